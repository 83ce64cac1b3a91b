"""tcgen05 (TF32 tensor-core) convolution kernels against the fp32 FFMA kernels and fp64 references.
Tolerance: TF32 keeps 10 mantissa bits of each operand (fp32 accumulate) -> ~1e-3 relative to the output scale."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TF32_TOL = 2e-3


@pytest.fixture()
def ops():
    from speechdrivestemplates_b200 import ops as o
    o.set_conv_math(1)
    yield o
    o.set_conv_math(0)


def dev():
    return torch.device("cuda:0")


def tc_launches():
    from speechdrivestemplates_b200 import _lib
    return _lib.call("sdt_tc_launches")


def to_cl(x):
    return x.permute(0, 2, 3, 1).contiguous()


def from_cl(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


# encoder layer geometries 0.1 .. 3.1 on reduced spatial sizes (odd widths, partial tiles) + one full-size layer
GEOMS = [
    # cin, cout, kh, kw, s, p, H, W, B
    (64, 64, 4, 4, 2, 1, 20, 37, 3),
    (64, 128, 3, 3, 1, 1, 10, 18, 3),
    (128, 128, 4, 4, 2, 1, 11, 19, 2),
    (128, 256, 3, 3, 1, 1, 10, 13, 2),
    (256, 256, 4, 4, 2, 1, 20, 27, 2),
    (256, 256, 3, 3, 1, 1, 10, 53, 2),
    (256, 256, 6, 3, 1, 0, 10, 53, 2),
    (64, 64, 4, 4, 2, 1, 80, 427, 2),
]


@pytest.mark.parametrize("cfg", GEOMS)
def test_tc_forward_with_loader_transform_and_stats(ops, cfg):
    cin, cout, kh, kw, s, p, H, W, B = cfg
    g = torch.Generator().manual_seed(21)
    x = torch.randn(B, cin, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, kh, kw, generator=g, dtype=torch.float64) / math.sqrt(cin * kh * kw)
    sc = torch.rand(B, cin, generator=g, dtype=torch.float64) + 0.5
    sh = torch.randn(B, cin, generator=g, dtype=torch.float64)
    a = F.leaky_relu(x * sc[:, :, None, None] + sh[:, :, None, None], 0.2)
    ref = F.conv2d(a, w, None, s, p)
    geom = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
    xf = (sc.float().to(dev()).contiguous(), sh.float().to(dev()).contiguous(), cin)
    n0 = tc_launches()
    y, partial = ops.conv_forward(to_cl(x.float()).to(dev()), w.float().to(dev()).contiguous(), geom, xf=xf, slope=0.2,
                                  want_stats=True, per_image=True)
    torch.cuda.synchronize()
    assert rel(from_cl(y), ref) < TF32_TOL
    assert tc_launches() == n0 + 1                       # the tcgen05 kernel ran (no silent FFMA fallback)
    tiles = partial.shape[0] // B
    ps = partial.view(B, tiles, 2, cout).double().sum(1).cpu()
    got = from_cl(y).double().cpu()
    assert rel(ps[:, 0], got.sum((2, 3))) < 1e-4          # the statistics describe the tensor that was stored
    assert rel(ps[:, 1], (got * got).sum((2, 3))) < 1e-4


@pytest.mark.parametrize("cfg", GEOMS[:7])
def test_tc_dgrad(ops, cfg):
    cin, cout, kh, kw, s, p, H, W, B = cfg
    g = torch.Generator().manual_seed(22)
    x = torch.randn(B, cin, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cout, cin, kh, kw, generator=g, dtype=torch.float64) / math.sqrt(cin * kh * kw)
    y = F.conv2d(x, w, None, s, p)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    geom = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
    n0 = tc_launches()
    dx = ops.conv_dgrad(to_cl(dy.float()).to(dev()), w.float().to(dev()).contiguous(), geom, H, W)
    torch.cuda.synchronize()
    assert rel(from_cl(dx), x.grad) < TF32_TOL
    assert tc_launches() == n0 + s * s                   # one tcgen05 launch per stride-parity class


def test_tc_conv1d_forward_bias_and_accumulating_dgrad(ops):
    B, L, cin, cout = 8, 64, 256, 256
    g = torch.Generator().manual_seed(23)
    x = torch.randn(B, cin, L, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cout, cin, 3, generator=g, dtype=torch.float64) / math.sqrt(cin * 3)
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    y = F.conv1d(x, w, b, 1, 1)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    geom = ops.ConvGeom.conv1d(cin, cout, 3, 1, 1)
    xcl = x.detach().float().permute(0, 2, 1).contiguous().view(B, 1, L, cin).to(dev())
    yk = ops.conv_forward(xcl, w.float().to(dev()).contiguous(), geom, bias=b.float().to(dev()))
    assert rel(yk.view(B, L, cout).permute(0, 2, 1), y) < TF32_TOL
    base = torch.ones(B, 1, L, cin, device=dev())
    dx = ops.conv_dgrad(dy.float().permute(0, 2, 1).contiguous().view(B, 1, L, cout).to(dev()), w.float().to(dev()).contiguous(),
                        geom, 1, L, out=base, accumulate=True)
    assert rel(dx.view(B, L, cin).permute(0, 2, 1), x.grad + 1.0) < TF32_TOL


@pytest.mark.parametrize("cfg", GEOMS)
@pytest.mark.parametrize("splits", [None, 1, 5])
def test_tc_wgrad(ops, cfg, splits):
    cin, cout, kh, kw, s, p, H, W, B = cfg
    if splits is not None and H * W > 2000:
        pytest.skip("one split setting is enough for the large case")
    g = torch.Generator().manual_seed(24)
    x = torch.randn(B, cin, H, W, generator=g, dtype=torch.float64)
    w = (torch.randn(cout, cin, kh, kw, generator=g, dtype=torch.float64) / math.sqrt(cin * kh * kw)).requires_grad_(True)
    sc = torch.rand(B, cin, generator=g, dtype=torch.float64) + 0.5
    sh = torch.randn(B, cin, generator=g, dtype=torch.float64)
    a = F.leaky_relu(x * sc[:, :, None, None] + sh[:, :, None, None], 0.2)
    y = F.conv2d(a, w, None, s, p)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    geom = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
    xf = (sc.float().to(dev()).contiguous(), sh.float().to(dev()).contiguous(), cin)
    n0 = tc_launches()
    dw = ops.conv_weight_grad(to_cl(x.float()).to(dev()), to_cl(dy.float()).to(dev()), geom, xf=xf, slope=0.2, splits=splits)
    torch.cuda.synchronize()
    assert tc_launches() == n0 + 1
    assert rel(dw, w.grad) < TF32_TOL


def test_tc_wgrad_conv1d(ops):
    B, L, cin, cout = 8, 64, 256, 256
    g = torch.Generator().manual_seed(25)
    x = torch.randn(B, cin, L, generator=g, dtype=torch.float64)
    w = (torch.randn(cout, cin, 4, generator=g, dtype=torch.float64) / math.sqrt(cin * 4)).requires_grad_(True)
    y = F.conv1d(x, w, None, 2, 1)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    geom = ops.ConvGeom.conv1d(cin, cout, 4, 2, 1)
    xcl = x.float().permute(0, 2, 1).contiguous().view(B, 1, L, cin).to(dev())
    dycl = dy.float().permute(0, 2, 1).contiguous().view(B, 1, -1, cout).to(dev())
    dw = ops.conv_weight_grad(xcl, dycl, geom)
    assert rel(dw.view(cout, cin, 4), w.grad) < TF32_TOL


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_sdt_bp_train_step_tf32_mode_vs_reference_fixture(mode):
    """Whole fused train step with the tcgen05 TF32 convolutions against the reference's recorded fp32 step.
    Stated TF32 tolerance: losses 1e-3, prediction 5e-3 of its range, final f64 results 5e-3; gradients agree in
    direction and size (relative L2 error < 0.25 per tensor: TF32 operand rounding moves ~0.1 % of the LeakyReLU units
    across zero, see tests/diag_grad_noise.py), the fp32 FFMA mode carries the tight gradient checks."""
    import numpy as np
    from oracle import sdt_oracle as O
    from speechdrivestemplates_b200 import _lib, config, ops as o, pipeline
    from test_gpu_step import _to_host_batch
    from util import golden, oliver_stat, rel_err
    g = golden("sdt_bp_step_golden")
    try:
        tr = pipeline.Voice2PoseTrainer(config.get_cfg("voice2pose_sdt_bp"), 16, dev(), use_cuda_graph=False, seed=0, conv_math=mode)
        tr.model.clips_code.data.copy_(0.1 * torch.randn(16, 32, generator=torch.Generator().manual_seed(11)))
        n0 = tc_launches()
        out = tr.train_step(_to_host_batch(O.synthetic_batch(2, 16, oliver_stat(True), seed=100)))
        host = tr.losses_to_host(out)
        assert tc_launches() - n0 >= 60                    # encoder + 1-D stacks ran on the tensor cores
        for k in ("G_reg_loss", "G_loss", "G_clipcode_kl_loss", "L2_dist", "lip_sync_error_n"):
            ref = float(g["step0/loss/" + k])
            assert abs(host[k] - ref) <= 1e-3 * max(1.0, abs(ref)), (k, host[k], ref)
        assert rel_err(out["poses_pred_batch"].cpu().numpy(), g["step0/pred"]) < 5e-3
        assert rel_err(out["final_pred"].cpu().numpy(), g["step0/final_pred"]) < 5e-3
        b0 = O.synthetic_batch(2, 16, oliver_stat(True), seed=100)                    # f64 path on identical f32 input: bit-exact
        st = b0["speaker_stat"]
        assert np.array_equal(out["final_gt"].cpu().numpy(),
                              O.get_final_results(b0["poses"].numpy(), st["mean"], st["std"], st["scale_factor"], True))
        # gradients: compare against the fp32 FFMA trainer on the same inputs
        tr0 = pipeline.Voice2PoseTrainer(config.get_cfg("voice2pose_sdt_bp"), 16, dev(), use_cuda_graph=False, seed=0, conv_math=0)
        tr0.model.clips_code.data.copy_(0.1 * torch.randn(16, 32, generator=torch.Generator().manual_seed(11)))
        tr0.train_step(_to_host_batch(O.synthetic_batch(2, 16, oliver_stat(True), seed=100)))
        for n in tr.grads:
            a, b = tr.grads[n].double().flatten(), tr0.grads[n].double().flatten()
            err = float((a - b).norm() / (b.norm() + 1e-30))
            cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
            assert err < 0.25 and cos > 0.97, (n, err, cos)
    finally:
        o.set_conv_math(0)


# ---------------------------------------------------------------- math modes 2 / 3: TMA-fed tcgen05 kernels
# (3 = operand reuse across vertical taps and accumulators in shared memory, csrc/tc_conv_ytap.cu)
@pytest.fixture(params=[2, 3, 4], ids=["tma", "reuse", "pair"])
def ops_tma(request):
    from speechdrivestemplates_b200 import ops as o
    o.set_conv_math(request.param)
    yield o
    o.set_conv_math(0)


# geometries that stress the sub-tile bookkeeping of mode 3: odd sub-tile counts (last CTA partially filled), maps much
# smaller than a patch, tall kernels, batch 1
REUSE_GEOMS = [
    # cin, cout, kh, kw, s, p, H, W, B
    (64, 64, 3, 3, 1, 1, 7, 9, 1),
    (64, 64, 3, 3, 1, 1, 40, 213, 1),
    (64, 128, 3, 3, 1, 1, 33, 50, 3),
    (128, 64, 4, 4, 2, 1, 30, 62, 5),
    (64, 256, 6, 3, 1, 0, 10, 53, 3),
    (128, 256, 3, 3, 1, 1, 17, 23, 2),
    (64, 64, 5, 5, 1, 2, 19, 21, 2),
    (64, 64, 4, 4, 2, 1, 21, 27, 3),          # odd input: the four parity classes of the data gradient have different grids
]


@pytest.mark.parametrize("cfg", REUSE_GEOMS)
def test_reuse_forward_and_dgrad_odd_geometries(cfg):
    from speechdrivestemplates_b200 import ops
    cin, cout, kh, kw, s, p, H, W, B = cfg
    g = torch.Generator().manual_seed(41)
    x = torch.randn(B, cin, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cout, cin, kh, kw, generator=g, dtype=torch.float64) / math.sqrt(cin * kh * kw)
    y = F.conv2d(x, w, None, s, p)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    geom = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
    ops.set_conv_math(3)
    try:
        n0 = tc_launches()
        yk, partial = ops.conv_forward(to_cl(x.detach().float()).to(dev()), w.float().to(dev()).contiguous(), geom, want_stats=True,
                                       per_image=True)
        dx = ops.conv_dgrad(to_cl(dy.float()).to(dev()), w.float().to(dev()).contiguous(), geom, H, W)
        torch.cuda.synchronize()
        assert tc_launches() > n0
    finally:
        ops.set_conv_math(0)
    assert rel(from_cl(yk), y) < TF32_TOL
    assert rel(from_cl(dx), x.grad) < TF32_TOL
    tiles = partial.shape[0] // B
    ps = partial.view(B, tiles, 2, cout).double().sum(1).cpu()
    got = from_cl(yk).double().cpu()
    assert rel(ps[:, 0], got.sum((2, 3))) < 1e-4
    assert rel(ps[:, 1], (got * got).sum((2, 3))) < 1e-4


def _reuse_case(cfg, seed=41):
    """Forward (+ statistics partials) and data gradient of one layer in math mode 3 against float64 torch."""
    from speechdrivestemplates_b200 import ops
    cin, cout, kh, kw, s, p, H, W, B = cfg
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cin, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cout, cin, kh, kw, generator=g, dtype=torch.float64) / math.sqrt(cin * kh * kw)
    y = F.conv2d(x, w, None, s, p)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    geom = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
    n0 = tc_launches()
    yk, partial = ops.conv_forward(to_cl(x.detach().float()).to(dev()), w.float().to(dev()).contiguous(), geom, want_stats=True,
                                   per_image=True)
    dx = ops.conv_dgrad(to_cl(dy.float()).to(dev()), w.float().to(dev()).contiguous(), geom, H, W)
    torch.cuda.synchronize()
    assert tc_launches() > n0
    assert rel(from_cl(yk), y) < TF32_TOL
    assert rel(from_cl(dx), x.grad) < TF32_TOL
    tiles = partial.shape[0] // B
    ps = partial.view(B, tiles, 2, cout).double().sum(1).cpu()
    got = from_cl(yk).double().cpu()
    assert rel(ps[:, 0], got.sum((2, 3))) < 1e-4
    assert rel(ps[:, 1], (got * got).sum((2, 3))) < 1e-4


SPAN_CASES = [
    # (cin, cout, kh, kw, s, p, H, W, B), forced (N tile, accumulators, patch width, images per patch)
    ((64, 128, 3, 3, 1, 1, 10, 53, 9), (0, 0, 8, 8)),          # 16 statistics runs per patch, B not a multiple of nb
    ((64, 128, 3, 3, 1, 1, 10, 53, 9), (128, 2, 8, 4)),
    ((64, 64, 4, 4, 2, 1, 20, 106, 6), (64, 2, 8, 8)),         # stride 2: two y phases; dgrad = 4 parity classes
    ((64, 64, 4, 4, 2, 1, 20, 106, 6), (64, 2, 16, 4)),        # 8 runs per patch
    ((128, 64, 3, 3, 1, 1, 20, 26, 5), (0, 0, 32, 2)),         # runs of 32 rows: one image per epilogue warp
    ((64, 64, 6, 3, 1, 0, 10, 53, 4), (64, 1, 32, 2)),         # six vertical taps, no padding
    ((64, 64, 3, 3, 1, 1, 5, 13, 21), (0, 0, 8, 16)),          # bh = 1, 16 images per patch
    ((64, 64, 5, 5, 1, 2, 19, 21, 3), (0, 0, 8, 2)),
]


@pytest.mark.parametrize("cfg,force", SPAN_CASES)
def test_reuse_image_spanning_patches(cfg, force):
    """Patches that take their pixels from several consecutive images (tensor map ordered (C, W, B, H)): results, masking of
    the images past the batch, and the per-image statistics rows."""
    import ctypes
    from speechdrivestemplates_b200 import _lib, ops
    lib = _lib.load()
    ops.set_conv_math(3)
    lib.sdt_debug_conv_force(*force)
    try:
        cin, cout, kh, kw, s, p, H, W, B = cfg
        geom = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
        x = torch.zeros(B, H, W, cin, device=dev())
        oh, ow = geom.out_hw(H, W)
        d = ops.fwd_desc(geom, x, torch.empty(geom.k, cout, device=dev()), torch.empty(B, oh, ow, cout, device=dev()), B, H, W,
                         None, 1.0, None, None, True, wt_nk=torch.empty(cout, geom.k, device=dev()))
        plan = (ctypes.c_int32 * 10)()
        assert lib.sdt_conv_plan(ctypes.byref(d), plan) == 0
        assert plan[0] == 3 and plan[3] >> 8 == force[3] and plan[4] == force[2], list(plan)
        _reuse_case(cfg, seed=43)
    finally:
        lib.sdt_debug_conv_force(0, 0, 0, 0)
        ops.set_conv_math(0)


@pytest.mark.parametrize("cfg", [(128, 64, 4, 4, 2, 1, 30, 62, 5), (64, 64, 4, 4, 2, 1, 21, 27, 3), (256, 256, 4, 4, 2, 1, 20, 106, 8)])
def test_parity_classes_in_one_launch_match_separate_launches(cfg):
    """sdt_conv_gemm_multi: the four stride-parity classes of a data gradient as ONE persistent launch give bit-identical
    results to four launches (same K order per output element), with one kernel instead of four."""
    from speechdrivestemplates_b200 import _lib, ops
    cin, cout, kh, kw, s, p, H, W, B = cfg
    g = torch.Generator().manual_seed(45)
    geom = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
    oh, ow = geom.out_hw(H, W)
    dy = torch.randn(B, oh, ow, cout, generator=g).to(dev())
    w = (torch.randn(cout, cin, kh, kw, generator=g) / math.sqrt(cin * kh * kw)).to(dev())
    ops.set_conv_math(3)
    try:
        descs, keep = [], []
        for cls in geom.dgrad_classes(H, W):
            wt_nk = torch.empty(geom.cin, cls["th"] * cls["tw"] * geom.cout, device=dev())
            ops.weight_prep_dgrad_nk(w, geom, cls, wt_nk)
            keep.append(wt_nk)
            descs.append((cls, wt_nk))
        dx1 = torch.full((B, H, W, cin), float("nan"), device=dev())
        for cls, wt_nk in descs:
            ops.conv_gemm(ops.dgrad_desc(geom, cls, dy, None, dx1, B, H, W, wt_nk=wt_nk))
        dx2 = torch.full((B, H, W, cin), float("nan"), device=dev())
        n0, k0 = tc_launches(), _lib.launch_count
        ops.conv_gemm_multi([ops.dgrad_desc(geom, cls, dy, None, dx2, B, H, W, wt_nk=wt_nk) for cls, wt_nk in descs])
        torch.cuda.synchronize()
        assert tc_launches() - n0 == 1 and _lib.launch_count - k0 == 1
    finally:
        ops.set_conv_math(0)
    assert torch.isfinite(dx2).all()
    assert torch.equal(dx1, dx2)


def test_reuse_matches_tma_mode_bitwise_products():
    """Modes 2 and 3 multiply the same TF32-rounded operands; only the fp32 accumulation order differs."""
    from speechdrivestemplates_b200 import ops
    cin, cout, kh, kw, s, p, H, W, B = GEOMS[1]
    g = torch.Generator().manual_seed(42)
    x = to_cl(torch.randn(B, cin, H, W, generator=g)).to(dev())
    w = (torch.randn(cout, cin, kh, kw, generator=g) / math.sqrt(cin * kh * kw)).to(dev())
    geom = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
    outs = []
    for mode in (2, 3):
        ops.set_conv_math(mode)
        try:
            outs.append(ops.conv_forward(x, w, geom).clone())
        finally:
            ops.set_conv_math(0)
    assert rel(outs[1], outs[0]) < 2e-6


@pytest.mark.parametrize("cfg", GEOMS)
def test_tma_forward_plain_source_with_stats(ops_tma, cfg):
    ops = ops_tma
    cin, cout, kh, kw, s, p, H, W, B = cfg
    g = torch.Generator().manual_seed(31)
    x = torch.randn(B, cin, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, kh, kw, generator=g, dtype=torch.float64) / math.sqrt(cin * kh * kw)
    ref = F.conv2d(x, w, None, s, p)
    geom = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
    n0 = tc_launches()
    y, partial = ops.conv_forward(to_cl(x.float()).to(dev()), w.float().to(dev()).contiguous(), geom, want_stats=True, per_image=True)
    torch.cuda.synchronize()
    assert tc_launches() == n0 + 1
    assert rel(from_cl(y), ref) < TF32_TOL
    tiles = partial.shape[0] // B
    ps = partial.view(B, tiles, 2, cout).double().sum(1).cpu()
    got = from_cl(y).double().cpu()
    assert rel(ps[:, 0], got.sum((2, 3))) < 1e-4           # patch rows outside the output grid are masked out of the statistics
    assert rel(ps[:, 1], (got * got).sum((2, 3))) < 1e-4


@pytest.mark.parametrize("cfg", GEOMS[:7])
def test_tma_dgrad(ops_tma, cfg):
    ops = ops_tma
    cin, cout, kh, kw, s, p, H, W, B = cfg
    g = torch.Generator().manual_seed(32)
    x = torch.randn(B, cin, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cout, cin, kh, kw, generator=g, dtype=torch.float64) / math.sqrt(cin * kh * kw)
    y = F.conv2d(x, w, None, s, p)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    geom = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
    dx = ops.conv_dgrad(to_cl(dy.float()).to(dev()), w.float().to(dev()).contiguous(), geom, H, W)
    torch.cuda.synchronize()
    assert rel(from_cl(dx), x.grad) < TF32_TOL


@pytest.mark.parametrize("k,s,L", [(3, 1, 64), (4, 2, 64), (4, 2, 33), (3, 1, 2)])
def test_tma_conv1d(ops_tma, k, s, L):
    ops = ops_tma
    B, cin, cout = 8, 256, 256
    g = torch.Generator().manual_seed(33)
    x = torch.randn(B, cin, L, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cout, cin, k, generator=g, dtype=torch.float64) / math.sqrt(cin * k)
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    y = F.conv1d(x, w, b, s, 1)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    geom = ops.ConvGeom.conv1d(cin, cout, k, s, 1)
    xcl = x.detach().float().permute(0, 2, 1).contiguous().view(B, 1, L, cin).to(dev())
    yk = ops.conv_forward(xcl, w.float().to(dev()).contiguous(), geom, bias=b.float().to(dev()))
    assert rel(yk.view(B, -1, cout).permute(0, 2, 1), y) < TF32_TOL
    base = torch.ones(B, 1, L, cin, device=dev())
    dx = ops.conv_dgrad(dy.float().permute(0, 2, 1).contiguous().view(B, 1, -1, cout).to(dev()), w.float().to(dev()).contiguous(),
                        geom, 1, L, out=base, accumulate=True)
    assert rel(dx.view(B, L, cin).permute(0, 2, 1), x.grad + 1.0) < TF32_TOL


@pytest.mark.parametrize("cfg", GEOMS)
@pytest.mark.parametrize("splits", [None, 3])
def test_tma_wgrad_plain_source(ops_tma, cfg, splits):
    ops = ops_tma
    cin, cout, kh, kw, s, p, H, W, B = cfg
    if splits is not None and H * W > 2000:
        pytest.skip("one split setting is enough for the large case")
    g = torch.Generator().manual_seed(34)
    x = torch.randn(B, cin, H, W, generator=g, dtype=torch.float64)
    w = (torch.randn(cout, cin, kh, kw, generator=g, dtype=torch.float64) / math.sqrt(cin * kh * kw)).requires_grad_(True)
    y = F.conv2d(x, w, None, s, p)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    geom = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
    n0 = tc_launches()
    dw = ops.conv_weight_grad(to_cl(x.float()).to(dev()), to_cl(dy.float()).to(dev()), geom, splits=splits)
    torch.cuda.synchronize()
    assert tc_launches() == n0 + 1
    assert rel(dw, w.grad) < TF32_TOL


@pytest.mark.parametrize("k,s,L,B", [(3, 1, 64, 8), (4, 2, 64, 8), (4, 2, 4, 32), (3, 1, 2, 5)])
def test_tma_wgrad_conv1d_small_maps(ops_tma, k, s, L, B):
    ops = ops_tma
    cin, cout = 256, 256
    g = torch.Generator().manual_seed(35)
    x = torch.randn(B, cin, L, generator=g, dtype=torch.float64)
    w = (torch.randn(cout, cin, k, generator=g, dtype=torch.float64) / math.sqrt(cin * k)).requires_grad_(True)
    y = F.conv1d(x, w, None, s, 1)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    geom = ops.ConvGeom.conv1d(cin, cout, k, s, 1)
    xcl = x.float().permute(0, 2, 1).contiguous().view(B, 1, L, cin).to(dev())
    dycl = dy.float().permute(0, 2, 1).contiguous().view(B, 1, -1, cout).to(dev())
    dw = ops.conv_weight_grad(xcl, dycl, geom)
    assert rel(dw.view(cout, cin, k), w.grad) < TF32_TOL
