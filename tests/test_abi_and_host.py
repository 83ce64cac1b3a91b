"""CPU-side checks: the C-ABI library loads and exports every symbol include/sdt_b200.h declares (no compute calls
without a GPU), the ctypes mirror of sdt_conv_desc matches the header, host-side geometry logic, drop-in state-dict
layout, loud failure without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    with open(os.path.join(ROOT, "include", "sdt_b200.h")) as f:
        return f.read()


def _built():
    import __graft_entry__ as ge
    ge.build()


def test_library_exports_every_declared_symbol():
    _built()
    from speechdrivestemplates_b200 import _lib
    lib = _lib.load()
    declared = set(re.findall(r"^(?:const char\*|int|int64_t)\s+(sdt_\w+)\s*\(", _header(), re.M))
    assert len(declared) >= 35
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sdt_version() >= 100
    assert lib.sdt_get_conv_math() in (0, 1, 2, 3, 4)


def test_conv_desc_mirror_matches_header():
    from speechdrivestemplates_b200._lib import ConvDesc
    body = re.search(r"typedef struct sdt_conv_desc \{(.*?)\} sdt_conv_desc;", _header(), re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(float\*|int32_t|float)\s*", "", decl)
        names += [n.strip().lstrip("*") for n in decl.split(",")]
    assert names == [f[0] for f in ConvDesc._fields_]
    assert ctypes.sizeof(ConvDesc) == 10 * 8 + 28 * 4 + 3 * 8 + 4 * 4    # 27 scalars + pad, 3 pointers, 3 scalars + tail padding


def test_argument_errors_are_reported_not_thrown():
    _built()
    from speechdrivestemplates_b200 import _lib
    with pytest.raises(_lib.SdtError, match="null pointer"):
        _lib.call("sdt_mel_fwd", None, 1, 1000, None, None, None, None, 1, None, None)
    with pytest.raises(_lib.SdtError, match="mode must be"):
        _lib.call("sdt_set_conv_math", 7)


def test_dgrad_parity_classes_cover_every_input_once():
    from speechdrivestemplates_b200.ops import ConvGeom
    for g, h, w in [(ConvGeom.conv2d(64, 64, 4, 4, 2, 1), 80, 427), (ConvGeom.conv2d(8, 8, 3, 3, 1, 1), 5, 7),
                    (ConvGeom.conv1d(256, 256, 4, 2, 1), 1, 63), (ConvGeom.conv2d(4, 4, 6, 3, 1, 0), 10, 53)]:
        seen = torch.zeros(h, w, dtype=torch.int32)
        taps = 0
        for c in g.dgrad_classes(h, w):
            ys = torch.arange(c["gh"]) * g.sh + c["py"]
            xs = torch.arange(c["gw"]) * g.sw + c["px"]
            assert ys.max() < h and xs.max() < w
            seen[ys[:, None], xs[None, :]] += 1
            taps += c["th"] * c["tw"]
        assert bool((seen == 1).all())
        assert taps == g.kh * g.kw                 # every kernel tap belongs to exactly one class


def test_dropin_state_dict_layout_and_cpu_refusal():
    from speechdrivestemplates_b200 import config, pipeline
    from util import golden
    cfg = config.get_cfg("voice2pose_sdt_bp")
    torch.manual_seed(0)
    m = pipeline.Voice2PoseModel(cfg, num_train_samples=16)
    g = golden("sdt_bp_zero_code_golden")
    ref_keys = sorted(k[len("init/"):-len("/digest")] for k in g.files if k.startswith("init/") and k.endswith("/digest"))
    assert sorted(m.state_dict().keys()) == ref_keys
    assert m.pose_encoder.training is False                 # voice2pose.py:77
    with pytest.raises(RuntimeError):                       # no CPU fallback: fails loudly
        m.netG(torch.zeros(1, 80, 427), 64, torch.zeros(1, 32))
    with pytest.raises(KeyError, match="Unknown model"):
        from speechdrivestemplates_b200 import networks
        networks.get_model("NoSuchNet")


def test_config_overlays_match_reference_yaml_when_available():
    ref = "/root/reference/configs"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    import yaml
    from speechdrivestemplates_b200 import config
    for name in config.OVERLAYS:
        with open(os.path.join(ref, name + ".yaml")) as f:
            y = yaml.safe_load(f)
        cfg = config.get_cfg(name)

        def walk(node, d):
            for k, v in d.items():
                if isinstance(v, dict):
                    walk(node[k], v)
                else:
                    assert float(node[k]) == float(v) if isinstance(v, str) and k == "LR" else node[k] == v, (name, k)
        walk(cfg, y)


def test_plugin_registers_into_reference_registries():
    """The drop-in seam itself: with the reference importable, plugin.register() makes the reference's own
    get_model / pipeline module resolve to the CUDA-backed classes, and they construct from the reference's yacs config
    with the reference's state-dict layout."""
    sys_path_ref = "/root/reference"
    if not os.path.isdir(os.path.join(sys_path_ref, "core", "networks")):
        pytest.skip("reference tree not present")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import refshim
    from speechdrivestemplates_b200 import networks, pipeline, plugin
    cfg = refshim.get_cfg("voice2pose_s2g")          # installs the shims, puts the reference on sys.path
    done = plugin.register()
    import core.networks as ref_networks
    import core.pipelines.voice2pose as ref_v2p
    import core.pipelines.pose2pose as ref_p2p
    assert set(done) >= {"SequenceGeneratorCNN", "PoseSequenceDiscriminator", "Autoencoder", "PoseSeqEncoder", "Voice2PoseModel"}
    assert ref_networks.get_model("SequenceGeneratorCNN") is networks.SequenceGeneratorCNN
    assert ref_v2p.Voice2PoseModel is pipeline.Voice2PoseModel and ref_p2p.Pose2PoseModel is pipeline.Pose2PoseModel
    with pytest.raises(KeyError, match="Unknown model"):
        ref_networks.get_model("Nope")
    torch.manual_seed(0)
    m = ref_v2p.Voice2PoseModel(cfg, num_train_samples=4)       # built from the reference's own (shimmed yacs) config object
    from util import golden
    g = golden("s2g_step_golden")
    ref_keys = sorted(k[len("init/"):-len("/digest")] for k in g.files if k.startswith("init/") and k.endswith("/digest"))
    assert sorted(m.state_dict().keys()) == ref_keys
    m2 = ref_p2p.Pose2PoseModel(refshim.get_cfg("pose2pose"), num_train_samples=4)
    assert "clip_code_mu" in m2.state_dict() and "ae.decoder.blocks.4.bias" in m2.state_dict()
    # the pipeline registry (core/pipelines/__init__.py:5-16): get_pipeline resolves to subclasses of the reference's own classes
    # that override exactly the three methods the fused trainers replace; registering twice does not stack subclasses
    import core.pipelines as ref_pipelines
    assert {"pipeline:Voice2Pose", "pipeline:Pose2Pose"} <= set(done)
    for name, base in (("Voice2Pose", ref_v2p.Voice2Pose), ("Pose2Pose", ref_p2p.Pose2Pose)):
        cls = ref_pipelines.get_pipeline(name)
        assert cls is not base and issubclass(cls, base) and cls.__mro__[1] is base
        assert {k for k in vars(cls) if not k.startswith("_") and callable(vars(cls)[k])} == {"setup_model", "setup_optimizer", "train_step"}
        import inspect
        for meth in ("setup_model", "setup_optimizer", "train_step"):
            assert inspect.signature(getattr(cls, meth)) == inspect.signature(getattr(base, meth)), (name, meth)
    plugin.register()
    assert ref_pipelines.get_pipeline("Voice2Pose").__mro__[1] is ref_v2p.Voice2Pose
    with pytest.raises(KeyError, match="Unknown pipeline"):
        ref_pipelines.get_pipeline("Nope")


def test_multistep_handle_follows_torch_multisteplr():
    """pipelines.MultiStepHandle == torch.optim.lr_scheduler.MultiStepLR on the reference's milestones (voice2pose.py:251-257)."""
    from speechdrivestemplates_b200 import pipelines

    class FakeTrainer:
        lr = None

        def set_lr(self, lr):
            self.lr = lr
    epochs, base = 14, 1e-4
    ms = [epochs - 10, epochs - 2]
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=base)
    ref = torch.optim.lr_scheduler.MultiStepLR(opt, ms, gamma=0.1, last_epoch=-1)
    tr = FakeTrainer()
    h = pipelines.MultiStepHandle(tr, base, ms, 0.1, -1)
    follower = pipelines.MultiStepHandle(tr, base, ms, 0.1, -1, drives=False)
    assert tr.lr == pytest.approx(opt.param_groups[0]["lr"])
    for _ in range(epochs):
        opt.step()
        ref.step()
        h.step()
        follower.step()
        assert tr.lr == pytest.approx(opt.param_groups[0]["lr"], rel=1e-12)
        assert h.get_last_lr()[0] == pytest.approx(ref.get_last_lr()[0], rel=1e-12)
    assert tr.lr == pytest.approx(base * 1e-2)
    # resume at epoch 5 (setup_experiment passes last_epoch=epoch, trainer.py:184): already past the first milestone
    tr2 = FakeTrainer()
    pipelines.MultiStepHandle(tr2, base, ms, 0.1, last_epoch=5)
    assert tr2.lr == pytest.approx(base * 0.1)


def test_default_conv_math_precedence(monkeypatch):
    from speechdrivestemplates_b200 import config, pipelines
    cfg = config.get_cfg("voice2pose_sdt_bp")
    monkeypatch.delenv("SDT_CONV_MATH", raising=False)
    assert pipelines.default_conv_math(cfg) == 3
    cfg.SYS["SDT_CONV_MATH"] = 0
    assert pipelines.default_conv_math(cfg) == 0
    monkeypatch.setenv("SDT_CONV_MATH", "2")
    assert pipelines.default_conv_math(cfg) == 2


# ---------------------------------------------------------------- tile planner of the mode-3 convolution (host only)
ENC_LAYERS = [(64, 64, 4, 4, 2, 1, 80, 427), (64, 128, 3, 3, 1, 1, 40, 213), (128, 128, 4, 4, 2, 1, 40, 213),
              (128, 256, 3, 3, 1, 1, 20, 106), (256, 256, 4, 4, 2, 1, 20, 106), (256, 256, 3, 3, 1, 1, 10, 53),
              (256, 256, 6, 3, 1, 0, 10, 53)]


def _fake_fwd_desc(ops, cin, cout, kh, kw, s, p, H, W, B):
    g = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
    oh, ow = g.out_hw(H, W)
    d = ops.ConvDesc()
    d.src = d.wt = d.wt_nk = d.dst = 0x1000                 # never dereferenced: sdt_conv_plan is host-only
    d.B, d.SH, d.SW, d.C = B, H, W, cin
    d.GH, d.GW, d.TH, d.TW = oh, ow, kh, kw
    d.y_mul, d.ty_mul, d.y_off = s, 1, -p
    d.x_mul, d.tx_mul, d.x_off = s, 1, -p
    d.N, d.DH, d.DW, d.dy_mul, d.dy_off, d.dx_mul, d.dx_off = cout, oh, ow, 1, 0, 1, 0
    d.per_image_tiles = 1
    return d, oh, ow


@pytest.mark.parametrize("B", [1, 2, 32, 128])
def test_conv_plan_for_the_encoder_layers(B):
    """Math mode 3 plans every 2-D encoder layer on the persistent tcgen05 kernel, within the hardware limits: TMEM columns
    (2 accumulator sets), shared memory of one CTA per SM, TMA box extents, and enough sub-tiles to cover the output grid."""
    _built()
    from speechdrivestemplates_b200 import _lib, ops
    lib = _lib.load()
    assert lib.sdt_set_conv_math(3) == 0
    try:
        for cfg in ENC_LAYERS:
            d, oh, ow = _fake_fwd_desc(ops, *cfg, B)
            out = (ctypes.c_int32 * 10)()
            assert lib.sdt_conv_plan(ctypes.byref(d), out) == 0
            kind, bn, mt, bh_nb, bw, box_rows, a_st, b_st, smem, tiles = list(out)
            bh, nb = bh_nb & 255, bh_nb >> 8                     # patch rows, images per patch
            assert kind == 3, (cfg, list(out))
            assert bn in (64, 128) and mt in (1, 2, 4) and 2 * mt * bn <= 512
            assert bh * bw * nb == 128 and bw % 8 == 0 and nb in (1, 2, 4, 8, 16) and nb < 2 * B
            assert smem <= 227 * 1024 and a_st >= 2 and b_st >= 2
            assert bw * cfg[4] <= 256 and box_rows * cfg[4] <= 256 and box_rows >= bh
            tpi = (-(-oh // bh)) * (-(-ow // bw))
            sub = -(-B // nb) * tpi
            assert tiles == -(-sub // mt) * (cfg[1] // bn)
            if torch.cuda.is_available():                        # needs the driver's cuTensorMapEncodeTiled to be eligible
                assert lib.sdt_conv_row_tiles(ctypes.byref(d)) == B * tpi  # one statistics row per (image, patch position)
    finally:
        lib.sdt_set_conv_math(0)


def test_conv_plan_invariants_on_random_layers():
    """Property test of the mode-3 planner (host only): whatever layer it accepts, the plan respects the hardware limits
    (shared memory of one CTA per SM, 512 TMEM columns for two accumulator sets, TMA box extents <= 256) and tiles the
    whole output grid; image-spanning patches never leave more than one half-empty group of images."""
    _built()
    import random
    from speechdrivestemplates_b200 import _lib, ops
    lib = _lib.load()
    rng = random.Random(7)
    assert lib.sdt_set_conv_math(3) == 0
    planned = spanning = 0
    try:
        for _ in range(300):
            cin, cout = rng.choice([32, 64, 96, 128, 256]), rng.choice([64, 128, 192, 256])
            k, s = rng.choice([(3, 1), (4, 2), (5, 1), (1, 1), (2, 2), (6, 1)])
            kh, kw = k, rng.choice([k, 3]) if s == 1 else k
            p = rng.choice([0, 1, k // 2])
            H, W, B = rng.randint(max(kh, 2), 90), rng.randint(max(kw, 2), 450), rng.choice([1, 2, 3, 7, 32, 33, 128])
            if (H + 2 * p - kh) // s + 1 < 2 or (W + 2 * p - kw) // s + 1 < 1:
                continue
            d, oh, ow = _fake_fwd_desc(ops, cin, cout, kh, kw, s, p, H, W, B)
            out = (ctypes.c_int32 * 10)()
            assert lib.sdt_conv_plan(ctypes.byref(d), out) == 0
            kind, bn, mt, bh_nb, bw, box_rows, a_st, b_st, smem, tiles = list(out)
            if kind != 3:
                continue
            planned += 1
            bh, nb = bh_nb & 255, bh_nb >> 8
            spanning += nb > 1
            assert cout % bn == 0 and bn in (64, 128) and mt in (1, 2, 4) and 2 * mt * bn <= 512
            assert bh * bw * nb == 128 and bw % 8 == 0 and (nb == 1 or nb // 2 < B)
            assert smem <= 227 * 1024 and 2 <= a_st <= 4 and 2 <= b_st <= 6
            assert bw * s <= 256 and box_rows * s <= 256 and box_rows >= bh
            sub = -(-B // nb) * (-(-oh // bh)) * (-(-ow // bw))
            assert tiles == -(-sub // mt) * (cout // bn)
            # ring + epilogue scratch as the kernel lays them out
            slots = 4 if (nb == 1 or bw >= 32) else 128 // bw
            epi = 4 * 32 * 36 * 4 + (1 if slots > 4 else mt) * 2 * slots * bn * 4       # 4 epilogue warps x 32 staged rows
            assert smem == a_st * mt * box_rows * nb * bw * 128 + b_st * bn * 128 + epi + (2 * 4 + 2 * 6 + 4) * 8 + 16
    finally:
        lib.sdt_set_conv_math(0)
    assert planned > 100 and spanning > 10


def test_conv_gemm_multi_rejects_bad_arguments_without_launching():
    """sdt_conv_gemm_multi (the parity classes of one data gradient as one launch): argument errors are reported through the
    status code + sdt_last_error(), before anything touches the GPU."""
    _built()
    from speechdrivestemplates_b200 import _lib, ops
    with pytest.raises(_lib.SdtError, match="sdt_conv_gemm_multi"):
        _lib.call("sdt_conv_gemm_multi", None, 2, None, None)
    arr = (ops.ConvDesc * 17)()
    with pytest.raises(_lib.SdtError, match="17 problems"):
        _lib.call("sdt_conv_gemm_multi", arr, 17, None, None)
    with pytest.raises(_lib.SdtError):
        _lib.call("sdt_conv_gemm_multi", arr, 0, None, None)


def test_conv_plan_falls_back_outside_mode_3_and_remaps_1d():
    _built()
    from speechdrivestemplates_b200 import _lib, ops
    lib = _lib.load()
    out = (ctypes.c_int32 * 10)()
    d, _, _ = _fake_fwd_desc(ops, *ENC_LAYERS[1], 4)
    for mode in (0, 1, 2):
        lib.sdt_set_conv_math(mode)
        assert lib.sdt_conv_plan(ctypes.byref(d), out) == 0 and out[0] != 3
    lib.sdt_set_conv_math(3)
    try:
        d1 = ops.ConvDesc()                                  # a 1-D layer: the batch becomes the image height for the planner
        d1.src = d1.wt = d1.wt_nk = d1.dst = 0x1000
        d1.B, d1.SH, d1.SW, d1.C = 4, 1, 64, 256
        d1.GH, d1.GW, d1.TH, d1.TW = 1, 64, 1, 3
        d1.y_mul, d1.ty_mul, d1.y_off, d1.x_mul, d1.tx_mul, d1.x_off = 1, 1, 0, 1, 1, -1
        d1.N, d1.DH, d1.DW, d1.dy_mul, d1.dx_mul = 256, 1, 64, 1, 1
        if os.environ.get("SDT_REMAP_1D") == "1":           # experimental switch (slower at B = 32, see conv_gemm.cu)
            assert lib.sdt_conv_plan(ctypes.byref(d1), out) == 0 and out[0] == 3 and (out[3] & 255) * (out[3] >> 8) * out[4] == 128
        else:
            assert lib.sdt_conv_plan(ctypes.byref(d1), out) == 0 and out[0] != 3
        d1.B = 1                                             # a single clip has nothing to stack: TMA kernel of mode 2
        assert lib.sdt_conv_plan(ctypes.byref(d1), out) == 0 and out[0] != 3
        with pytest.raises(_lib.SdtError):
            _lib.call("sdt_conv_plan", ctypes.byref(d1), None)
    finally:
        lib.sdt_set_conv_math(0)


def test_wgrad_splits_keep_at_least_256_pixels_per_split():
    from speechdrivestemplates_b200 import ops
    g1d = ops.ConvGeom.conv1d(256, 256, 3, 1, 1)
    assert ops.wgrad_splits(g1d, 32, 1, 64) == 8             # 2048 pixels
    assert ops.wgrad_splits(g1d, 1, 1, 2) == 1
    g2d = ops.ConvGeom.conv2d(64, 128, 3, 3, 1, 1)
    s = ops.wgrad_splits(g2d, 32, 40, 213)
    assert 1 <= s <= 32 * 40 * 213 // 256 and s * 5 >= 400   # ~3 CTAs per SM
    assert ops.bwd_tiles(255, 32) >= 8 and ops.bwd_tiles(255, 32) * 32 >= 255 // 8
    assert ops.bwd_tiles(40 * 213, 32) == -(-40 * 213 // 256)


@pytest.mark.parametrize("name", ["voice2pose_sdt_bp", "voice2pose_s2g", "pose2pose"])
def test_checkpoint_layout_loads_into_the_reference_classes_strictly(name):
    """a model_state_dict in the layout checkpoint.py writes ('module.' + the drop-in's state-dict keys) loads into
    the REFERENCE's own step model with strict=True, the reference's state dict loads back into the drop-in, and the parameter ORDER
    (which indexes torch.optim.Adam's per-parameter state in the checkpoint, trainer.py:318-319) is the same on both sides."""
    if not os.path.isdir("/root/reference/core/pipelines"):
        pytest.skip("reference tree not present")
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import refshim
    from speechdrivestemplates_b200 import pipeline
    cfg = refshim.get_cfg(name)
    import core.networks as ref_networks
    import core.pipelines.pose2pose as ref_p2p
    import core.pipelines.voice2pose as ref_v2p
    importlib.reload(ref_networks)                 # undo a plugin.register() of an earlier test: the reference's own classes
    importlib.reload(ref_v2p)
    importlib.reload(ref_p2p)
    if name == "pose2pose":
        RefModel, Own = ref_p2p.Pose2PoseModel, pipeline.Pose2PoseModel
    else:
        RefModel, Own = ref_v2p.Voice2PoseModel, pipeline.Voice2PoseModel
    assert RefModel.__module__.startswith("core.pipelines")
    torch.manual_seed(1)
    ref = RefModel(cfg, None, 6, 0)
    torch.manual_seed(2)
    own = Own(cfg, num_train_samples=6)
    written = {"module." + k: v.detach().clone() for k, v in own.state_dict().items()}         # checkpoint.voice2pose_checkpoint's layout
    # (no GPU here, so no DataParallel wrapper around the reference model: strip the prefix by hand)
    missing, unexpected = ref.load_state_dict({k[len("module."):]: v for k, v in written.items()}, strict=True)
    assert not missing and not unexpected
    for k, v in ref.state_dict().items():
        assert torch.equal(v, own.state_dict()[k]), k
    torch.manual_seed(3)
    ref2 = RefModel(cfg, None, 6, 0)
    own.load_state_dict(ref2.state_dict(), strict=True)
    for k, v in own.state_dict().items():
        assert torch.equal(v, ref2.state_dict()[k]), k
    sub = "ae" if name == "pose2pose" else "netG"
    assert [n for n, _ in getattr(own, sub).named_parameters()] == [n for n, _ in getattr(ref, sub).named_parameters()]
    if name == "voice2pose_s2g":
        assert [n for n, _ in own.netD_pose.named_parameters()] == [n for n, _ in ref.netD_pose.named_parameters()]


def test_p2p_allreduce_rejects_bad_arguments_without_launching():
    """sdt_p2p_allreduce (csrc/p2p.cu) validates rank / world / alignment / element count before touching the device."""
    _built()
    from speechdrivestemplates_b200 import _lib
    lib = _lib.load()
    ptrs = (ctypes.c_uint64 * 2)(0x10000, 0x20000)
    assert lib.sdt_p2p_allreduce(ptrs, 0, 2, 2, 1024, None, None, 0, None) != 0 and "rank" in _lib.last_error()
    assert lib.sdt_p2p_allreduce(ptrs, 0, 0, 2, 1022, None, None, 0, None) != 0 and "multiple of 4" in _lib.last_error()
    bad = (ctypes.c_uint64 * 2)(0x10000, 0x20004)
    assert lib.sdt_p2p_allreduce(bad, 0, 0, 2, 1024, None, None, 0, None) != 0 and "aligned" in _lib.last_error()
    assert lib.sdt_p2p_allreduce(ptrs, 0, 0, 2, 1024, None, None, 8, None) != 0 and "scalar" in _lib.last_error()
    assert lib.sdt_p2p_allreduce(ptrs, 0, 0, 17, 1024, None, None, 0, None) != 0
