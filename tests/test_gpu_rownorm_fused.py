"""Fused channel-LayerNorm + activation epilogue of the 1-D convolutions (sdt_conv_desc.rn_*, csrc/tc_conv_tma.cu: a cluster of four
CTAs per row tile exchanges the row sums through distributed shared memory) against the two-launch path it replaces
(sdt_conv_gemm + sdt_rownorm_act_fwd) and against an fp64 torch reference of the block (building_blocks.py:31-54)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,L,k,s,p,cin", [(32, 64, 3, 1, 1, 256), (4, 64, 4, 2, 1, 256), (3, 32, 4, 2, 1, 256), (2, 8, 3, 1, 1, 256),
                                           (5, 2, 3, 1, 1, 256), (128, 64, 3, 1, 1, 256), (7, 64, 3, 1, 1, 288), (2, 150, 3, 1, 1, 256)])
def test_fused_rownorm_epilogue_matches_two_launch_path(B, L, k, s, p, cin):
    from speechdrivestemplates_b200 import ops
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(B * 1000 + L)
    g = ops.ConvGeom.conv1d(cin, 256, k, s, p)
    x = torch.randn(B, L, cin, generator=gen).to(dev)
    w = (torch.randn(256, cin, k, generator=gen) / math.sqrt(cin * k)).to(dev)
    lo = g.out_hw(1, L)[1]
    wt_nk = torch.empty(256, g.k, device=dev)
    ops.weight_prep_fwd_nk(w.view(256, cin, 1, k), g, wt_nk)
    slope = 0.2
    # two launches
    raw_a = torch.empty(B, lo, 256, device=dev)
    ops.conv_gemm(ops.fwd_desc(g, x, None, raw_a, B, 1, L, wt_nk=wt_nk, math=3))
    act_a, mean_a, rstd_a = ops.rownorm_act_fwd(raw_a, slope, tf32=True)
    # one launch
    raw_b = torch.full((B, lo, 256), float("nan"), device=dev)
    act_b = torch.full((B, lo, 256), float("nan"), device=dev)
    mean_b = torch.full((B * lo,), float("nan"), device=dev)
    rstd_b = torch.full((B * lo,), float("nan"), device=dev)
    d = ops.fwd_desc(g, x, None, raw_b, B, 1, L, wt_nk=wt_nk, math=3)
    assert ops.conv_rownorm_ok(d)
    ops.conv_gemm(ops.set_rownorm(d, act_b, mean_b, rstd_b, slope, True))
    torch.cuda.synchronize()
    assert torch.equal(raw_a, raw_b)                                     # the same accumulators
    assert torch.allclose(mean_a, mean_b, rtol=0, atol=2e-6) and torch.allclose(rstd_a, rstd_b, rtol=2e-6, atol=0)
    # activations are stored rounded to TF32 (10 mantissa bits): a 1e-7 difference in the statistics may flip a rounding
    assert float((act_a - act_b).abs().max()) <= 2 ** -10 * float(act_a.abs().max())
    assert float(((act_a - act_b).abs() > 1e-6).float().mean()) < 1e-2
    # fp64 reference of the block on the TF32-rounded operands the kernel saw
    xr = x.double().cpu()
    wr = wt_nk.double().cpu().view(256, k, cin).permute(0, 2, 1)          # (Cout, Cin, k) from the K-major (tap, channel) operand
    y = torch.nn.functional.conv1d(xr.permute(0, 2, 1), wr, stride=s, padding=p).permute(0, 2, 1)
    mu = y.mean(-1, keepdim=True)
    var = y.var(-1, unbiased=False, keepdim=True)
    ref = torch.nn.functional.leaky_relu((y - mu) / torch.sqrt(var + 1e-5), slope)
    err = float((act_b.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err < 3e-3, err                                                # TF32 inputs (x is not pre-rounded here) + TF32-rounded output


def test_fused_rownorm_is_refused_where_it_does_not_apply():
    from speechdrivestemplates_b200 import _lib, ops
    dev = torch.device("cuda:0")
    g = ops.ConvGeom.conv1d(256, 128, 3, 1, 1)                            # N = 128: not four 64-column quarters
    x = torch.randn(2, 16, 256, device=dev)
    wt_nk = torch.empty(128, g.k, device=dev)
    raw = torch.empty(2, 16, 128, device=dev)
    d = ops.fwd_desc(g, x, None, raw, 2, 1, 16, wt_nk=wt_nk, math=3)
    assert not ops.conv_rownorm_ok(d)
    with pytest.raises(_lib.SdtError, match="rownorm"):
        ops.conv_gemm(ops.set_rownorm(d, raw, raw.view(-1)[:32], raw.view(-1)[:32], 0.2, True))
    g2 = ops.ConvGeom.conv1d(256, 256, 3, 1, 1)
    wt2 = torch.empty(256, g2.k, device=dev)
    raw2 = torch.empty(2, 16, 256, device=dev)
    d0 = ops.fwd_desc(g2, x, None, raw2, 2, 1, 16, wt_nk=wt2, math=0)      # fp32 FFMA mode: no tensor-core kernel, no fused epilogue
    assert not ops.conv_rownorm_ok(d0)
