"""Tensor-level wrappers over the C ABI (include/sdt_b200.h).

Every function takes CUDA torch tensors (channels-last fp32 activations), launches on torch's CURRENT stream, and
fails loudly (``SdtError`` / ``ValueError``) -- there is no eager/CPU fallback.  PyTorch is used only for device
memory and streams.
"""
import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import ConvDesc, call

EPS_NORM = 1e-5


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    """Device pointer of a tensor (or NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _chk(t, dtype=torch.float32, name="tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (no CPU fallback exists)" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


# ------------------------------------------------------------------------------------------------
# mel
# ------------------------------------------------------------------------------------------------

def mel_band_tables(fb):
    """(257, 80) filterbank buffer -> band form (start, count, weights[80, stride]) on fb's device."""
    fbc = fb.detach().float().cpu()
    n_freq, n_mel = fbc.shape
    starts, counts = [], []
    for m in range(n_mel):
        nz = torch.nonzero(fbc[:, m]).flatten()
        if nz.numel() == 0:
            starts.append(0)
            counts.append(0)
        else:
            starts.append(int(nz[0]))
            counts.append(int(nz[-1]) - int(nz[0]) + 1)
    stride = max(1, max(counts))
    w = torch.zeros(n_mel, stride)
    for m in range(n_mel):
        w[m, :counts[m]] = fbc[starts[m]:starts[m] + counts[m], m]
    dev = fb.device
    return (torch.tensor(starts, dtype=torch.int32, device=dev), torch.tensor(counts, dtype=torch.int32, device=dev),
            w.to(dev).contiguous(), stride)


def mel_fwd(audio, window, tables, out=None):
    """audio (B, L) -> mel (B, 80, 1 + L//160).  voice2pose.py:125."""
    _chk(audio, name="audio")
    _chk(window, name="window")
    start, count, weight, stride = tables
    B, L = audio.shape
    T = 1 + L // 160
    if out is None:
        out = torch.empty(B, 80, T, device=audio.device, dtype=torch.float32)
    call("sdt_mel_fwd", _p(audio), B, L, _p(window), _p(start), _p(count), _p(weight), stride, _p(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# convolution geometry -> descriptors
# ------------------------------------------------------------------------------------------------

@dataclass(frozen=True)
class ConvGeom:
    """One nn.Conv1d / nn.Conv2d of the reference; 1-D convs are 2-D with H = KH = 1."""
    cin: int
    cout: int
    kh: int
    kw: int
    sh: int
    sw: int
    ph: int
    pw: int

    @staticmethod
    def conv1d(cin, cout, k, s, p):
        return ConvGeom(cin, cout, 1, k, 1, s, 0, p)

    @staticmethod
    def conv2d(cin, cout, kh, kw, s, p):
        return ConvGeom(cin, cout, kh, kw, s, s, p, p)

    def out_hw(self, h, w):
        return (h + 2 * self.ph - self.kh) // self.sh + 1, (w + 2 * self.pw - self.kw) // self.sw + 1

    @property
    def k(self):
        return self.kh * self.kw * self.cin

    def dgrad_classes(self, h, w):
        """Stride-parity decomposition of the data gradient: list of dicts (one implicit GEMM each)."""
        out = []
        for py in range(self.sh):
            ky0 = (py + self.ph) % self.sh
            th = max(0, -(-(self.kh - ky0) // self.sh))
            gh = max(0, -(-(h - py) // self.sh))
            for px in range(self.sw):
                kx0 = (px + self.pw) % self.sw
                tw = max(0, -(-(self.kw - kx0) // self.sw))
                gw = max(0, -(-(w - px) // self.sw))
                if gh == 0 or gw == 0:
                    continue
                out.append(dict(py=py, px=px, ky0=ky0, kx0=kx0, th=th, tw=tw, gh=gh, gw=gw,
                                y_off=(py + self.ph - ky0) // self.sh, x_off=(px + self.pw - kx0) // self.sw))
        return out


def set_conv_math(mode):
    """PROCESS DEFAULT of the convolution math mode: 0 = fp32 FFMA kernels, 1 = tcgen05 TF32 tensor-core kernels where a layer
    is eligible, 2 = 1 with TMA-delivered operands, 3 = 2 with the persistent kernel (shared-memory operand reuse across
    vertical taps / accumulators, image-spanning patches, parity classes in one launch), 4 = 3 with CTA pairs (experimental).

    The default starts at 3 (SDT_CONV_MATH in the environment overrides it).  It is what an engine built without an explicit
    ``conv_math`` captures at construction and what the single-shot helpers of this module use; every engine then passes ITS
    mode per descriptor (``sdt_conv_desc.math``), so changing the default never changes a live engine."""
    call("sdt_set_conv_math", int(mode))


def get_conv_math():
    return call("sdt_get_conv_math")


def resolve_math(math):
    """An explicit mode, or the process default when ``math`` is None."""
    return get_conv_math() if math is None else int(math)


def _set_math(d, math):
    d.math = 0 if math is None else int(math) + 1          # 0 = process default (include/sdt_b200.h)


def tc_eligible(cin, cout):
    """Static part of the tcgen05 path's eligibility (csrc/tc_conv.cu): K-block = 32 channels, N tile = 64/128/256."""
    return cin % 32 == 0 and cout in (64, 128, 256)


def fwd_desc(g, x, wt, dst, B, H, W, xf=None, slope=1.0, bias=None, stat_partial=None, per_image=False, accumulate=False,
             wt_nk=None, math=None):
    oh, ow = g.out_hw(H, W)
    d = ConvDesc()
    _set_math(d, math)
    d.src, d.bias, d.dst = x.data_ptr(), bias.data_ptr() if bias is not None else None, dst.data_ptr()
    d.wt = wt.data_ptr() if wt is not None else None
    d.wt_nk = wt_nk.data_ptr() if wt_nk is not None else None
    if xf is not None:
        d.xf_scale, d.xf_shift, d.xf_bstride = xf[0].data_ptr(), xf[1].data_ptr(), xf[2]
    d.xf_slope = slope
    d.stat_partial = stat_partial.data_ptr() if stat_partial is not None else None
    d.B, d.SH, d.SW, d.C = B, H, W, g.cin
    d.GH, d.GW, d.TH, d.TW = oh, ow, g.kh, g.kw
    d.y_mul, d.ty_mul, d.y_off = g.sh, 1, -g.ph
    d.x_mul, d.tx_mul, d.x_off = g.sw, 1, -g.pw
    d.N = g.cout
    d.DH, d.DW, d.dy_mul, d.dy_off, d.dx_mul, d.dx_off = oh, ow, 1, 0, 1, 0
    d.accumulate = int(accumulate)
    d.per_image_tiles = int(per_image)
    d.splits = 1
    return d


def dgrad_desc(g, cls, dy, wt_cls, dx, B, H, W, accumulate=False, wt_nk=None, math=None):
    """Data gradient for one stride-parity class: dx[b, q*s+py, ...] = sum_taps dy[...] * W."""
    oh, ow = g.out_hw(H, W)
    d = ConvDesc()
    _set_math(d, math)
    d.src, d.dst = dy.data_ptr(), dx.data_ptr()
    d.wt = wt_cls.data_ptr() if wt_cls is not None else None
    d.wt_nk = wt_nk.data_ptr() if wt_nk is not None else None
    d.xf_slope = 1.0
    d.B, d.SH, d.SW, d.C = B, oh, ow, g.cout
    d.GH, d.GW, d.TH, d.TW = cls["gh"], cls["gw"], cls["th"], cls["tw"]
    d.y_mul, d.ty_mul, d.y_off = 1, -1, cls["y_off"]
    d.x_mul, d.tx_mul, d.x_off = 1, -1, cls["x_off"]
    d.N = g.cin
    d.DH, d.DW = H, W
    d.dy_mul, d.dy_off, d.dx_mul, d.dx_off = g.sh, cls["py"], g.sw, cls["px"]
    d.accumulate = int(accumulate)
    d.per_image_tiles = 0
    d.splits = 1
    return d


def wgrad_desc(g, x, dy, wpart, B, H, W, splits, xf=None, slope=1.0, math=None):
    oh, ow = g.out_hw(H, W)
    d = ConvDesc()
    _set_math(d, math)
    d.src, d.dy, d.wpart = x.data_ptr(), dy.data_ptr(), wpart.data_ptr()
    if xf is not None:
        d.xf_scale, d.xf_shift, d.xf_bstride = xf[0].data_ptr(), xf[1].data_ptr(), xf[2]
    d.xf_slope = slope
    d.B, d.SH, d.SW, d.C = B, H, W, g.cin
    d.GH, d.GW, d.TH, d.TW = oh, ow, g.kh, g.kw
    d.y_mul, d.ty_mul, d.y_off = g.sh, 1, -g.ph
    d.x_mul, d.tx_mul, d.x_off = g.sw, 1, -g.pw
    d.N = g.cout
    d.splits = splits
    return d


def wgrad_splits(g, B, oh, ow, target_ctas=444, math=None):
    """Number of K (pixel) splits so that the weight-gradient GEMM fills the 148 SMs ~3 times over -- but never fewer than
    256 pixels (8 k-blocks of the tensor-core kernel) per split: every split writes, and the reduction re-reads, a full
    copy of the layer's gradient, which for the 1-D stacks (2048 pixels, 0.8 MB of weights) used to be 32 copies per layer."""
    if g.k <= 16 and g.cout <= 64:          # streaming small-K kernel: one partial per CTA, 4 CTAs per SM
        return min(148 * 4, max(1, -(-(B * oh * ow) // 64)))
    if resolve_math(math) >= 3 and g.cin % 128 == 0 and g.cout % 128 == 0 and g.kh >= 2 and oh >= 2:
        # csrc/tc_wgrad_ytap.cu: one CTA per SM and per (128-channel block, kernel column, group of <= 3 vertical taps, N tile)
        groups = sum(-(-(-(-(g.kh - p) // g.sh)) // 3) for p in range(min(g.sh, g.kh)))
        tiles = (g.cin // 128) * g.kw * groups * (g.cout // 128)
        return max(1, min(148 // tiles, -(-(B * oh * ow) // 256)))
    tiles = -(-g.k // 128) * -(-g.cout // (64 if g.cout <= 64 else 128))
    pixels = B * oh * ow
    return max(1, min(-(-target_ctas // tiles), -(-pixels // 256)))


def conv_gemm_multi(descs):
    """Several convolution problems in stream order (the stride-parity classes of one data gradient): one persistent launch
    in math mode 3 when they are compatible, else one launch each (csrc/conv_gemm.cu: sdt_conv_gemm_multi)."""
    if len(descs) == 1:
        return conv_gemm(descs[0])
    arr = (ConvDesc * len(descs))(*descs)
    launched = C.c_int(0)
    call("sdt_conv_gemm_multi", arr, len(descs), _stream(), C.byref(launched))
    _lib.launch_count += launched.value - 1          # call() counted one kernel


def row_tiles(desc):
    n = call("sdt_conv_row_tiles", C.byref(desc))
    if n < 0:
        raise _lib.SdtError("sdt_conv_row_tiles: " + _lib.last_error())
    return n


def conv_gemm(desc):
    call("sdt_conv_gemm", C.byref(desc), _stream())


def conv_rownorm_ok(desc):
    """Can sdt_conv_gemm run this (1-D forward) problem with the fused channel-LayerNorm + activation epilogue?"""
    return bool(call("sdt_conv_rownorm_ok", C.byref(desc)))


def set_rownorm(desc, act, mean, rstd, slope, tf32):
    """Ask sdt_conv_gemm(desc) to also write act = act((dst - mean_row) * rstd_row) and the row statistics (include/sdt_b200.h rn_*)."""
    desc.rn_act, desc.rn_mean, desc.rn_rstd = act.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    desc.rn_eps, desc.rn_slope, desc.rn_out_tf32 = EPS_NORM, float(slope), int(tf32)
    return desc


def conv_wgrad(desc):
    call("sdt_conv_wgrad", C.byref(desc), _stream())


def wgrad_reduce(wpart, splits, g, grad, accumulate=False):
    call("sdt_conv_wgrad_reduce", _p(wpart), splits, g.cout, g.cin, g.kh * g.kw, _p(grad), int(accumulate), _stream())


def wgrad_reduce_batch(table, n_items, max_ctas, max_T):
    """table: device uint8 tensor holding n_items sdt_reduce_item records (include/sdt_b200.h)."""
    call("sdt_conv_wgrad_reduce_batch", C.c_void_p(table.data_ptr()), n_items, max_ctas, max_T, _stream())


def reduce_item_table(items, device):
    """items: [(wpart, grad, splits, N, C, T, accumulate)] -> (device table, max_ctas, max_T)."""
    import numpy as np
    dt = np.dtype([("wpart", "<u8"), ("grad", "<u8"), ("i", "<i4", (6,))])
    arr = np.zeros(len(items), dtype=dt)
    max_ctas = max_T = 1
    for k, (wp, gr, splits, N, Cc, T, acc) in enumerate(items):
        assert Cc % 32 == 0 and T <= 256
        arr[k] = (wp.data_ptr(), gr.data_ptr(), (splits, N, Cc, T, int(acc), 0))
        max_ctas, max_T = max(max_ctas, N * (Cc // 32)), max(max_T, T)
    return torch.from_numpy(arr.view(np.uint8).copy()).to(device), max_ctas, max_T


def weight_prep_fwd(w, g, out):
    """(Cout,Cin,kh,kw) reference-layout parameter -> (K, Cout) forward GEMM operand."""
    call("sdt_weight_prep", _p(w), g.cout, g.cin, g.kh, g.kw, 0, 0, 0, 1, g.kh, g.kw, _p(out), _stream())


def weight_prep_dgrad(w, g, cls, out):
    """... -> (th*tw*Cout, Cin) data-gradient operand of one parity class."""
    call("sdt_weight_prep", _p(w), g.cout, g.cin, g.kh, g.kw, 1, cls["ky0"], cls["kx0"], g.sw,
         cls["th"], cls["tw"], _p(out), _stream())


def weight_prep_fwd_nk(w, g, out):
    """... -> (Cout, K) K-major tensor-core operand of the forward GEMM."""
    call("sdt_weight_prep", _p(w), g.cout, g.cin, g.kh, g.kw, 2, 0, 0, 1, g.kh, g.kw, _p(out), _stream())


def weight_prep_dgrad_nk(w, g, cls, out):
    """... -> (Cin, th*tw*Cout) K-major tensor-core operand of one data-gradient parity class."""
    call("sdt_weight_prep", _p(w), g.cout, g.cin, g.kh, g.kw, 3, cls["ky0"], cls["kx0"], g.sw,
         cls["th"], cls["tw"], _p(out), _stream())


# ------------------------------------------------------------------------------------------------
# convenience single-shot convolution ops (allocate their own scratch; used by tests and module boundaries)
# ------------------------------------------------------------------------------------------------

def _as_w4(w):
    return w if w.dim() == 4 else w.unsqueeze(2)


def conv_forward(x, w, g, xf=None, slope=1.0, bias=None, want_stats=False, per_image=False):
    """x (B,H,W,Cin) channels-last, w reference layout -> y (B,OH,OW,Cout) [, partial (tiles,2,Cout)]."""
    _chk(x, name="x")
    _chk(w, name="w")
    B, H, W, _ = x.shape
    oh, ow = g.out_hw(H, W)
    wt = torch.empty(g.k, g.cout, device=x.device)
    weight_prep_fwd(w, g, wt)
    wt_nk = None
    if get_conv_math() >= 1 and tc_eligible(g.cin, g.cout):
        wt_nk = torch.empty(g.cout, g.k, device=x.device)
        weight_prep_fwd_nk(w, g, wt_nk)
    y = torch.empty(B, oh, ow, g.cout, device=x.device)
    d = fwd_desc(g, x, wt, y, B, H, W, xf, slope, bias, None, per_image, wt_nk=wt_nk)
    partial = None
    if want_stats:
        partial = torch.empty(row_tiles(d), 2, g.cout, device=x.device)
        d.stat_partial = partial.data_ptr()
    conv_gemm(d)
    return (y, partial) if want_stats else y


def conv_dgrad(dy, w, g, H, W, out=None, accumulate=False):
    """dy (B,OH,OW,Cout) -> dx (B,H,W,Cin)."""
    _chk(dy, name="dy")
    B = dy.shape[0]
    dx = out if out is not None else torch.empty(B, H, W, g.cin, device=dy.device)
    # every input position belongs to exactly one stride-parity class, so the classes together write all of dx
    descs, keep = [], []          # keep: the operand tensors must outlive the (asynchronous) launch
    for cls in g.dgrad_classes(H, W):
        wt = torch.empty(cls["th"] * cls["tw"] * g.cout, g.cin, device=dy.device)
        weight_prep_dgrad(w, g, cls, wt)
        wt_nk = None
        if get_conv_math() >= 1 and tc_eligible(g.cout, g.cin):
            wt_nk = torch.empty(g.cin, cls["th"] * cls["tw"] * g.cout, device=dy.device)
            weight_prep_dgrad_nk(w, g, cls, wt_nk)
        descs.append(dgrad_desc(g, cls, dy, wt, dx, B, H, W, accumulate, wt_nk=wt_nk))
        keep.append((wt, wt_nk))
    conv_gemm_multi(descs)
    return dx


def conv_weight_grad(x, dy, g, xf=None, slope=1.0, splits=None):
    """-> gradient in the reference parameter layout (Cout, Cin, kh, kw)."""
    _chk(x, name="x")
    _chk(dy, name="dy")
    B, H, W, _ = x.shape
    oh, ow = g.out_hw(H, W)
    splits = splits or wgrad_splits(g, B, oh, ow)
    wpart = torch.empty(splits, g.cout, g.k, device=x.device)
    conv_wgrad(wgrad_desc(g, x, dy, wpart, B, H, W, splits, xf, slope))
    grad = torch.empty(g.cout, g.cin, g.kh, g.kw, device=x.device)
    wgrad_reduce(wpart, splits, g, grad)
    return grad


# ------------------------------------------------------------------------------------------------
# normalisation
# ------------------------------------------------------------------------------------------------

def norm_finalize(partial, groups, C, count, gamma=None, beta=None, running=None, momentum=0.1, out=None):
    """-> (scale, shift, mean, rstd), each (groups, C). running = (running_mean, running_var, num_batches_tracked)."""
    tiles_per_group = partial.shape[0] // groups
    dev = partial.device
    scale, shift, mean, rstd = out if out is not None else [torch.empty(groups, C, device=dev) for _ in range(4)]
    rm, rv, nbt = running if running is not None else (None, None, None)
    call("sdt_norm_finalize", _p(partial), groups, tiles_per_group, C, float(count), _p(gamma), _p(beta), EPS_NORM,
         _p(scale), _p(shift), _p(mean), _p(rstd), _p(rm), _p(rv), _p(nbt), momentum, _stream())
    return scale, shift, mean, rstd


def chan_stats(x, w0, w1, rows_per_part, partial):
    """x (B, H, W, C) channels-last; partial (B, n_parts, 2, C) <- sums / sums of squares over the columns [w0, w1) (include/sdt_b200.h)."""
    B, H, W, Cc = x.shape
    n_parts = partial.shape[1]
    call("sdt_chan_stats", _p(x), B, H, W, Cc, int(w0), int(w1), int(rows_per_part), _p(partial), n_parts, _stream())
    return partial


def bn_eval_scale_shift(rm, rv, gamma, beta, out=None):
    Cc = rm.numel()
    scale, shift = out if out is not None else (torch.empty(1, Cc, device=rm.device), torch.empty(1, Cc, device=rm.device))
    call("sdt_bn_eval_scale_shift", _p(rm), _p(rv), _p(gamma), _p(beta), EPS_NORM, Cc, _p(scale), _p(shift), _stream())
    return scale, shift


BWD_ROWS = 256


def bwd_tiles(P, B):
    """Tiles per image of the norm-backward reduction: at most 256 rows per tile, and enough tiles (down to 32 rows each)
    for ~600 CTAs -- with one 256-row tile per image the small maps ran 32 CTAs of serial loads (41 us for 17 MB)."""
    return max(-(-P // BWD_ROWS), min(-(-P // 32), -(-600 // max(B, 1))))


# 0 = off (default).  Measured at B = 32 (profiles/r2_ablation_classminor_normbwd_groups.txt): image groups of 40 / 64 / 100 MB make the
# step 0.08 / 0.04 / 0.02 ms SLOWER than the whole-batch passes -- the extra launches and the smaller grids cost more than the L2 hits
# of the apply pass save (the whole-batch passes already run at 5.4 / 6.4 TB/s).
NORM_BWD_L2_BYTES = int(float(__import__("os").environ.get("SDT_NORM_BWD_L2_MB", "0")) * 1e6)


def bwd_partial_tiles(P, B, Cc):
    """Tiles per image to ALLOCATE for norm_backward's partial sums: the image-group path uses more tiles per image than the
    whole-batch path (fewer images per launch, same ~600 CTAs)."""
    per_image = 2 * P * Cc * 4
    nb = max(1, min(B, NORM_BWD_L2_BYTES // per_image)) if NORM_BWD_L2_BYTES > 0 else B
    return max(bwd_tiles(P, B), min(bwd_tiles(P, nb), P))




def norm_backward(g, x, mean, rstd, groups, slope, gamma=None, beta=None, dgamma=None, dbeta=None, accumulate=False,
                  scratch=None, tf32=False):
    """In place: g (B,P,C) := d loss / d x for [normalise(groups) -> affine -> act]; returns g.
    tf32: store the result rounded to TF32 (it is the next data / weight gradient's tensor-core operand, sdt_b200.h "out_tf32").

    Two passes over g and x (reduce, apply).  With per-image statistics (groups == B) the images are independent, so maps larger
    than L2 are processed in image groups of <= NORM_BWD_L2_BYTES (g + x): the apply pass of a group then reads what its reduce
    pass just brought into the 126 MB L2 instead of streaming 2 x 140 MB from HBM a second time."""
    B = g.shape[0]
    Cc = g.shape[-1]
    P = g.numel() // (B * Cc)
    tpi = bwd_tiles(P, B)
    if scratch is None:
        partial = torch.empty(B * tpi, 2, Cc, device=g.device)
        m1 = torch.empty(groups, Cc, device=g.device)
        m2 = torch.empty(groups, Cc, device=g.device)
    else:
        partial, m1, m2 = scratch
    per_image = 2 * P * Cc * 4
    if groups == B and B > 1 and NORM_BWD_L2_BYTES > 0 and B * per_image > NORM_BWD_L2_BYTES and gamma is None:
        nb = max(1, min(B, NORM_BWD_L2_BYTES // per_image))
        nb = -(-B // -(-B // nb))                      # balanced groups
        st = _stream()
        for b0 in range(0, B, nb):
            n = min(nb, B - b0)
            tg = max(1, min(bwd_tiles(P, nb), partial.shape[0] // B, P))
            gs, xs = g[b0:b0 + n], x[b0:b0 + n]
            ps = partial[b0 * tg:(b0 + n) * tg]
            call("sdt_norm_bwd_reduce", _p(gs), _p(xs), _p(mean[b0:b0 + n]), _p(rstd[b0:b0 + n]), None, None, n, P, Cc, n, slope, _p(ps), tg, st)
            call("sdt_norm_bwd_finalize", _p(ps), n, tg, Cc, float(P), _p(m1[b0:b0 + n]), _p(m2[b0:b0 + n]), None, None, 0, st)
            call("sdt_norm_bwd_apply", _p(gs), _p(xs), _p(mean[b0:b0 + n]), _p(rstd[b0:b0 + n]), None, None, _p(m1[b0:b0 + n]), _p(m2[b0:b0 + n]),
                 n, P, Cc, n, slope, int(tf32), st)
        return g
    call("sdt_norm_bwd_reduce", _p(g), _p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), B, P, Cc, groups, slope, _p(partial),
         tpi, _stream())
    call("sdt_norm_bwd_finalize", _p(partial), groups, (B * tpi) // groups, Cc, float(P * (B // groups)), _p(m1), _p(m2),
         _p(dgamma), _p(dbeta), int(accumulate), _stream())
    call("sdt_norm_bwd_apply", _p(g), _p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(m1), _p(m2), B, P, Cc, groups, slope,
         int(tf32), _stream())
    return g


def rownorm_act_fwd(x, slope, out=None, tf32=False):
    """x (..., C) -> (y, mean (R), rstd (R)): channel LayerNorm (the reference's 1-D 'IN') + activation."""
    Cc = x.shape[-1]
    R = x.numel() // Cc
    if out is None:
        y = torch.empty_like(x)
        mean = torch.empty(R, device=x.device)
        rstd = torch.empty(R, device=x.device)
    else:
        y, mean, rstd = out
    call("sdt_rownorm_act_fwd", _p(x), R, Cc, EPS_NORM, slope, _p(y), _p(mean), _p(rstd), int(tf32), _stream())
    return y, mean, rstd


def rownorm_act_bwd(g_y, x, mean, rstd, slope, out=None, tf32=False):
    Cc = x.shape[-1]
    R = x.numel() // Cc
    g_x = out if out is not None else torch.empty_like(x)
    call("sdt_rownorm_act_bwd", _p(g_y), _p(x), _p(mean), _p(rstd), R, Cc, slope, _p(g_x), int(tf32), _stream())
    return g_x


def scale_shift_act(x, scale, shift, bstride, slope, out=None, tf32=False):
    B, Cc = x.shape[0], x.shape[-1]
    P = x.numel() // (B * Cc)
    y = out if out is not None else torch.empty_like(x)
    call("sdt_scale_shift_act", _p(x), _p(scale), _p(shift), B, P, Cc, bstride, slope, _p(y), int(tf32), _stream())
    return y


def first_layer_units(H, W):
    return call("sdt_first_layer_units", H, W)


def first_layer_fwd(x, w, slope, eps=None, out=None, scratch=None, tf32=False):
    """x (B,H,W) one-channel image, w (64,1,3,3) -> (act (B,H,W,64), scale (B,64), shift (B,64), moments (B,54) f64):
    Conv2d(1,64,3,1,1) + InstanceNorm2d + LeakyReLU in one pass over the output (generator.py:17)."""
    B, H, W = x.shape
    Cc = w.shape[0]
    units = first_layer_units(H, W)
    if out is None:
        out = (torch.empty(B, H, W, Cc, device=x.device), torch.empty(B, Cc, device=x.device),
               torch.empty(B, Cc, device=x.device), torch.empty(B, 54, device=x.device, dtype=torch.float64))
    act, sc, sh, mom = out
    part = scratch if scratch is not None else torch.empty(B, units, 54, device=x.device, dtype=torch.float64)
    call("sdt_first_layer_fwd", _p(x), _p(w), B, H, W, Cc, EPS_NORM if eps is None else eps, slope, _p(part), _p(mom), _p(sc), _p(sh), _p(act), int(tf32), _stream())
    return act, sc, sh, mom


def first_layer_stats(x, w, eps=None, out=None, scratch=None):
    """Statistics half of first_layer_fwd: x (B,H,W), w (64,1,3,3) -> (scale (B,64) = rstd, shift (B,64) = -mean*rstd, moments (B,54) f64)."""
    B, H, W = x.shape
    Cc = w.shape[0]
    units = first_layer_units(H, W)
    sc, sh, mom = out if out is not None else (torch.empty(B, Cc, device=x.device), torch.empty(B, Cc, device=x.device),
                                               torch.empty(B, 54, device=x.device, dtype=torch.float64))
    part = scratch if scratch is not None else torch.empty(B, units, 54, device=x.device, dtype=torch.float64)
    call("sdt_first_layer_fwd", _p(x), _p(w), B, H, W, Cc, EPS_NORM if eps is None else eps, 1.0, _p(part), _p(mom), _p(sc), _p(sh), None, 0, _stream())
    _lib.launch_count -= 1                      # call() counted three kernels; the activation pass is skipped
    return sc, sh, mom


def first_layer_act(x, w, scale, shift, slope, out=None, tf32=False):
    """Activation half: x (B,H,W) (a tile of the image), per-(image, channel) scale / shift -> act (B,H,W,64)."""
    B, H, W = x.shape
    Cc = w.shape[0]
    act = out if out is not None else torch.empty(B, H, W, Cc, device=x.device)
    call("sdt_first_layer_act", _p(x), _p(w), _p(scale), _p(shift), B, H, W, Cc, slope, _p(act), int(tf32), _stream())
    return act


def first_layer_bwd(g_act, act, x, w, mom, sc, sh, slope, dw, scratch=None):
    """Weight gradient of the first block from dLoss/d act (IN + LeakyReLU backward folded in); dw (64,1,3,3) overwritten."""
    B, H, W = x.shape
    Cc = w.shape[0]
    units = first_layer_units(H, W)
    part = scratch if scratch is not None else torch.empty(B, units, 11, Cc, device=x.device)
    call("sdt_first_layer_bwd", _p(g_act), _p(act), _p(x), _p(w), _p(mom), _p(sc), _p(sh), B, H, W, Cc, slope, _p(part), _p(dw),
         _stream())
    return dw


# ------------------------------------------------------------------------------------------------
# resampling / losses / heads / optimizer
# ------------------------------------------------------------------------------------------------

def enc_to_seq_fwd(x, scale, shift, bstride, slope, code, F, out=None, tf32=False):
    B, H, W, Cc = x.shape
    D = 0 if code is None else code.shape[1]
    y = out if out is not None else torch.empty(B, F, Cc + D, device=x.device)
    call("sdt_enc_to_seq_fwd", _p(x), _p(scale), _p(shift), bstride, slope, B, H, W, Cc, _p(code), D, F, _p(y), int(tf32), _stream())
    return y


def enc_to_seq_bwd(g_out, H, W, Cc, D, g_act=None, g_code=None):
    B, F, _ = g_out.shape
    if g_act is None:
        g_act = torch.empty(B, H, W, Cc, device=g_out.device)
    if D > 0 and g_code is None:
        g_code = torch.empty(B, D, device=g_out.device)
    call("sdt_enc_to_seq_bwd", _p(g_out), B, H, W, Cc, D, F, _p(g_act), _p(g_code), _stream())
    return g_act, g_code


def upsample_add_fwd(x, skip, Lout, out=None, tf32=False):
    B, Lin, Cc = x.shape
    y = out if out is not None else torch.empty(B, Lout, Cc, device=x.device)
    call("sdt_upsample_add_fwd", _p(x), _p(skip), B, Lin, Lout, Cc, _p(y), int(tf32), _stream())
    return y


def upsample_bwd(g_out, Lin, out=None, accumulate=False):
    B, Lout, Cc = g_out.shape
    g_x = out if out is not None else torch.empty(B, Lin, Cc, device=g_out.device)
    call("sdt_upsample_bwd", _p(g_out), B, Lin, Lout, Cc, _p(g_x), int(accumulate), _stream())
    return g_x


def l1_loss(pred, gt, lam, loss_out, g_pred, partial):
    call("sdt_l1_loss", _p(pred), _p(gt), pred.numel(), lam, _p(loss_out), _p(g_pred), _p(partial), _stream())


def code_gather_kl(table, idx, lam, code, out, g_code):
    B, D = code.shape
    call("sdt_code_gather_kl", _p(table), _p(idx), B, D, lam, _p(code), _p(out), _p(g_code), _stream())


def code_scatter_grad(ga, gb, idx, g_table):
    B = idx.numel()
    D = g_table.shape[1]
    call("sdt_code_scatter_grad", _p(ga), _p(gb), _p(idx), B, D, _p(g_table), _stream())


def code_store_rows(src_a, table_a, idx, src_b=None, table_b=None):
    """table_a[idx[b]] = src_a[b] (and table_b / src_b): pose2pose.py:135-137, last duplicate wins."""
    B, D = src_a.shape
    call("sdt_code_store_rows", _p(src_a), _p(table_a), _p(src_b), _p(table_b), _p(idx), B, D, _stream())


def colsum(g, out, accumulate=False):
    Cc = g.shape[-1]
    call("sdt_colsum", _p(g), g.numel() // Cc, Cc, _p(out), int(accumulate), _stream())


def mse_const_loss(s, target, lam, out, g_s=None):
    call("sdt_mse_const_loss", _p(s), s.numel(), target, lam, _p(out), _p(g_s), _stream())


def motion_diff_fwd(x, out=None):
    B, T, Cc = x.shape
    y = out if out is not None else torch.empty(B, T - 1, Cc, device=x.device)
    call("sdt_motion_diff_fwd", _p(x), B, T, Cc, _p(y), _stream())
    return y


def motion_diff_bwd(g_out, out=None, accumulate=False):
    B, Tm1, Cc = g_out.shape
    g_x = out if out is not None else torch.empty(B, Tm1 + 1, Cc, device=g_out.device)
    call("sdt_motion_diff_bwd", _p(g_out), B, Tm1 + 1, Cc, _p(g_x), int(accumulate), _stream())
    return g_x


def pose_head_fwd(x, scale, shift, slope, mu=None, logvar=None):
    B, L, D2 = x.shape
    if mu is None:
        mu = torch.empty(B, D2 // 2, device=x.device)
        logvar = torch.empty(B, D2 // 2, device=x.device)
    call("sdt_pose_head_fwd", _p(x), _p(scale), _p(shift), slope, B, L, D2, _p(mu), _p(logvar), _stream())
    return mu, logvar


def vae_reparam_kl(mu, logvar, eps, lam, code, out):
    call("sdt_vae_reparam_kl", _p(mu), _p(logvar), _p(eps), mu.numel(), lam, _p(code), _p(out), _stream())


def vae_reparam_kl_bwd(mu, logvar, eps, g_code, lam, g_mu, g_logvar):
    call("sdt_vae_reparam_kl_bwd", _p(mu), _p(logvar), _p(eps), _p(g_code), mu.numel(), lam, _p(g_mu), _p(g_logvar), _stream())


def pose_head_bwd(g_mu, g_logvar, L, g_act):
    B, D = g_mu.shape
    call("sdt_pose_head_bwd", _p(g_mu), _p(g_logvar), B, L, 2 * D, _p(g_act), _stream())


def pose_preprocess(raw, mean, std, hierarchical=True, out=None):
    """raw (T,3,137) f32 -> normalised (T,2,121) f32; bit-exact with gesture_dataset.py:95-105.  T may be B*frames."""
    _chk(raw, name="raw")
    T = raw.shape[0]
    if out is None:
        out = torch.empty(T, 2, 121, device=raw.device)
    call("sdt_pose_preprocess", _p(raw), T, _p(_chk(mean, name="mean")), _p(_chk(std, name="std")), int(hierarchical), _p(out),
         _stream())
    return out


def pose_final_results(poses, mean, std, scale, hierarchical=True, out=None):
    """poses (B,T,2,121) f32; mean/std (B,242) f64; scale (B) f64 -> f64; bit-exact with gesture_dataset.py:213-220."""
    _chk(poses, name="poses")
    _chk(mean, torch.float64, "mean")
    _chk(std, torch.float64, "std")
    _chk(scale, torch.float64, "scale")
    B, T = poses.shape[0], poses.shape[1]
    if out is None:
        out = torch.empty(B, T, 2, 121, device=poses.device, dtype=torch.float64)
    call("sdt_pose_final_results", _p(poses), B, T, _p(mean), _p(std), _p(scale), int(hierarchical), _p(out), _stream())
    return out


def pose_parted2global(poses, mean_p, std_p, mean_g, std_g, out=None):
    """(…,2,121) f32 normalised parted poses -> normalised global poses (gesture_dataset.py:221-234)."""
    _chk(poses, name="poses")
    if out is None:
        out = torch.empty_like(poses)
    call("sdt_pose_parted2global", _p(poses), poses.numel() // 242, _p(mean_p), _p(std_p), _p(mean_g), _p(std_g), _p(out), _stream())
    return out


def pose_metrics(pred, gt, partial=None, out=None):
    """final-result poses (B,T,2,121) f64 -> tensor [L2_dist, lip_sync_error_n] (voice2pose.py:412-430)."""
    B, T = pred.shape[0], pred.shape[1]
    if partial is None:
        partial = torch.empty(2 * B, device=pred.device, dtype=torch.float64)
    if out is None:
        out = torch.empty(2, device=pred.device, dtype=torch.float64)
    call("sdt_pose_metrics", _p(pred), _p(gt), B, T, _p(partial), _p(out), _stream())
    return out


def adam_advance(scalars, lr, beta1=0.9, beta2=0.999):
    call("sdt_adam_advance", _p(scalars), lr, beta1, beta2, _stream())


def adam_flat(param, grad, exp_avg, exp_avg_sq, scalars, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0, weight_decay=0.0):
    call("sdt_adam_flat", _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), _p(scalars), beta1, beta2, eps,
         grad_scale, float(weight_decay), _stream())
