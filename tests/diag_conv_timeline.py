"""Diagnostic (not a test): per-CTA timeline of the mode-3 convolution kernel on the encoder layer shapes.
    python tests/diag_conv_timeline.py [batch]
Per persistent CTA: life = CTA start -> end; clocks the MMA thread spent blocked on operands (wait_full) and on the epilogue
handing an accumulator set back (wait_acc), the TMA thread on ring slots (wait_empty), an epilogue warp on accumulators (wait_tmem)."""
import ctypes as C
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speechdrivestemplates_b200 import _lib, ops  # noqa: E402

if os.environ.get("SDT_DIAG_LIB"):          # A/B against another build of the library (diagnostic only)
    _lib.LIB_PATH = os.environ["SDT_DIAG_LIB"]

LAYERS = [(64, 64, 4, 4, 2, 1, 80, 427), (64, 128, 3, 3, 1, 1, 40, 213), (128, 128, 4, 4, 2, 1, 40, 213),
          (128, 256, 3, 3, 1, 1, 20, 106), (256, 256, 4, 4, 2, 1, 20, 106), (256, 256, 3, 3, 1, 1, 10, 53),
          (256, 256, 6, 3, 1, 0, 10, 53)]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 3]
    flags_list = [int(f) for f in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
    dev = torch.device("cuda:0")
    lib = _lib.load()
    lib.sdt_debug_conv_timeline.argtypes = [C.c_void_p, C.c_int]
    ncta = 4096
    tlbuf = torch.zeros(ncta, 8, dtype=torch.int64, device=dev)
    for (cin, cout, kh, kw, s, p, H, W) in LAYERS:
        g = ops.ConvGeom.conv2d(cin, cout, kh, kw, s, p)
        x = torch.randn(B, H, W, cin, device=dev)
        w = torch.randn(cout, cin, kh, kw, device=dev) / math.sqrt(cin * kh * kw)
        oh, ow = g.out_hw(H, W)
        flops = 2.0 * B * oh * ow * cout * cin * kh * kw
        for mode in modes:
            ops.set_conv_math(mode)
            wt = torch.empty(g.k, g.cout, device=dev)
            ops.weight_prep_fwd(w, g, wt)
            wt_nk = torch.empty(g.cout, g.k, device=dev)
            ops.weight_prep_fwd_nk(w, g, wt_nk)
            y = torch.empty(B, oh, ow, cout, device=dev)
            d = ops.fwd_desc(g, x, wt, y, B, H, W, None, 1.0, None, None, True, wt_nk=wt_nk)
            partial = torch.empty(ops.row_tiles(d), 2, cout, device=dev)
            d.stat_partial = partial.data_ptr()
            plan = (C.c_int32 * 10)()
            lib.sdt_conv_plan(C.byref(d), plan)
            for _ in range(3):
                ops.conv_gemm(d)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.conv_gemm(d)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            line = "L %3d->%3d %dx%d s%d %3dx%3d mode %d: %.1f us  %.0f TFLOP/s  plan %s" % (
                cin, cout, kh, kw, s, H, W, mode, ms * 1e3, flops / ms / 1e9, list(plan))
            d.stat_partial = None
            for _ in range(3):
                ops.conv_gemm(d)
            e0.record()
            for _ in range(10):
                ops.conv_gemm(d)
            e1.record()
            torch.cuda.synchronize()
            line += " | no stats %.1f us" % (e0.elapsed_time(e1) * 100)
            dy = torch.randn(B, oh, ow, cout, device=dev)
            dx = torch.empty(B, H, W, cin, device=dev)
            for _ in range(3):
                ops.conv_dgrad(dy, w, g, H, W, out=dx)
            e0.record()
            for _ in range(5):
                ops.conv_dgrad(dy, w, g, H, W, out=dx)
            e1.record()
            torch.cuda.synchronize()
            line += " | dgrad (incl. weight prep) %.1f us" % (e0.elapsed_time(e1) * 200)
            d.stat_partial = partial.data_ptr()
            for fl in (flags_list if (mode == 3 and plan[0] == 3 and len(sys.argv) > 3) else []):
                lib.sdt_debug_conv_flags(fl)
                for _ in range(2):
                    ops.conv_gemm(d)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    ops.conv_gemm(d)
                e1.record()
                torch.cuda.synchronize()
                line += "\n      flags %d: %.1f us" % (fl, e0.elapsed_time(e1) * 100)
                tlbuf.zero_()
                lib.sdt_debug_conv_timeline(C.c_void_p(tlbuf.data_ptr()), ncta)
                ops.conv_gemm(d)
                torch.cuda.synchronize()
                lib.sdt_debug_conv_timeline(None, 0)
                n = min(plan[9], 148)
                t = tlbuf[:n].cpu().double()
                life = t[:, 5] - t[:, 1]
                span = float(t[:, 5].max() - t[:, 1].min())
                t0 = t[:, 1:2]
                rel = (t[:, 1:6] - t0) / 1e3
                line += "\n      ctas %d span %.1f us | mean us since CTA start: prologue done %.2f, first tile issued %.2f, all tiles issued %.2f, end %.2f (max %.2f) | clk: mma wait_full %.0f | tma wait_empty %.0f | SM clock %.0f MHz" % (
                    n, span / 1e3, rel[:, 1].mean(), rel[:, 2].mean(), rel[:, 3].mean(), rel[:, 4].mean(), rel[:, 4].max(), t[:, 7].mean(), t[:, 6].mean(), float((t[:, 0] / (t[:, 5] - t[:, 1])).mean() * 1e3))
            lib.sdt_debug_conv_flags(0)
            print(line, flush=True)
    ops.set_conv_math(0)


if __name__ == "__main__":
    main()
