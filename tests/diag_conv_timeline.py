"""Diagnostic (not a test): per-CTA timeline of the mode-3 convolution kernel on the encoder layer shapes.
    python tests/diag_conv_timeline.py [batch]
Columns (ns, mean over CTAs): life = CTA start -> end, setup = start -> MMA thread ready, first = ready -> first operands landed,
main = first operands -> accumulators complete, epi = accumulators complete -> CTA end; wait_full / wait_empty = clocks the MMA
thread / the TMA thread spent blocked on the ring barriers."""
import ctypes as C
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speechdrivestemplates_b200 import _lib, ops  # noqa: E402

LAYERS = [(64, 64, 4, 4, 2, 1, 80, 427), (64, 128, 3, 3, 1, 1, 40, 213), (128, 128, 4, 4, 2, 1, 40, 213),
          (128, 256, 3, 3, 1, 1, 20, 106), (256, 256, 4, 4, 2, 1, 20, 106), (256, 256, 3, 3, 1, 1, 10, 53),
          (256, 256, 6, 3, 1, 0, 10, 53)]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 3]
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
            if mode == 3 and plan[0] == 3:
                tlbuf.zero_()
                lib.sdt_debug_conv_timeline(C.c_void_p(tlbuf.data_ptr()), ncta)
                ops.conv_gemm(d)
                torch.cuda.synchronize()
                lib.sdt_debug_conv_timeline(None, 0)
                n = min(plan[9], ncta)
                t = tlbuf[:n].cpu().double()
                life, setup, first = t[:, 5] - t[:, 1], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2]
                mainl, epi = t[:, 4] - t[:, 3], t[:, 5] - t[:, 4]
                span = float(t[:, 5].max() - t[:, 1].min())
                line += "\n      ctas %d span %.1f us | ns: life %.0f setup %.0f first %.0f main %.0f epi %.0f | clk: wait_full %.0f wait_empty %.0f" % (
                    n, span / 1e3, life.mean(), setup.mean(), first.mean(), mainl.mean(), epi.mean(), t[:, 7].mean(), t[:, 6].mean())
                # concurrency on SM of CTA 0
                sm0 = t[0, 0]
                on = t[t[:, 0] == sm0]
                line += "\n      SM %d ran %d CTAs; their (start,end) us: %s" % (
                    int(sm0), on.shape[0], " ".join("(%.1f,%.1f)" % ((a - t[:, 1].min()) / 1e3, (b - t[:, 1].min()) / 1e3)
                                                      for a, b in sorted(zip(on[:, 1].tolist(), on[:, 5].tolist()))[:10]))
            print(line, flush=True)
    ops.set_conv_math(0)


if __name__ == "__main__":
    main()
