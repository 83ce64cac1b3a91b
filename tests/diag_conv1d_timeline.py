"""Diagnostic (not a test): phases of tc_conv_tma_kernel on the 1-D layer shapes (B x L x 256 -> 256, k = 3 / 4).
    python tests/diag_conv1d_timeline.py [batch]
Per CTA (globaltimer, ns since the kernel's first CTA start): prologue done, first operand stage landed, last MMA issued, accumulator
complete, epilogue done, CTA end."""
import ctypes as C
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speechdrivestemplates_b200 import _lib, ops  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda:0")
    lib = _lib.load()
    lib.sdt_debug_tma_timeline.argtypes = [C.c_void_p, C.c_int]
    ops.set_conv_math(3)
    ncta = 1024
    tlbuf = torch.zeros(ncta, 8, dtype=torch.int64, device=dev)
    for (L, k, s, p) in [(64, 3, 1, 1), (64, 4, 2, 1), (2, 3, 1, 1)]:
        g = ops.ConvGeom.conv1d(256, 256, k, s, p)
        x = torch.randn(B, L, 256, device=dev)
        w = torch.randn(256, 256, k, device=dev) / math.sqrt(256 * k)
        lo = g.out_hw(1, L)[1]
        wt_nk = torch.empty(256, g.k, device=dev)
        ops.weight_prep_fwd_nk(w.view(256, 256, 1, k), g, wt_nk)
        y = torch.empty(B, lo, 256, device=dev)
        d = ops.fwd_desc(g, x, None, y, B, 1, L, wt_nk=wt_nk, math=3)
        for _ in range(3):
            ops.conv_gemm(d)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv_gemm(d)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 50
        for flags in (0, 2, 3, 4):
            tlbuf.zero_()
            lib.sdt_debug_tma_flags(flags)
            lib.sdt_debug_tma_timeline(C.c_void_p(tlbuf.data_ptr()), ncta)
            ops.conv_gemm(d)
            torch.cuda.synchronize()
            lib.sdt_debug_tma_timeline(None, 0)
            lib.sdt_debug_tma_flags(0)
            t = tlbuf.cpu().double()
            t = t[t[:, 0] > 0]
            base = t[:, 0].min()
            rel = (t[:, :7] - base) / 1e3
            names = ["start", "prologue", "1st stage", "MMAs issued", "acc done", "epilogue", "end"]
            what = ["", " [1 of 4 MMAs per k-block]", " [no MMAs]", " [no MMAs, stages released by a plain mbarrier arrive]", " [no MMAs, plain arrive, weight boxes only]"][flags]
            print("L %2d k %d s %d%s: %d CTAs, %.1f us back to back | mean us since the first CTA start: %s | span %.2f us" % (
                L, k, s, what, t.shape[0], us, "  ".join("%s %.2f" % (n, rel[:, i].mean()) for i, n in enumerate(names)), float(rel[:, 6].max())), flush=True)
    ops.set_conv_math(0)


if __name__ == "__main__":
    main()
