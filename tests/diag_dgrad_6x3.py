"""Diagnostic (not a test): the data gradient of the 6x3 / no-padding encoder layer (10x53 grid over a 5x51 source) under forced
tile geometries of tc_conv_ytap_kernel, checked against the fp32 FFMA kernel.   python tests/diag_dgrad_6x3.py [batch]"""
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
    g = ops.ConvGeom.conv2d(256, 256, 6, 3, 1, 0)
    H, W = 10, 53
    oh, ow = g.out_hw(H, W)
    torch.manual_seed(0)
    dy = torch.randn(B, oh, ow, 256, device=dev)
    w = torch.randn(256, 256, 6, 3, device=dev) / math.sqrt(256 * 18)
    ops.set_conv_math(0)
    ref = ops.conv_dgrad(dy, w, g, H, W).clone()
    ops.set_conv_math(3)
    cls = g.dgrad_classes(H, W)[0]
    wt_nk = torch.empty(g.cin, cls["th"] * cls["tw"] * g.cout, device=dev)
    ops.weight_prep_dgrad_nk(w, g, cls, wt_nk)
    dx = torch.empty(B, H, W, 256, device=dev)
    d = ops.dgrad_desc(g, cls, dy, None, dx, B, H, W, False, wt_nk=wt_nk, math=3)
    flops = 2.0 * B * oh * ow * 256 * 256 * 18
    for force in [(0, 0, 0, 0), (128, 2, 8, 1), (128, 2, 16, 1), (128, 2, 32, 1), (128, 2, 8, 2), (128, 2, 16, 2), (128, 2, 8, 4), (128, 1, 64, 1),
                  (128, 1, 8, 8), (128, 1, 32, 2), (64, 2, 32, 1), (64, 4, 16, 1), (64, 4, 8, 1)]:
        lib.sdt_debug_conv_force(*force)
        plan = (C.c_int32 * 10)()
        lib.sdt_conv_plan(C.byref(d), plan)
        dx.zero_()
        for _ in range(3):
            ops.conv_gemm(d)
        torch.cuda.synchronize()
        err = float((dx - ref).abs().max() / ref.abs().max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv_gemm(d)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("force %-18s plan %-52s %.1f us  %.0f TFLOP/s (algorithmic)  err %.1e" % (force, list(plan), ms * 1e3, flops / ms / 1e9, err), flush=True)
    lib.sdt_debug_conv_force(0, 0, 0, 0)
    ops.set_conv_math(0)


if __name__ == "__main__":
    main()
