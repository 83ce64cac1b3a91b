"""Diagnostic: voice2pose_s2g eval forward (BatchNorm from running statistics) in math mode 3 vs mode 0, per-layer deviation."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_gpu_eval import _eval_model
from util import oliver_stat
from oracle import sdt_oracle as O

g, model, n_train, bs = _eval_model("s2g", "voice2pose_s2g", [])
b = O.synthetic_batch(bs, n_train, oliver_stat(False), seed=410, stat_parted=oliver_stat(True), stat_global=oliver_stat(False))
hb = dict(b)
hb["speaker_stat"] = {k: torch.from_numpy(np.asarray(v)) for k, v in b["speaker_stat"].items()}
bufs = {}
for mode in (0, 3):
    model.set_conv_math(mode)
    with torch.no_grad():
        losses, results = model(hb, None)
    eng = model.netG.engine()
    bufs[mode] = {k: v.clone() for k, v in eng.arena.bufs.items() if k.startswith("raw:") or k in ("x0", "pred")}
    print(mode, {k: float(v) for k, v in losses.items()})
for k in bufs[0]:
    if k in bufs[3] and bufs[0][k].shape == bufs[3][k].shape:
        a, r = bufs[3][k].double(), bufs[0][k].double()
        print("%-60s max|ref| %.3e rms %.3e  max err/max %.2e  rel-L2 %.2e" % (k, r.abs().max(), r.pow(2).mean().sqrt(), (a - r).abs().max() / r.abs().max(), (a - r).norm() / r.norm()))
