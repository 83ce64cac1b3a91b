"""Diagnostic: accuracy of the TF32 tensor-core math mode on the sdt_bp train step vs the reference fixture (B=2)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import golden, oliver_stat, samples_of, rel_err
from oracle import sdt_oracle as O
from speechdrivestemplates_b200 import pipeline, config, ops, _lib
from test_gpu_step import _to_host_batch

g = golden("sdt_bp_step_golden")
for mode in (0, 1):
    tr = pipeline.Voice2PoseTrainer(config.get_cfg("voice2pose_sdt_bp"), 16, "cuda:0", use_cuda_graph=False, seed=0, conv_math=mode)
    tr.model.clips_code.data.copy_(0.1 * torch.randn(16, 32, generator=torch.Generator().manual_seed(11)))
    n0 = _lib.call("sdt_tc_launches")
    out = tr.train_step(_to_host_batch(O.synthetic_batch(2, 16, oliver_stat(True), seed=100)))
    host = tr.losses_to_host(out)
    print("mode", mode, "tc launches", _lib.call("sdt_tc_launches") - n0, "kernels/step", tr.kernels_per_step)
    for k, v in host.items():
        print("   %-22s %.7f  ref %.7f  diff %.2e" % (k, v, float(g["step0/loss/" + k]), abs(v - float(g["step0/loss/" + k]))))
    print("   pred rel err %.2e   mu_pred %.2e" % (rel_err(out["poses_pred_batch"].cpu().numpy(), g["step0/pred"]),
                                                  rel_err(out["mu_pred"].cpu().numpy(), g["step0/mu_pred"])))
    worst = []
    for n, t in tr.grads.items():
        key = "step0/grad/netG.%s" % n
        ref = g[key + "/samples"].astype(np.float64)
        rms = np.sqrt(g[key + "/digest"][1] / t.numel())
        e = np.abs(samples_of(t).astype(np.float64) - ref) / rms
        worst.append((float(e.max()), float(np.median(e)), n))
    for w in worst:
        print("   grad %-55s max %.2e med %.2e" % (w[2], w[0], w[1]))
ops.set_conv_math(0)
