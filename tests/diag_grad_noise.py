"""Diagnostic (not a test): per-parameter gradient error of the CUDA path and of the fp32 CPU oracle against the fp64 oracle."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import oliver_stat
from oracle import sdt_oracle as O
from speechdrivestemplates_b200 import pipeline, config
from test_gpu_step import _to_host_batch

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_train = 16
gen = torch.Generator().manual_seed(11)
code0 = 0.1 * torch.randn(n_train, 32, generator=gen)
batch = O.synthetic_batch(bs, n_train, oliver_stat(True), seed=321)
res = {}
for dt in (torch.float64, torch.float32):
    orc = O.Voice2PoseOracle(O.make_cfg("voice2pose_sdt_bp"), n_train, seed=0, dtype=dt)
    orc.sd["clips_code"] = code0.to(dt)
    res[dt] = orc.train_step(batch)[2]
tr = pipeline.Voice2PoseTrainer(config.get_cfg("voice2pose_sdt_bp"), n_train, "cuda:0", use_cuda_graph=False, seed=0)
tr.model.clips_code.data.copy_(code0)
tr.train_step(_to_host_batch(batch))
print("B=%d  %-58s %10s %10s" % (bs, "param", "cuda-vs-64", "cpu32-vs-64"))
for n, t in tr.grads.items():
    ref = res[torch.float64]["netG." + n].numpy()
    rms = np.sqrt((ref ** 2).mean())
    e1 = np.abs(t.cpu().numpy() - ref).max() / rms
    e2 = np.abs(res[torch.float32]["netG." + n].numpy() - ref).max() / rms
    print("     %-58s %10.2e %10.2e" % (n, e1, e2))

# ---- activation sign mismatches (LeakyReLU mask flips) between the CUDA forward and the fp64 oracle forward
taps = {}
orc = O.Voice2PoseOracle(O.make_cfg("voice2pose_sdt_bp"), n_train, seed=0, dtype=torch.float64)
orc.sd["clips_code"] = code0.double()
orc.forward(batch, taps)
bufs = tr.model.netG.engine().arena.bufs
names = {"unet.e%d" % i: "unet.e%d" % i for i in range(7)}
names.update({"unet.d%d" % i: "unet.d%d" % i for i in range(1, 6)})
names.update({"dec.%d" % i: "decoder.%d" % i for i in range(4)})
for tk, bk in names.items():
    ref = taps[tk].permute(0, 2, 1)
    got = bufs["act:" + bk].cpu().double()
    flips = ((ref > 0) != (got > 0))
    idx = flips.nonzero()
    print("%-10s flips=%d  max|act diff|=%.2e %s" % (tk, int(flips.sum()), float((ref - got).abs().max()),
          [(float(ref[tuple(i)]), float(got[tuple(i)])) for i in idx[:3]]))
