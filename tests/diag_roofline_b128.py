"""Diagnostic (not a test): north-star roofline figures at batch 128 -- mel front end, 2-D encoder forward, UNet + pose decoder
forward -- each kernel family timed with CUDA events around its launches in an eager generator forward.
    python tests/diag_roofline_b128.py [batch]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from speechdrivestemplates_b200 import _lib, config, pipeline  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    dev = torch.device("cuda:0")
    tr = pipeline.Voice2PoseTrainer(config.get_cfg("voice2pose_sdt_bp"), bench.N_TRAIN, dev, use_cuda_graph=False, conv_math=3)
    tr.set_overlap(False)
    hb = bench.make_batches(B, 0, count=1)[0]
    for _ in range(3):
        tr.train_step(hb)
    torch.cuda.synchronize()
    prof = bench.EventProfiler()
    tr._stage(hb)
    _lib.hooks = (prof.pre, prof.post)
    reps = 3
    for _ in range(reps):
        tr.run_staged()
    _lib.hooks = None
    agg = prof.summary()
    pk = bench.peaks()
    # ---- per-family records: (ms per step, launches per step, flops per step)
    rec = {k: (v[0] / reps, v[1] // reps, v[2] / reps) for k, v in agg.items()}
    out = {"batch": B, "hbm_peak_gbs": pk["hbm"], "bf16_sustained_tflops": pk["tf_sustained"]}
    mel_ms = rec["sdt_mel_fwd"][0]
    mel_bytes = 409704.0 * B                                  # audio read once + mel written once
    out["mel"] = {"ms": mel_ms, "algorithmic_GBps": mel_bytes / mel_ms / 1e6, "frac_of_hbm": mel_bytes / mel_ms / 1e6 / pk["hbm"]}
    yt = rec.get("sdt_conv_gemm[tc_conv_ytap_kernel]")
    if yt:
        out["encoder_convs_fwd_plus_dgrad"] = {"ms": yt[0], "launches": yt[1], "TFLOPs": yt[2] / yt[0] / 1e9,
                                               "frac_of_bf16_sustained": yt[2] / yt[0] / 1e9 / pk["tf_sustained"],
                                               "frac_of_tf32_equivalent": yt[2] / yt[0] / 1e9 / (0.5 * pk["tf_sustained"])}
    g = rec.get("sdt_conv_gemm")
    if g:
        out["other_convs_1d_and_ffma"] = {"ms": g[0], "launches": g[1], "TFLOPs": g[2] / g[0] / 1e9}
    out["by_kernel_ms"] = {k: round(v[0], 4) for k, v in sorted(rec.items(), key=lambda kv: -kv[1][0])[:14]}
    out["step_ms_serial_sum"] = sum(v[0] for v in rec.values())
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
