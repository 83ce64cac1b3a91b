"""Diagnostic (not a test): which stream bounds the overlapped step?  Times the CUDA-graph step (B=32, math mode 3) as is, with the
weight gradients skipped, with the FGD / metrics side work skipped, and on a single stream (each variant in its own process: the
diagnostic switches are read at import).      python tests/diag_critical_path.py [batch]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(batch, overlap):
    sys.path.insert(0, ROOT)
    import torch
    import bench
    from speechdrivestemplates_b200 import config, pipeline
    dev = torch.device("cuda:0")
    tr = pipeline.Voice2PoseTrainer(config.get_cfg("voice2pose_sdt_bp"), bench.N_TRAIN, dev, conv_math=3)
    tr.model.clips_code.data.copy_(bench.initial_codes())
    if not overlap:
        tr.set_overlap(False)
    hbs = bench.make_batches(batch, 0, count=2)
    for i in range(6):
        tr.train_step(hbs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 30
    e0.record()
    for _ in range(K):
        tr.run_staged()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"ms_per_step": e0.elapsed_time(e1) / K}))


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    variants = [("full step, three streams", {}, 1), ("weight gradients skipped", {"SDT_DIAG_SKIP_WGRAD": "1"}, 1),
                ("FGD / metrics side work skipped", {"SDT_DIAG_SKIP_SIDE": "1"}, 1),
                ("both skipped (forward + dgrad chain + Adam)", {"SDT_DIAG_SKIP_WGRAD": "1", "SDT_DIAG_SKIP_SIDE": "1"}, 1),
                ("full step, single stream", {}, 0)]
    for name, env, overlap in variants:
        e = dict(os.environ)
        e.update(env)
        out = subprocess.run([sys.executable, __file__, "--child", str(batch), str(overlap)], env=e, capture_output=True, text=True)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]
        print("%-48s %s" % (name, line))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), bool(int(sys.argv[3])))
    else:
        main()
