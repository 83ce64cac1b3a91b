"""Diagnostic (not a test): where does the host-fed step loop spend its time?  python tests/diag_e2e.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from speechdrivestemplates_b200 import config, pipeline  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    tr = pipeline.Voice2PoseTrainer(config.get_cfg("voice2pose_sdt_bp"), bench.N_TRAIN, dev, conv_math=3)
    hbs = bench.make_batches(32, 0)
    dbs = [{"audio": h["audio"].to(dev), "poses": h["poses"].to(dev), "clip_index": h["clip_index"].to(dev),
            "speaker_stat": {k: v.to(dev) for k, v in h["speaker_stat"].items()}} for h in hbs]
    for i in range(6):
        tr.train_step(hbs[i % 4])
    torch.cuda.synchronize()
    K = 20

    def timed(name, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        t_cpu = time.perf_counter() - t0
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        print("%-48s %.3f ms/step (host enqueue %.3f ms/step)" % (name, t / K * 1e3, t_cpu / K * 1e3), flush=True)

    def dev_loop():
        for k in range(K):
            tr._stage(dbs[k % 4])
            tr.run_staged()
    timed("device-resident batches", dev_loop)

    def graphs_only():
        for k in range(K):
            tr.run_staged()
    timed("graph replays only", graphs_only)

    def h2d_main_nosync():
        for k in range(K):
            tr.train_step(hbs[k % 4])
    timed("H2D on the main stream, no loss read", h2d_main_nosync)

    def h2d_main_sync():
        for k in range(K):
            tr.losses_to_host(tr.train_step(hbs[k % 4]))
    timed("H2D on the main stream, blocking loss read", h2d_main_sync)

    def prefetch_nolosses():
        tr.run_epoch((hbs[k % 4] for k in range(K)))
    timed("run_epoch, no loss callback", prefetch_nolosses)

    def prefetch_losses():
        tr.run_epoch((hbs[k % 4] for k in range(K)), on_losses=lambda i, d: None)
    timed("run_epoch, losses every step", prefetch_losses)

    # phase timing inside run_epoch
    ph = {"prefetch": 0.0, "train_step": 0.0, "post": 0.0, "collect": 0.0}
    it = iter(hbs[k % 4] for k in range(K))
    nxt = next(it)
    tr.prefetch(nxt)
    pending = None
    k = 0
    torch.cuda.synchronize()
    while nxt is not None:
        t0 = time.perf_counter()
        tr.train_step(nxt)
        t1 = time.perf_counter()
        nxt = next(it, None)
        if nxt is not None:
            tr.prefetch(nxt)
        t2 = time.perf_counter()
        tr.post_losses(k & 1)
        t3 = time.perf_counter()
        if pending is not None:
            tr.collect_losses(pending & 1)
        t4 = time.perf_counter()
        ph["train_step"] += t1 - t0
        ph["prefetch"] += t2 - t1
        ph["post"] += t3 - t2
        ph["collect"] += t4 - t3
        pending = k
        k += 1
    torch.cuda.synchronize()
    print({k2: round(v / K * 1e3, 3) for k2, v in ph.items()})

    # with the bench's clock sampler running
    for rep in range(3):
        smp = bench.ClockSampler(0)
        smp.start()
        time.sleep(0.2)
        timed("run_epoch + NVML sampler (rep %d)" % rep, prefetch_losses)
        print("   ", smp.stop())
        timed("run_epoch, sampler off (rep %d)" % rep, prefetch_losses)

    # H2D alone
    def h2d_only():
        for k in range(K):
            tr.prefetch(hbs[k % 4])
    timed("prefetch copies alone", h2d_only)


if __name__ == "__main__":
    main()
