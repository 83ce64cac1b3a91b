"""Benchmark of the Voice2Pose SDT-BP train step (BASELINE.json: configs[1], metric clips/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

* own arm: one process per GPU (torchrun for N>1), per-GPU batch 32 x 64-frame clips (weak scaling), synthetic data,
  one flat NCCL all-reduce per step; prints ONE JSON line with `value` (inputs resident in HBM), `e2e` (host batch ->
  H2D -> step -> D2H of the loss scalars every step, through Voice2PoseTrainer.run_epoch), `roofline` of the dominant kernel
  family measured live with CUDA events, `cpu_baseline` (the CPU oracle port on the host cores, bounded sample),
  `clocks` sampled with nvidia-smi during the timed regions.
* --impl reference: the reference's CPU implementation of the same step = the oracle port (the reference is Python
  and /root/reference does not exist on the GPU box), all host threads, rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_TRAIN = 32768          # synthetic number of training clips
METRIC = "sdt_bp_train_clips_per_s"
UNIT = "clips/s"


def oliver_stat():
    z = np.load(os.path.join(ROOT, "tests", "golden", "speaker_stat_oliver.npz"))
    return {"mean": z["parted_mean"], "std": z["parted_std"], "scale_factor": float(z["parted_scale_factor"])}


def workload_name(batch):
    return "voice2pose_sdt_bp train step, batch %d x 64-frame clips (68266 samples @16 kHz, 121 keypoints) per GPU" % batch


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clocks / throttle reasons during the timed regions (B200_PROFILING.md recipe), sampled in-process through NVML
    every 20 ms.  (A polling `nvidia-smi -lms` child was measured to stall the host-synchronised e2e loop by ~2 ms per
    step -- each query takes driver locks -- so it is only the fallback when NVML cannot be loaded.)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.how = None

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for n, b in bits.items():
                    if r & b:
                        self.reasons.add(n)
            except Exception:            # noqa: BLE001 - a failed sample is just skipped
                pass
            self._stop.wait(0.02)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES may remap indices: resolve through the PCI bus id of the torch device
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(nv.nvmlDeviceGetCount()):
                    hi = nv.nvmlDeviceGetHandleByIndex(i)
                    if int(nv.nvmlDeviceGetPciInfo(hi).bus) == int(bus):
                        h = hi
                        break
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.how = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:                # noqa: BLE001 - fall back to the nvidia-smi child
            self.how = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.how = "nvidia-smi"
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.how == "nvml":
            self._stop.set()
            self.thread.join(timeout=2)
            sm = self.samples
            if not sm:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "how": "nvml"}
            busy = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
            return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm),
                    "how": "nvml, 20 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(self.NAMES, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "how": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------------
def make_batches(batch, rank, count=4):
    from oracle import sdt_oracle as O     # only for the shared seeded synthetic-input generator
    st = oliver_stat()
    out = []
    for i in range(count):
        b = O.synthetic_batch(batch, N_TRAIN, st, seed=1000 + 17 * rank + i)
        hb = {"audio": b["audio"].pin_memory(), "poses": b["poses"].pin_memory(), "clip_index": b["clip_index"].pin_memory(),
              "num_frames": b["num_frames"],
              "speaker_stat": {k: torch.from_numpy(np.asarray(v)).pin_memory() for k, v in b["speaker_stat"].items()}}
        out.append(hb)
    return out


def cpu_oracle_clips_per_s(batch, steps, warmup, threads):
    """Reference CPU path (oracle port): forward + backward + Adam of voice2pose_sdt_bp on the host cores."""
    from oracle import sdt_oracle as O
    torch.set_num_threads(threads)
    orc = O.Voice2PoseOracle(O.make_cfg("voice2pose_sdt_bp"), N_TRAIN, seed=0)
    orc.sd["clips_code"] = 0.1 * torch.randn(N_TRAIN, 32, generator=torch.Generator().manual_seed(11))
    st = oliver_stat()
    times = []
    for i in range(warmup + steps):
        b = O.synthetic_batch(batch, N_TRAIN, st, seed=2000 + i)
        t0 = time.perf_counter()
        _l, _r, grads = orc.train_step(b)
        orc.apply_optimizers(grads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return batch * len(times) / sum(times), sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = args.batch
    # bounded sample: keep the whole run within a few minutes on any host
    t0 = time.perf_counter()
    v1, s1 = cpu_oracle_clips_per_s(min(batch, 8), 1, 0, threads)
    est = s1 * (batch / min(batch, 8)) * (args.steps + args.warmup)
    sample_batch = batch if est < 240 else 8
    steps = args.steps if s1 * (sample_batch / min(batch, 8)) * (args.steps + args.warmup) < 300 else max(2, int(240 / max(s1, 1e-3)) - args.warmup)
    value, sec = cpu_oracle_clips_per_s(sample_batch, steps, args.warmup, threads)
    sample = "%d timed steps of %d clips (of the %d-clip batch) after %d warm-up" % (steps, sample_batch, batch, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(batch), "per_gpu_batch": batch,
                   "note": "CPU oracle port of the reference step (reference is Python/torch; /root/reference is absent on the GPU box)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
_JSON_OUT = None


def claim_stdout():
    """stdout carries the ONE JSON line.  Keep the real stdout for it and point fd 1 at stderr, so that nothing a library writes
    to stdout (NCCL's version banner, a C-level printf) can end up next to it."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


class EventProfiler:
    """Brackets every C-ABI kernel call of one eager step with CUDA events on the launching (current) stream."""

    def __init__(self):
        self.records = []

    def pre(self, name, args):
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        flops = 0.0
        if name == "sdt_conv_gemm_multi":             # the parity classes of one data gradient (one launch in math mode 3)
            import ctypes as C
            from speechdrivestemplates_b200 import _lib
            ds = [args[0][i] for i in range(args[1])]
            flops = sum(2.0 * d.B * d.GH * d.GW * d.N * d.TH * d.TW * d.C for d in ds)
            plan = (C.c_int32 * 10)()
            _lib.load().sdt_conv_plan(C.byref(ds[0]), plan)
            name = "sdt_conv_gemm[tc_conv_ytap_kernel]" if plan[0] in (3, 4) else "sdt_conv_gemm"
        elif name in ("sdt_conv_gemm", "sdt_conv_wgrad"):
            d = args[0]._obj
            flops = 2.0 * d.B * d.GH * d.GW * d.N * d.TH * d.TW * d.C
            if name == "sdt_conv_gemm":
                # which kernel does the library launch for this descriptor?  3 = tc_conv_ytap_kernel (2-D convolutions)
                import ctypes as C
                from speechdrivestemplates_b200 import _lib
                plan = (C.c_int32 * 10)()
                _lib.load().sdt_conv_plan(args[0], plan)
                if plan[0] in (3, 4):
                    name = "sdt_conv_gemm[tc_conv_ytap_kernel]"
        return (name, e0, flops)

    def post(self, tok):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.records.append(tok + (e1,))

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, e0, flops, e1 in self.records:
            a = agg.setdefault(name, [0.0, 0, 0.0])
            a[0] += e0.elapsed_time(e1)
            a[1] += 1
            a[2] += flops
        return agg


def run_own(args):
    import torch.distributed as dist
    from speechdrivestemplates_b200 import _lib, config, pipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        # stdout carries the ONE JSON line: NCCL's own banner / debug lines (NCCL_DEBUG=VERSION|WARN|INFO in the environment) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    B, K, W = args.batch, args.steps, max(args.warmup, 3)

    conv_math = {"fp32": 0, "tf32": 1, "tf32-tma": 2, "tf32-reuse": 3, "tf32-pair": 4}[args.math]
    tr = pipeline.Voice2PoseTrainer(config.get_cfg("voice2pose_sdt_bp"), N_TRAIN, dev, use_cuda_graph=not args.no_graph,
                                    process_group=pg, seed=0, conv_math=conv_math)
    tr.model.clips_code.data.copy_(0.1 * torch.randn(N_TRAIN, 32, generator=torch.Generator().manual_seed(11)))
    host_batches = make_batches(B, rank)
    dev_batches = []
    for hb in host_batches:
        dev_batches.append({"audio": hb["audio"].to(dev), "poses": hb["poses"].to(dev), "clip_index": hb["clip_index"].to(dev),
                            "speaker_stat": {k: v.to(dev) for k, v in hb["speaker_stat"].items()}})

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (eager twice -> CUDA-graph capture -> replays)
    for i in range(W + 2):
        tr.train_step(host_batches[i % len(host_batches)])
    torch.cuda.synchronize()
    launches_per_step = tr.kernels_per_step

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # ---- value: inputs resident in HBM (a device->device restage of the rotating batch is inside the region)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        tr._stage(dev_batches[k % len(dev_batches)])
        tr.run_staged()
    e1.record()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))

    # ---- e2e: pinned host batch -> H2D -> step -> D2H of the loss scalars, every step
    # Voice2PoseTrainer.run_epoch = the reference's `for batch in dataloader: train_step; log` loop: every step's batch
    # is copied host->device (copy stream, one batch ahead) and every step's scalars are read back (one step behind)
    tr.run_epoch((host_batches[k % len(host_batches)] for k in range(W)), on_losses=lambda i, d: None)   # warm-up of this path
    barrier()
    t0 = time.perf_counter()
    e0.record()
    per_step, stamps = [], [t0]

    def on_losses(i, d):
        per_step.append(d)
        stamps.append(time.perf_counter())
    tr.run_epoch((host_batches[k % len(host_batches)] for k in range(K)), on_losses=on_losses)
    e1.record()
    assert len(per_step) == K
    gaps = sorted((b - a) * 1e3 for a, b in zip(stamps[:-1], stamps[1:]))
    last = per_step[-1]
    barrier()
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), 0.0))
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    hb = host_batches[0]
    h2d = sum(t.numel() * t.element_size() for t in (hb["audio"], hb["poses"], hb["clip_index"])) + \
        sum(t.numel() * t.element_size() for t in hb["speaker_stat"].values())
    d2h = 8 * 8

    # ---- roofline of the dominant kernel family: one eager step bracketed with CUDA events per launch
    prof = EventProfiler()
    graphs, tr.use_graph = tr._graphs, False
    tr.set_overlap(False)                  # single stream: every launch is timed alone
    tr._stage(dev_batches[0])
    tr.run_staged()                        # un-profiled eager step (re-warm)
    _lib.hooks = (prof.pre, prof.post)
    reps = 3
    for _ in range(reps):
        tr.run_staged()
    _lib.hooks = None
    agg = prof.summary()
    tr.set_overlap(True)
    tr._graphs, tr.use_graph = graphs, not args.no_graph
    total_ms = sum(a[0] for a in agg.values())
    fam_keys = ("sdt_conv_gemm", "sdt_conv_gemm[tc_conv_ytap_kernel]", "sdt_conv_wgrad")
    conv_ms = sum(agg.get(k, [0, 0, 0])[0] for k in fam_keys)
    conv_fl = sum(agg.get(k, [0, 0, 0])[2] for k in fam_keys)
    conv_n = sum(agg.get(k, [0, 0, 0])[1] for k in fam_keys)
    pk = peaks()
    family = {"kernels": "all implicit-GEMM convolution launches (2-D + 1-D forward / data-gradient, weight gradients, FFMA leftovers)",
              "achieved_tflops": conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0,
              "share_of_step": conv_ms / total_ms if total_ms else None, "launches_per_step": conv_n // reps}
    dom = agg.get("sdt_conv_gemm[tc_conv_ytap_kernel]")
    if dom is not None and dom[0] > 0:
        # dominant kernel: the persistent tcgen05 kernel of the 2-D encoder convolutions (forward + data gradient)
        achieved = dom[2] / (dom[0] * 1e-3) / 1e12
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
            "frac": achieved / pk["tf_sustained"],
            # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the 14 launches of one step in one `ncu --set full`
            # capture of this command (profiles/r1_ncu_tc_conv_mode3_final.txt: 1762.6 MB over 14 launches)
            "traffic": 125.9e6,
            "kernel": "tc_conv_ytap_kernel (tcgen05 kind::tf32, TMA operands, persistent, double-buffered TMEM): 2-D encoder "
                      "convolutions forward + data gradient",
            "algorithmic_flops_per_launch": dom[2] / max(dom[1], 1),
            # share of the serial sum of event-bracketed launches; every bracket around one of the ~250 small launches also
            # holds a few us of launch latency, so this reads lower than the share in the ncu launch list of the same command
            # (profiles/r1_step_breakdown_mode3.txt: 1.081 of 4.039 ms)
            "share_of_step": dom[0] / total_ms if total_ms else None,
            "share_of_step_ncu_launch_list": 0.268,
            "launches_per_step": dom[1] // reps, "avg_launch_ms": dom[0] / max(dom[1], 1),
            "peak_source": pk["src"] + " bf16 dense, sustained (kernel timed inside a long step); the kernel computes in TF32, whose "
                                       "nominal dense rate is half of bf16 (1.1 vs 2.25 PFLOP/s)",
            "ncu_tensor_pipe_active_pct": "51-85 (Cout >= 128), 32-51 (64-channel stride-2 layer): profiles/r1_ncu_tc_conv_mode3_final.txt",
            # the same measured peak scaled to the arithmetic type the kernel uses (TF32 = half the bf16 rate)
            "frac_of_tf32_equivalent_peak": achieved / (0.5 * pk["tf_sustained"]),
        }
    else:
        achieved = family["achieved_tflops"]
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
            "frac": achieved / pk["tf_sustained"], "traffic": None,
            "kernel": "implicit-GEMM convolutions: " + ("conv_gemm_kernel + conv_wgrad_kernel (fp32 FFMA)" if conv_math == 0
                                                          else "tc_conv_kernel + tc_wgrad_kernel (tcgen05 TF32) + FFMA kernels for ineligible layers"),
            "share_of_step": family["share_of_step"], "launches_per_step": family["launches_per_step"],
            "avg_launch_ms": conv_ms / max(conv_n, 1),
            "peak_source": pk["src"] + " bf16 dense, sustained (kernel timed inside a long step)",
        }
    roofline["conv_family"] = family
    roofline["by_kernel_ms_per_step"] = {k: round(v[0] / reps, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:8]}
    roofline["hbm_peak_gbs"] = pk["hbm"]

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline: oracle port on the host cores, bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        v, sec = cpu_oracle_clips_per_s(8, 2, 1, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "2 timed steps of 8 clips (of the %d-clip batch) after 1 warm-up, torch CPU fp32, %.2f s/step" % (B, sec)}

    clips = B * world * K
    line = {
        "metric": METRIC, "value": clips / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if conv_math == 0 else "tf32",
        "data": "synthetic",
        "config": {"workload": workload_name(B), "per_gpu_batch": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                   "n_train_clips": N_TRAIN, "cuda_graph": tr._graphs is not None,
                   "l2": "no explicit flush: each step streams ~%.1f GB of activations/gradients (>> 126 MB L2) and rotates over 4 distinct input batches" % (tr.model.netG.engine().arena.nbytes() / 1e9),
                   "conv_math": ["fp32 FFMA", "tcgen05 TF32 operands, fp32 accumulate (FFMA for ineligible layers)",
                                 "tcgen05 TF32, TMA-fed operands for forward/dgrad, fp32 accumulate (FFMA for ineligible layers)",
                                 "tcgen05 TF32, TMA-fed operands reused across vertical taps and accumulators in shared memory, fp32 "
                                 "accumulate (FFMA for ineligible layers)",
                                 "as tf32-reuse plus CTA pairs (cta_group::2) where N % 128 == 0 (experimental)"][conv_math]},
        "e2e": {"value": clips / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K, "wall_s": wall_e2e, "host_gap_ms": {"median": gaps[len(gaps) // 2], "max": gaps[-1]}},
        "gpu_launches": launches_per_step * K * 2,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "last_losses": last,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU (BASELINE configs[1]: 32)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--math", default="tf32-reuse", choices=["fp32", "tf32", "tf32-tma", "tf32-reuse", "tf32-pair"],
                    help="convolution math: fp32 FFMA kernels or tcgen05 TF32 tensor-core kernels (fp32 accumulate)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
