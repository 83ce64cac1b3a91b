"""Benchmarks of the Voice2Pose / Pose2Pose hot path (BASELINE.json: metric on configs[1]; configs[2..4] through --config).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference] [--config sdt_bp|sdt_vae|pose2pose|demo] [--batch B]

* own arm: one process per GPU (torchrun for N>1), fixed per-GPU batch (weak scaling), synthetic data.  ONE JSON line with
  `value` (inputs resident in HBM), `e2e` (pinned host batch -> H2D -> step -> D2H of the loss scalars, every step, through the
  trainers' public API), `roofline` of the dominant kernel (measured live with CUDA events; algorithmic FLOPs), the north-star
  sub-rooflines `mel_enc` / `decoder` / `mel` (graph-replayed forward segments) at the bench batch and at batch 128, `cpu_baseline`
  (the CPU oracle port on the host cores, bounded sample), `gpu_torch_baseline` (the same oracle graph in stock torch on this
  GPU with cuDNN TF32: the existing Blackwell kernels this work has to beat), `clocks` sampled through NVML during the timed
  regions and, for N>1, `rank_param_spread` (max - min over ranks of every parameter after the timed loop; must be 0.0).
* --impl reference: the reference's CPU implementation of the same step = the oracle port (the reference is Python over torch and
  /root/reference does not exist on the GPU box), all host threads, rank 0 only, SAME per-GPU batch, steps and warm-up at every N.
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
UNIT = "clips/s"
AUDIO_LEN, FRAMES = 68266, 64
# ---- algorithmic work per clip
MEL_BYTES = 409704.0                       # audio read once + mel written once
MEL_ENC_BYTES = 39794344.0                 # fused mel + 2-D encoder forward, layer-boundary traffic
MEL_ENC_FLOPS = 7118.6e6
DEC_BYTES, DEC_WEIGHT_BYTES = 1366528.0, 14.24e6      # UNet + pose decoder forward: activations per clip, weights per step
DEC_FLOPS = 243.36e6
STEP_FLOPS = 22190e6                       # sdt_bp / sdt_vae full train step
STEP_BYTES, STEP_FIXED_BYTES = 102.4e6, 283e6
P2P_FLOPS, P2P_BYTES, P2P_FIXED_BYTES = 710e6, 4.0e6, 127e6
GEN_FWD_FLOPS = 7356.0e6

CONFIGS = {
    # name: (metric, default per-GPU batch, reference cfg name, BASELINE.json config it measures)
    "sdt_bp": ("sdt_bp_train_clips_per_s", 32, "voice2pose_sdt_bp", "configs[1]: voice2pose_sdt_bp train step, batch 32 per GPU"),
    "sdt_vae": ("sdt_vae_train_clips_per_s", 16, "voice2pose_sdt_vae", "configs[2]: voice2pose_sdt_vae train step, global batch 128 on 8 GPUs = 16 per GPU"),
    "pose2pose": ("pose2pose_train_clips_per_s", 32, "pose2pose", "configs[3]: pose2pose VAE train step, global batch 256 on 8 GPUs = 32 per GPU"),
    "demo": ("demo_frames_per_s", 1, "voice2pose_sdt_bp", "configs[4]: demo inference, 600 s of 16 kHz audio -> 9000 frames x 121 keypoints"),
}


def oliver_stat():
    z = np.load(os.path.join(ROOT, "tests", "golden", "speaker_stat_oliver.npz"))
    return {"mean": z["parted_mean"], "std": z["parted_std"], "scale_factor": float(z["parted_scale_factor"])}


def workload_name(config, batch):
    if config == "demo":
        return "demo inference: one 600 s / 16 kHz utterance -> 9000 frames x 121 keypoints per GPU, host wav in / host poses out"
    what = {"sdt_bp": "voice2pose_sdt_bp train step", "sdt_vae": "voice2pose_sdt_vae train step (external frozen clip codes)",
            "pose2pose": "pose2pose VAE train step"}[config]
    return "%s, batch %d x 64-frame clips (68266 samples @16 kHz, 121 keypoints) per GPU" % (what, batch)


def config_dict(config, batch, world):
    """Identical for both arms (the reference arm is the same workload on the host cores)."""
    return {"workload": workload_name(config, batch), "per_gpu_batch": batch, "global_batch": batch * world,
            "parallelism": "dp%d" % world if config != "demo" else "%d independent replicas" % world, "n_train_clips": N_TRAIN,
            "baseline_config": CONFIGS[config][3],
            "l2": "no explicit flush: every step streams far more than the 126 MB L2 (activations + gradients: ~1.8 GB at batch 32) "
                  "and rotates over 4 distinct input batches"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        out = dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="MEASURED_PEAKS.json")
    # TF32 cuBLAS peak measured on this pool's B200 by profiles/measure_tf32_peak.py (committed with its clock record)
    tf32 = os.path.join(ROOT, "profiles", "r2_tf32_peak.json")
    if os.path.exists(tf32):
        with open(tf32) as f:
            t = json.load(f)
        out["tf32_burst"], out["tf32_sustained"] = t.get("tf32_tflops"), t.get("tf32_tflops_sustained")
    return out


def ncu_file_facts():
    """Figures that only an ncu capture can give (DRAM traffic per launch, the kernel's share of the serialised launch list).  They
    are read from the committed summary of the capture, never typed into this file; absent file -> null."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_dominant_kernel.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


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
            props = torch.cuda.get_device_properties(self.index)
            bus = props.pci_bus_id if hasattr(props, "pci_bus_id") else None
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
        busy = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "how": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------------
# synthetic inputs: seeded, identical for the GPU arm, the CPU arm and the tests
# ------------------------------------------------------------------------------------------------
def synthetic_batch(batch_size, seed):
    g = torch.Generator().manual_seed(seed)
    audio = 0.1 * torch.randn(batch_size, AUDIO_LEN, generator=g)
    poses = torch.randn(batch_size, FRAMES, 2, 121, generator=g)
    idx = torch.randint(0, N_TRAIN, (batch_size,), generator=g)
    st = oliver_stat()
    return {"audio": audio, "poses": poses, "clip_index": idx, "num_frames": torch.full((batch_size,), FRAMES, dtype=torch.long),
            "speaker": ["oliver"] * batch_size,
            "speaker_stat": {"mean": np.tile(st["mean"][None], (batch_size, 1)), "std": np.tile(st["std"][None], (batch_size, 1)),
                             "scale_factor": np.full((batch_size,), st["scale_factor"], np.float64)}}


def make_batches(batch, rank, count=4):
    out = []
    for i in range(count):
        b = synthetic_batch(batch, 1000 + 17 * rank + i)
        out.append({"audio": b["audio"].pin_memory(), "poses": b["poses"].pin_memory(), "clip_index": b["clip_index"].pin_memory(),
                    "num_frames": b["num_frames"],
                    "speaker_stat": {k: torch.from_numpy(np.asarray(v)).pin_memory() for k, v in b["speaker_stat"].items()}})
    return out


def initial_codes():
    return 0.1 * torch.randn(N_TRAIN, 32, generator=torch.Generator().manual_seed(11))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle port of the reference step
# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(config, batch, steps, warmup, threads):
    """(units per second, seconds per step) of the oracle port on the host cores: forward + backward + Adam of the train
    configs, mel + generator forward of the demo."""
    from oracle import sdt_oracle as O
    torch.set_num_threads(threads)
    times = []
    if config == "demo":
        cfg = O.make_cfg("voice2pose_sdt_bp")
        sd = O.init_voice2pose(cfg, 16, seed=0)
        nf = 9000
        audio = 0.1 * torch.randn(1, int(nf * 16000 / 15), generator=torch.Generator().manual_seed(8))
        code = 0.1 * torch.randn(1, 32, generator=torch.Generator().manual_seed(9))
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            with torch.no_grad():
                mel = O.mel_spectrogram(audio, sd["mel_transfm.spectrogram.window"], sd["mel_transfm.mel_scale.fb"])
                O.generator_forward(mel, nf, code, sd, cfg, False, "netG.")
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        return nf * len(times) / sum(times), sum(times) / len(times)
    if config == "pose2pose":
        orc = O.Pose2PoseOracle(O.make_cfg("pose2pose"), N_TRAIN, seed=0)
    else:
        orc = O.Voice2PoseOracle(O.make_cfg(CONFIGS[config][2]), N_TRAIN, seed=0)
        if config == "sdt_bp":
            orc.sd["clips_code"] = initial_codes()
        table = initial_codes()
    for i in range(warmup + steps):
        b = synthetic_batch(batch, 2000 + i)
        t0 = time.perf_counter()
        if config == "pose2pose":
            _l, _r, grads = orc.train_step(b, torch.randn(batch, 32))
        else:
            if config == "sdt_vae":
                b["external_code_table"] = table
            _l, _r, grads = orc.train_step(b)
        orc.apply_optimizers(grads)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1
    metric, default_batch = CONFIGS[args.config][0], CONFIGS[args.config][1]
    batch = args.batch or default_batch
    t0 = time.perf_counter()
    # The sample is the FULL per-GPU batch at every N (same clips per step as the own arm).  Only the step count is bounded:
    # one probe step sizes it so that the run ends within ~4 minutes on any host.
    _v, sec = cpu_oracle_rate(args.config, batch, 1, 0, threads)
    steps = max(1, min(args.steps, int(200.0 / max(sec, 1e-3)) - args.warmup))
    warmup = min(args.warmup, max(0, int(40.0 / max(sec, 1e-3))))
    value, sec = cpu_oracle_rate(args.config, batch, steps, warmup, threads)
    unit = "frames/s" if args.config == "demo" else UNIT
    sample = "%d timed steps of the full %d-clip per-GPU batch after %d warm-up (the same at every --gpus N)" % (steps, batch, warmup)
    if args.config == "demo":
        sample = "%d timed 600 s utterances after %d warm-up" % (steps, warmup)
    emit({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(args.config, batch, world),
        "note": "CPU oracle port of the reference step on rank 0's host cores (the reference is Python over torch; /root/reference "
                "is absent on the GPU box); value = clips/s of ONE rank's share of the global batch",
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    })


def gpu_torch_rate(config, batch, steps, warmup, dev):
    """`gpu_torch_baseline`: the oracle's torch graph of the same step on THIS GPU with stock torch kernels (cuDNN convolutions
    with TF32 allowed = the reference's own GPU default, main.py sets nothing else) -- the 'existing Blackwell kernels' bar of
    SURVEY §2.2.  Test-infrastructure use of the oracle, outside every timed region of the own arm.  Omits the f64 final results
    + metrics of the step (numpy in the oracle), which flatters this baseline slightly."""
    from oracle import sdt_oracle as O
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = False          # torch's default
    torch.backends.cudnn.benchmark = True                  # main.py:39
    if config == "pose2pose":
        orc = O.Pose2PoseOracle(O.make_cfg("pose2pose"), N_TRAIN, seed=0)
    else:
        orc = O.Voice2PoseOracle(O.make_cfg(CONFIGS[config][2]), N_TRAIN, seed=0)
        if config == "sdt_bp":
            orc.sd["clips_code"] = initial_codes()
    orc.sd = {k: v.to(dev) for k, v in orc.sd.items()}
    table = initial_codes().to(dev)
    batches = []
    for i in range(4):
        b = synthetic_batch(batch, 3000 + i)
        b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
        b["external_code_table"] = table
        batches.append(b)

    def step(b):
        if config == "pose2pose":
            _l, _r, grads = orc.train_step(b, torch.randn(batch, 32, device=dev))
        else:
            leaves = list(orc.g_names) + (["clips_code"] if orc.code_trainable else [])
            for n in leaves:
                orc.sd[n].requires_grad_(True)
            losses, _res = orc.forward(b)
            gs = torch.autograd.grad(losses["G_loss"], [orc.sd[n] for n in leaves], allow_unused=True)
            for n in leaves:
                orc.sd[n].requires_grad_(False)
            grads = {n: (g if g is not None else torch.zeros_like(orc.sd[n])) for n, g in zip(leaves, gs)}
        orc.apply_optimizers(grads)

    for i in range(warmup):
        step(batches[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(batches[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return batch / (ms * 1e-3), ms


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
    """Brackets every C-ABI kernel call of one eager step with CUDA events on the launching (current) stream.  FLOPs are
    ALGORITHMIC: 2 x (output pixels of the forward convolution) x N x taps x C -- a data gradient is credited with its forward
    problem's FLOPs, not with the gather-form grid it happens to run on (the 6x3 / no-padding layer executes 2.08x those)."""

    def __init__(self):
        self.records = []

    @staticmethod
    def _alg_flops(d):
        pixels = d.SH * d.SW if d.ty_mul < 0 else d.GH * d.GW          # dgrad descriptors walk taps backwards (ops.dgrad_desc)
        return 2.0 * d.B * pixels * d.N * d.TH * d.TW * d.C

    @staticmethod
    def _exec_flops(d):
        return 2.0 * d.B * d.GH * d.GW * d.N * d.TH * d.TW * d.C

    def pre(self, name, args):
        import ctypes as C
        from speechdrivestemplates_b200 import _lib
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        flops = executed = 0.0
        if name == "sdt_conv_gemm_multi":             # the parity classes of one data gradient (one launch in math mode 3)
            ds = [args[0][i] for i in range(args[1])]
            flops, executed = sum(self._alg_flops(d) for d in ds), sum(self._exec_flops(d) for d in ds)
            plan = (C.c_int32 * 10)()
            _lib.load().sdt_conv_plan(C.byref(ds[0]), plan)
            name = "sdt_conv_gemm[tc_conv_ytap_kernel]" if plan[0] in (3, 4) else "sdt_conv_gemm"
        elif name in ("sdt_conv_gemm", "sdt_conv_wgrad"):
            d = args[0]._obj
            flops, executed = self._alg_flops(d), self._exec_flops(d)
            if name == "sdt_conv_gemm":
                plan = (C.c_int32 * 10)()
                _lib.load().sdt_conv_plan(args[0], plan)
                if plan[0] in (3, 4):
                    name = "sdt_conv_gemm[tc_conv_ytap_kernel]"
        return (name, e0, flops, executed)

    def post(self, tok):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.records.append(tok + (e1,))

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, e0, flops, executed, e1 in self.records:
            a = agg.setdefault(name, [0.0, 0, 0.0, 0.0])
            a[0] += e0.elapsed_time(e1)
            a[1] += 1
            a[2] += flops
            a[3] += executed
        return agg


def time_graph(fn, reps, pre_each=None):
    """Capture fn() into a CUDA graph and return the mean device time (ms) of `reps` replays (CUDA events on the replay stream)."""
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for i in range(reps):
        if pre_each is not None:
            pre_each(i)
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1)
    del g
    return total / reps


def forward_segments(tr, dev_batches, pk, reps=20):
    """North-star sub-rooflines on the trainer's own engines: the fused mel + 2-D encoder forward, the whole generator forward
    (UNet + pose decoder = the difference) and the mel kernel alone, each as a CUDA graph replayed `reps` times over rotating
    input batches (each segment streams far more than L2 between replays: the encoder forward alone writes 40 MB per clip)."""
    m = tr.model
    eng = m.netG.engine()
    B = dev_batches[0]["audio"].shape[0]
    audio = torch.empty_like(dev_batches[0]["audio"])
    code = torch.empty(B, 32, device=audio.device).normal_(0, 0.1)
    gp = {n: p.detach() for n, p in m.netG.named_parameters()}
    mel_buf = torch.empty(B, 80, 1 + AUDIO_LEN // 160, device=audio.device)

    def rotate(i):
        audio.copy_(dev_batches[i % len(dev_batches)]["audio"])

    wg, eng.wg_stream = getattr(eng, "wg_stream", None), None       # forward only: weight prep inline, single stream
    try:
        t_mel = time_graph(lambda: m.mel_transfm(audio, out=mel_buf), reps, rotate)
        t_enc = time_graph(lambda: eng.forward(m.mel_transfm(audio, out=mel_buf), FRAMES, code, gp, True, None, upto="encoder"), reps, rotate)
        t_gen = time_graph(lambda: eng.forward(m.mel_transfm(audio, out=mel_buf), FRAMES, code, gp, True, None), reps, rotate)
        t_prep = time_graph(lambda: eng.wprep.run(), reps)
    finally:
        eng.wg_stream = wg
    t_enc_net = t_enc - t_prep                       # the per-step weight-operand refresh is not part of the forward roofline
    t_dec = max(t_gen - t_enc, 1e-6)
    enc_hbm_ms = MEL_ENC_BYTES * B / (pk["hbm"] * 1e9) * 1e3
    enc_tensor_ms = MEL_ENC_FLOPS * B / (pk["tf_sustained"] * 1e12) * 1e3
    dec_hbm_ms = (DEC_BYTES * B + DEC_WEIGHT_BYTES) / (pk["hbm"] * 1e9) * 1e3
    dec_tensor_ms = DEC_FLOPS * B / (pk["tf_sustained"] * 1e12) * 1e3
    return {
        "batch": B, "how": "CUDA-graph replays (%d) of forward segments of the trainer's engines, events on the replay stream; "
                           "mel_enc excludes the per-step weight-operand refresh (%.3f ms, sdt_weight_prep_batch)" % (reps, t_prep),
        "mel": {"ms": t_mel, "hbm_gbs": MEL_BYTES * B / t_mel / 1e6, "frac_of_hbm": MEL_BYTES * B / t_mel / 1e6 / pk["hbm"]},
        "mel_enc": {"ms": t_enc_net, "hbm_gbs": MEL_ENC_BYTES * B / t_enc_net / 1e6,
                    "frac_of_hbm": MEL_ENC_BYTES * B / t_enc_net / 1e6 / pk["hbm"],
                    "tflops": MEL_ENC_FLOPS * B / t_enc_net / 1e9, "frac_of_bf16_sustained": MEL_ENC_FLOPS * B / t_enc_net / 1e9 / pk["tf_sustained"],
                    "tighter_roofline": "hbm" if enc_hbm_ms >= enc_tensor_ms else "tensor",
                    "frac_of_tighter_roofline": max(enc_hbm_ms, enc_tensor_ms) / t_enc_net,
                    # the kernels compute in TF32: against the MEASURED cuBLAS TF32 peak the tensor roofline is the tighter one
                    "frac_of_tf32_sustained": (MEL_ENC_FLOPS * B / t_enc_net / 1e9 / pk["tf32_sustained"]) if pk.get("tf32_sustained") else None},
        "decoder": {"ms": t_dec, "what": "UNet_1D + pose decoder forward (generator forward minus mel + encoder)",
                    "hbm_gbs": (DEC_BYTES * B + DEC_WEIGHT_BYTES) / t_dec / 1e6, "tflops": DEC_FLOPS * B / t_dec / 1e9,
                    "tighter_roofline": "hbm" if dec_hbm_ms >= dec_tensor_ms else "tensor",
                    "frac_of_tighter_roofline": max(dec_hbm_ms, dec_tensor_ms) / t_dec},
        "generator_fwd_ms": t_gen,
    }


# ------------------------------------------------------------------------------------------------
def build_trainer(config, batch, dev, pg, args, conv_math):
    from speechdrivestemplates_b200 import checkpoint, config as C, pipeline
    if config == "pose2pose":
        return pipeline.Pose2PoseTrainer(C.get_cfg("pose2pose"), N_TRAIN, dev, use_cuda_graph=not args.no_graph, process_group=pg,
                                         seed=0, conv_math=conv_math)
    if config == "sdt_vae":
        # external frozen clip codes + FGD encoder weights come from a pose2pose-format checkpoint (voice2pose.py:40-55,234-242):
        # written here from a freshly initialised pose VAE with seeded codes (no trained checkpoint exists offline)
        path = os.path.join("/tmp", "sdt_b200_bench_p2p_rank%s.pth" % os.environ.get("RANK", "0"))
        p2p = pipeline.Pose2PoseTrainer(C.get_cfg("pose2pose"), N_TRAIN, dev, use_cuda_graph=False, seed=3, conv_math=conv_math)
        p2p.model.clip_code_mu.copy_(initial_codes())
        checkpoint.save_pose2pose(p2p, path, 0, 0)
        del p2p
        cfg = C.get_cfg("voice2pose_sdt_vae", ["VOICE2POSE.POSE_ENCODER.AE_CHECKPOINT", path])
        return pipeline.Voice2PoseTrainer(cfg, N_TRAIN, dev, use_cuda_graph=not args.no_graph, process_group=pg, seed=0, conv_math=conv_math)
    tr = pipeline.Voice2PoseTrainer(C.get_cfg("voice2pose_sdt_bp"), N_TRAIN, dev, use_cuda_graph=not args.no_graph, process_group=pg,
                                    seed=0, conv_math=conv_math)
    tr.model.clips_code.data.copy_(initial_codes())
    return tr


def run_epoch_generic(tr, batches, on_losses):
    """Voice2PoseTrainer.run_epoch, or the same software pipeline spelled out for trainers without it (pose2pose)."""
    if hasattr(tr, "run_epoch"):
        return tr.run_epoch(batches, on_losses=on_losses)
    k = 0
    for b in batches:
        out = tr.train_step(b)
        on_losses(k, tr.losses_to_host(out))
        k += 1
    return k


def run_train(args):
    import torch.distributed as dist
    from speechdrivestemplates_b200 import _lib

    for var in ("SDT_DIAG_SKIP_WGRAD", "SDT_DIAG_SKIP_SIDE"):
        if os.environ.get(var):
            raise SystemExit("%s is set: that diagnostic switch skips work of the step; refusing to benchmark" % var)
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
    config = args.config
    metric, default_batch = CONFIGS[config][0], CONFIGS[config][1]
    B, K, W = args.batch or default_batch, args.steps, max(args.warmup, 3)
    conv_math = {"fp32": 0, "tf32": 1, "tf32-tma": 2, "tf32-reuse": 3, "tf32-pair": 4}[args.math]
    tr = build_trainer(config, B, dev, pg, args, conv_math)
    host_batches = make_batches(B, rank)
    dev_batches = [{"audio": hb["audio"].to(dev), "poses": hb["poses"].to(dev), "clip_index": hb["clip_index"].to(dev),
                    "speaker_stat": {k: v.to(dev) for k, v in hb["speaker_stat"].items()}} for hb in host_batches]

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

    # ---- e2e: pinned host batch -> H2D -> step -> D2H of the loss scalars, every step, through the public API
    # (Voice2PoseTrainer.run_epoch = the reference's `for batch in dataloader: train_step; log` loop: every step's batch is copied
    # host->device on a copy stream one batch ahead, every step's scalars are read back one step behind)
    run_epoch_generic(tr, (host_batches[k % len(host_batches)] for k in range(W)), lambda i, d: None)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    per_step, stamps = [], [t0]

    def on_losses(i, d):
        per_step.append(d)
        stamps.append(time.perf_counter())
    run_epoch_generic(tr, (host_batches[k % len(host_batches)] for k in range(K)), on_losses)
    e1.record()
    assert len(per_step) == K
    gaps = sorted((b - a) * 1e3 for a, b in zip(stamps[:-1], stamps[1:]))
    last = per_step[-1]
    barrier()
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), 0.0))
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    hb = host_batches[0]
    keys = ("poses", "clip_index") if config == "pose2pose" else ("audio", "poses", "clip_index")
    h2d = sum(hb[k].numel() * hb[k].element_size() for k in keys) + sum(t.numel() * t.element_size() for t in hb["speaker_stat"].values())
    d2h = len(getattr(tr, "_scal", torch.zeros(8))) * 8

    # ---- multi-GPU consistency: every rank must hold bit-identical parameters after the timed loops
    spread = None
    if world > 1:
        hi, lo = tr.flat_p.clone(), tr.flat_p.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        spread = float((hi - lo).abs().max())

    pk = peaks()
    roofline, segments, b128 = None, None, None
    if config in ("sdt_bp", "sdt_vae"):
        roofline = conv_roofline(tr, dev_batches, args, pk, conv_math)
        if rank == 0 and not args.no_segments:
            segments = forward_segments(tr, dev_batches, pk)
    else:
        step_ms = ms_dev / K
        bytes_step = P2P_BYTES * B + P2P_FIXED_BYTES
        roofline = {"bound": "hbm", "achieved": bytes_step / step_ms / 1e6, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": bytes_step / step_ms / 1e6 / pk["hbm"], "traffic": None,
                    "kernel": "whole pose2pose step (a chain of ~%d launches of a few microseconds each: latency-bound, SURVEY 8d); "
                              "algorithmic bytes = B x 4.0 MB + 127 MB of weights / gradients / Adam state" % launches_per_step,
                    "peak_source": pk["src"]}

    if world > 1:
        dist.barrier()
    if rank != 0:
        finish(tr, world)
        return

    # ---- batch-128 block (north-star target batch): step time + the same sub-rooflines, on a second trainer
    if config == "sdt_bp" and world == 1 and not args.no_b128 and B != 128:
        tr.close()
        b128 = batch128_block(args, dev, pk, conv_math)

    # ---- CPU baseline: oracle port on the host cores, bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        cb = min(B, 8)
        v, sec = cpu_oracle_rate(config, cb, 2, 1, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "2 timed steps of %d clips (of the %d-clip batch) after 1 warm-up, torch CPU fp32, %.2f s/step" % (cb, B, sec)}
    gpu_torch = None
    if world == 1 and not args.no_gpu_torch:
        try:
            v, ms = gpu_torch_rate(config, B, 10, 4, dev)
            gpu_torch = {"value": v, "unit": UNIT, "ms_per_step": ms, "kind": "oracle graph in stock torch %s on this GPU: cuDNN convolutions with "
                         "TF32 allowed (the reference's GPU default), cudnn.benchmark, eager autograd, torch.optim-equivalent Adam; "
                         "without the step's f64 final results + metrics" % torch.__version__,
                         "sample": "10 timed steps of the full %d-clip batch after 4 warm-up, inputs resident in HBM" % B}
        except Exception as e:              # noqa: BLE001 - a baseline leg must never take the bench line down
            gpu_torch = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}

    clips = B * world * K
    line = {
        "metric": metric, "value": clips / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if conv_math == 0 else "tf32",
        "data": "synthetic", "config": config_dict(config, B, world),
        "impl_detail": {"cuda_graph": tr._graphs is not None or b128 is not None, "graphs_per_step": len(tr._graphs) if tr._graphs else (1 if b128 else 0),
                        "comm_mode": getattr(tr, "comm_mode", "none"),
                        "p2p_nvls_multicast": (bool(getattr(getattr(tr, "_px", None), "multicast", 0)) if getattr(tr, "_px", None) is not None else None),
                        "conv_math": ["fp32 FFMA", "tcgen05 TF32 operands, fp32 accumulate (FFMA for ineligible layers)",
                                      "tcgen05 TF32, TMA-fed operands for forward/dgrad, fp32 accumulate (FFMA for ineligible layers)",
                                      "tcgen05 kind::tf32 (operands stored round-to-nearest TF32, fp32 accumulation in TMEM), TMA-fed operands reused "
                                      "across vertical taps and accumulators in shared memory; FFMA for the 1->64 first layer",
                                      "as tf32-reuse plus CTA pairs (cta_group::2) where N % 128 == 0 (experimental)"][conv_math]},
        "e2e": {"value": clips / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K, "wall_s": wall_e2e, "host_gap_ms": {"median": gaps[len(gaps) // 2], "max": gaps[-1]}},
        "gpu_launches": launches_per_step * K * 2,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "gpu_torch_baseline": gpu_torch,
        "last_losses": last,
    }
    if config in ("sdt_bp", "sdt_vae"):
        step_ms = ms_dev / K
        line["step_roofline"] = {
            "algorithmic_flops_per_clip": STEP_FLOPS, "algorithmic_bytes_per_clip": STEP_BYTES, "fixed_bytes_per_step": STEP_FIXED_BYTES,
            "tflops": STEP_FLOPS * B / step_ms / 1e9, "frac_of_bf16_sustained": STEP_FLOPS * B / step_ms / 1e9 / pk["tf_sustained"],
            "hbm_gbs": (STEP_BYTES * B + STEP_FIXED_BYTES) / step_ms / 1e6, "frac_of_hbm": (STEP_BYTES * B + STEP_FIXED_BYTES) / step_ms / 1e6 / pk["hbm"]}
        if segments is not None:
            line["mel_enc_hbm_gbs"] = segments["mel_enc"]["hbm_gbs"]
            line["mel_enc_frac_of_hbm"] = segments["mel_enc"]["frac_of_hbm"]
            line["north_star"] = segments
    if b128 is not None:
        line["b128"] = b128
    if spread is not None:
        line["rank_param_spread"] = spread
    emit(line)
    finish(tr, world)


def conv_roofline(tr, dev_batches, args, pk, conv_math):
    """Roofline of the dominant kernel family: one eager single-stream step bracketed with CUDA events per launch."""
    from speechdrivestemplates_b200 import _lib
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
    conv_ms = sum(agg.get(k, [0, 0, 0, 0])[0] for k in fam_keys)
    conv_fl = sum(agg.get(k, [0, 0, 0, 0])[2] for k in fam_keys)
    conv_n = sum(agg.get(k, [0, 0, 0, 0])[1] for k in fam_keys)
    family = {"kernels": "all implicit-GEMM convolution launches (2-D + 1-D forward / data-gradient, weight gradients, FFMA leftovers)",
              "achieved_tflops": conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0,
              "share_of_step": conv_ms / total_ms if total_ms else None, "launches_per_step": conv_n // reps}
    dom = agg.get("sdt_conv_gemm[tc_conv_ytap_kernel]")
    facts = ncu_file_facts()
    if dom is not None and dom[0] > 0:
        achieved = dom[2] / (dom[0] * 1e-3) / 1e12
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
            "traffic": facts.get("dram_bytes_per_launch") if facts else None,
            "kernel": "tc_conv_ytap_kernel (tcgen05 kind::tf32, TMA operands, persistent, double-buffered TMEM): 2-D encoder "
                      "convolutions forward + data gradient",
            "algorithmic_flops_per_launch": dom[2] / max(dom[1], 1),
            "executed_over_algorithmic_flops": dom[3] / dom[2] if dom[2] else None,
            "share_of_step": dom[0] / total_ms if total_ms else None,
            "share_note": "share of the serial sum of event-bracketed launches of one eager single-stream step; each bracket around one of "
                          "the ~250 small launches also holds a few us of launch latency, so this reads lower than the ncu launch list",
            "launches_per_step": dom[1] // reps, "avg_launch_ms": dom[0] / max(dom[1], 1),
            "peak_source": pk["src"] + " bf16 dense, sustained (kernel timed inside a long step); the kernel computes in TF32",
            "frac_of_bf16_burst": achieved / pk["tf_burst"],
        }
        if pk.get("tf32_sustained"):
            roofline["tf32_peak_measured"] = {"burst": pk["tf32_burst"], "sustained": pk["tf32_sustained"],
                                              "source": "profiles/r2_tf32_peak.json (cuBLAS TF32 8192^3 on this pool's B200)"}
            roofline["frac_of_tf32_sustained"] = achieved / pk["tf32_sustained"]
        if facts:
            roofline["ncu"] = facts
    else:
        achieved = family["achieved_tflops"]
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
            "frac": achieved / pk["tf_sustained"], "traffic": None,
            "kernel": "implicit-GEMM convolutions: " + ("conv_gemm_kernel + conv_wgrad_kernel (fp32 FFMA)" if conv_math == 0
                                                          else "tc_conv_kernel + tc_wgrad_kernel (tcgen05 TF32) + FFMA kernels for ineligible layers"),
            "share_of_step": family["share_of_step"], "launches_per_step": family["launches_per_step"],
            "avg_launch_ms": conv_ms / max(conv_n, 1), "peak_source": pk["src"] + " bf16 dense, sustained",
        }
    roofline["conv_family"] = family
    roofline["by_kernel_ms_per_step"] = {k: round(v[0] / reps, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:10]}
    roofline["by_kernel_note"] = ("serial single-stream sum %.3f ms; the timed step overlaps three streams (weight gradients and the FGD / "
                                  "metrics side work run beside the data-gradient chain), so this sum exceeds ms_per_step" % (total_ms / reps))
    roofline["hbm_peak_gbs"] = pk["hbm"]
    return roofline


def batch128_block(args, dev, pk, conv_math):
    """north_star's target batch: the same trainer at 128 clips per GPU -- step time, dominant-kernel roofline, sub-rooflines."""
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    B = 128
    tr = build_trainer("sdt_bp", B, dev, None, args, conv_math)
    hbs = make_batches(B, 0, count=2)
    dbs = [{"audio": hb["audio"].to(dev), "poses": hb["poses"].to(dev), "clip_index": hb["clip_index"].to(dev),
            "speaker_stat": {k: v.to(dev) for k, v in hb["speaker_stat"].items()}} for hb in hbs]
    for i in range(5):
        tr.train_step(hbs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for k in range(K):
        tr._stage(dbs[k % 2])
        tr.run_staged()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    roof = conv_roofline(tr, dbs, args, pk, conv_math)
    seg = forward_segments(tr, dbs, pk, reps=10)
    out = {"batch": B, "ms_per_step": ms, "clips_per_s": B / (ms * 1e-3), "steps": K,
           "step_tflops": STEP_FLOPS * B / ms / 1e9, "step_frac_of_bf16_sustained": STEP_FLOPS * B / ms / 1e9 / pk["tf_sustained"],
           "dominant_kernel": {k: roof[k] for k in ("achieved", "frac", "unit", "launches_per_step", "avg_launch_ms", "share_of_step") if k in roof},
           "mel": seg["mel"], "mel_enc": seg["mel_enc"], "decoder": seg["decoder"]}
    tr.close()
    return out


def run_demo(args):
    """configs[4]: 600 s of 16 kHz audio -> 9000-frame pose stream: host wav in, host poses out, one utterance per GPU."""
    from speechdrivestemplates_b200 import config as C, data, inference
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    conv_math = {"fp32": 0, "tf32": 1, "tf32-tma": 2, "tf32-reuse": 3, "tf32-pair": 4}[args.math]
    seconds = 600
    alen, nf = data.parse_audio_length(seconds * 16000, 16000, 15)
    torch.manual_seed(0)
    store = tuple(int(x) for x in args.store_layers.split(",") if x != "")
    gen = inference.StreamingGenerator(C.get_cfg("voice2pose_sdt_bp"), dev, conv_math=conv_math, chunk_frames=args.chunk_frames, store_layers=store,
                                       graph=args.chunk_frames > 0 and not args.no_graph)
    audio = (0.1 * torch.randn(1, alen, generator=torch.Generator().manual_seed(8 + rank))).pin_memory()
    code = 0.1 * torch.randn(1, 32, generator=torch.Generator().manual_seed(9))
    K, W = args.steps, max(args.warmup, 3)
    for _ in range(W):
        out = gen(audio, nf, code)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        out = gen(audio, nf, code)            # H2D of the wav, mel + generator, D2H of the pose stream: the public call
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # device-resident variant: the wav already in HBM, poses left in HBM
    a_dev = audio.to(dev)
    c_dev = code.to(dev)
    e0.record()
    for _ in range(K):
        gen.forward_device(a_dev, nf, c_dev)
    e1.record()
    torch.cuda.synchronize()
    ms_dev = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_dev], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev = float(t.item())
        dist.barrier()
    if rank != 0:
        os._exit(0)
    clocks = sampler.stop()
    pk = peaks()
    flops = GEN_FWD_FLOPS * (nf / 64.0)
    tfl = flops / (ms_dev / K) / 1e9
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        v, sec = cpu_oracle_rate("demo", 1, 1, 0, threads)
        cpu = {"value": v, "unit": "frames/s", "cores": threads, "kind": "port", "sample": "1 utterance of 600 s, torch CPU fp32, %.2f s" % sec}
    emit({
        "metric": CONFIGS["demo"][0], "value": nf * world * K / (ms_dev * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if conv_math == 0 else "tf32",
        "data": "synthetic", "config": config_dict("demo", 1, world),
        "impl_detail": {"chunk_frames": gen.chunk_frames, "chunks": gen.last_chunks, "store_layers": list(gen.store_layers),
                        "cuda_graph": gen.use_graph,
                        "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                        "x_realtime": seconds / (ms / K * 1e-3), "out_shape": list(out.shape)},
        "e2e": {"value": nf * world * K / (ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": alen * 4, "d2h_bytes_per_step": nf * 242 * 4,
                "ms_per_step": ms / K},
        "gpu_launches": gen.launches_per_call * K * 2, "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": tfl, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tfl / pk["tf_sustained"], "traffic": None,
                     "kernel": "whole generator forward over the utterance (7,356 MFLOP per 64 frames, SURVEY 8d)", "peak_source": pk["src"]},
        "cpu_baseline": cpu,
    })
    os._exit(0)


def finish(tr, world):
    """Leave without tearing NCCL down: destroying the communicator after its kernels were captured into CUDA graphs has been seen
    to hang at exit (tests/multi_gpu_check.py); the graphs are dropped, every rank passes a barrier, the OS reclaims the rest."""
    if world > 1:
        import torch.distributed as dist
        tr.close()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        if _JSON_OUT is not None:
            _JSON_OUT.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--config", default="sdt_bp", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU (default: the BASELINE config's per-GPU share)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--math", default="tf32-reuse", choices=["fp32", "tf32", "tf32-tma", "tf32-reuse", "tf32-pair"],
                    help="convolution math: fp32 FFMA kernels or tcgen05 TF32 tensor-core kernels (fp32 accumulate)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gpu-torch", action="store_true", help="skip the gpu_torch_baseline leg")
    ap.add_argument("--no-b128", action="store_true", help="skip the batch-128 block")
    ap.add_argument("--no-segments", action="store_true", help="skip the forward-segment sub-rooflines")
    ap.add_argument("--chunk-frames", type=int, default=0, help="demo: frames per streamed chunk (0 = one-shot forward)")
    ap.add_argument("--store-layers", default="3,5", help="demo, streamed: encoder layers whose full-length raw map is kept (inference.py)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "demo":
        run_demo(args)
    else:
        run_train(args)


if __name__ == "__main__":
    main()
