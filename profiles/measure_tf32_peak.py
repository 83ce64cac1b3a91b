"""cuBLAS TF32 GEMM peak of this B200 ("builder must measure if TF32 MMA is used") -> profiles/r2_tf32_peak.json.
Same recipe as MEASURED_PEAKS.json's bf16 figure: torch.matmul 8192^3 (2*N^3 flop), best of 10 (burst) and back to back for 4 s
(sustained), SM clocks / power sampled through NVML meanwhile.  Also re-measures bf16 beside it for the ratio.
    python profiles/measure_tf32_peak.py [out.json]"""
import json
import sys
import threading
import time

import torch


def sample_clocks(stop, out):
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(0)
        out["sm_max_mhz"] = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not stop.is_set():
            out.setdefault("sm", []).append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            out.setdefault("w", []).append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            stop.wait(0.1)
    except Exception as e:          # noqa: BLE001
        out["error"] = str(e)


def gemm_rate(dtype, tf32, n=8192, sustain_s=4.0):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    flop = 2.0 * n ** 3
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        time.sleep(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(10, int(sustain_s * best * 1e12 / flop))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    e1.synchronize()
    return best, flop * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12, reps


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "profiles/r2_tf32_peak.json"
    stop, clk = threading.Event(), {}
    th = threading.Thread(target=sample_clocks, args=(stop, clk), daemon=True)
    th.start()
    tf_b, tf_s, reps = gemm_rate(torch.float32, True)
    time.sleep(1.0)
    bf_b, bf_s, _ = gemm_rate(torch.bfloat16, False)
    time.sleep(1.0)
    fp_b, fp_s, _ = gemm_rate(torch.float32, False, n=4096, sustain_s=1.0)
    stop.set()
    th.join(timeout=2)
    sm = sorted(clk.get("sm", []))
    res = {"tf32_tflops": tf_b, "tf32_tflops_sustained": tf_s, "bf16_tflops": bf_b, "bf16_tflops_sustained": bf_s,
           "fp32_ffma_tflops": fp_b, "tf32_over_bf16_sustained": tf_s / bf_s,
           "how": "torch.matmul fp32 8192^3 with torch.backends.cuda.matmul.allow_tf32=True (cuBLAS TF32 tensor-core GEMM): best of 10 "
                  "(burst) and %d back to back (sustained); bf16 the same way; fp32 without TF32 at 4096^3" % reps,
           "gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "clocks": {"sm_mhz_median_upper_half": sm[len(sm) * 3 // 4] if sm else None, "sm_mhz_min": sm[0] if sm else None,
                      "sm_max_mhz": clk.get("sm_max_mhz"), "power_w_max": max(clk.get("w", [0.0])), "samples": len(sm)},
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
