"""Facts about the dominant kernel that only ncu can give, written to the JSON file bench.py reads (never typed into bench.py):
    python profiles/ncu_dominant_facts.py <full.ncu-rep> <launch_list.csv> [out.json] [kernel substring]
* per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and tensor-pipe activity, averaged over the captured
  launches of the kernel (one train step's worth when captured with -c <launches per step>);
* the kernel's share of ONE step in the `--metrics gpu__time_duration.sum` launch list (step = launches between two mel_kernel's).
"""
import csv
import json
import re
import subprocess
import sys


def num(v):
    return float(v.replace(",", ""))


def main():
    rep, lst = sys.argv[1], sys.argv[2]
    out_path = sys.argv[3] if len(sys.argv) > 3 else "profiles/r2_ncu_dominant_kernel.json"
    pat = sys.argv[4] if len(sys.argv) > 4 else "tc_conv_ytap_kernel"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
    per = []
    for r in rows[2:]:
        if pat not in r[ix["Kernel Name"]]:
            continue
        def m(name):
            return num(r[ix[name]]) * scale.get(units[ix[name]], 1.0)
        per.append({"grid": r[ix["Grid Size"]], "time_us": m("gpu__time_duration.sum"),
                    "dram_bytes": m("dram__bytes_read.sum") + m("dram__bytes_write.sum"), "dram_read_bytes": m("dram__bytes_read.sum"),
                    "tensor_pipe_pct": num(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
                    if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in ix else None})
    with open(lst) as f:
        lines = [l for l in f if not l.startswith("==")]
    lrows = list(csv.DictReader(lines))
    idx = [i for i, r in enumerate(lrows) if "mel_kernel" in r["Kernel Name"]]
    a, b = idx[-2], idx[-1]                      # the last complete step of the list
    tot = dom = 0.0
    ndom = 0
    fam = {}
    for r in lrows[a:b]:
        v = num(r["Metric Value"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        fam[k] = fam.get(k, 0.0) + v
        tot += v
        if pat in r["Kernel Name"]:
            dom += v
            ndom += 1
    n = max(len(per), 1)
    tw = sum(p["time_us"] for p in per) or 1.0
    res = {"kernel": pat, "launches_captured": len(per),
           "dram_bytes_per_launch": sum(p["dram_bytes"] for p in per) / n,
           "dram_read_bytes_per_launch": sum(p["dram_read_bytes"] for p in per) / n,
           "tensor_pipe_active_pct_time_weighted": sum((p["tensor_pipe_pct"] or 0.0) * p["time_us"] for p in per) / tw,
           "avg_time_us_under_ncu": tw / n,
           "share_of_step_in_launch_list": dom / tot if tot else None, "launches_per_step_in_launch_list": ndom,
           "step_us_in_launch_list": tot,
           "top_kernels_us_per_step": dict(sorted(((k, round(v, 1)) for k, v in fam.items()), key=lambda kv: -kv[1])[:12]),
           "per_launch": per, "sources": [rep.split("/")[-1], lst.split("/")[-1]],
           "note": "ncu timings are cold-cache and serialised: compare the SHARE, not the absolute"}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k not in ("per_launch",)}, indent=1))


if __name__ == "__main__":
    main()
