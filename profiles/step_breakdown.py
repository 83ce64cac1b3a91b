"""Per-launch breakdown of ONE train step from an ncu launch list (the step = launches between two mel_kernel's)."""
import csv, re, sys
def main(path, thresh=0.05):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    idx = [i for i, r in enumerate(rows) if "mel_kernel" in r["Kernel Name"]]
    a, b = idx[0], idx[1]
    tot, out, fam = 0.0, [], {}
    for r in rows[a:b]:
        v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r["Metric Unit"], 1e-6)
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        out.append((k, r["Grid Size"], v)); tot += v
        f_ = fam.setdefault(k, [0, 0.0]); f_[0] += 1; f_[1] += v
    print("# one step: %d launches, %.3f ms device time (ncu, cold cache, serialised)" % (len(out), tot))
    for k, (n, v) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print("%-48s n=%4d %8.3f ms %5.1f%%" % (k[:48], n, v, 100 * v / tot))
    print("# launches above %.2f ms, in order" % thresh)
    for k, g, v in out:
        if v > thresh:
            print("%-48s %-18s %8.3f ms" % (k[:48], g, v))
if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.05)
