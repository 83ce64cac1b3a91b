"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total device time, share."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    print("# %s: %d launches, %.2f ms total device time (cold-cache, serialised: compare SHARES)" % (path, n, tot / 1e6))
    print("%-64s %7s %12s %7s" % ("kernel", "n", "total ms", "share"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-64s %7d %12.3f %6.1f%%" % (k[:64], c, t / 1e6, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
