"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_agg.py launches.csv [--last-fraction 0.5] [--top 30]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    frac = float(sys.argv[sys.argv.index("--last-fraction") + 1]) if "--last-fraction" in sys.argv else 1.0
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    data = []
    for r in rows[hdr + 1:]:
        if len(r) > vi and r[vi]:
            v = float(r[vi].replace(",", ""))
            v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(r[ui], 1.0)
            data.append((r[ki], v))
    data = data[int(len(data) * (1 - frac)):]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in data:
        k = re.sub(r"\(.*", "", k)
        k = re.sub(r"^void ", "", k)
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{len(data)} launches, {tot / 1e3:.2f} ms of kernel time")
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{v / 1e3:9.3f} ms {100 * v / tot:5.1f}%  n={c:5d}  {k[:110]}")


if __name__ == "__main__":
    main()
