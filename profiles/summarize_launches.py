"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of device time per kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0][:80]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / tot * 100:6.2f}%  n={v[0]:4d}  avg={v[1] / v[0] / 1e3:10.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
