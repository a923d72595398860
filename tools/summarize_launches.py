"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (per-launch times are
cold-cache and serialised: compare SHARES, not absolutes)."""
import collections
import csv
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row["Metric Unit"]
        t = t / 1000 if unit == "ns" else (t * 1000 if unit == "ms" else t)
        key = row["Kernel Name"].split("(")[0][:48] + " grid" + row.get("Grid Size", "")
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    print(f"{'us':>10} {'share':>6} {'n':>4}  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:10.1f} {100 * t / tot:5.1f}% {c:4d}  {k}")
    print(f"{tot:10.1f} total us over {sum(c for c, _ in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
