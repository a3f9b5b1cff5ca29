"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel family."""
import collections
import csv
import re
import sys


def main(path, tail=0):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
        seq.append((name, v, row.get("Grid Size", ""), row.get("Block Size", "")))
    tot = sum(v[1] for v in agg.values())
    print(f"{'total us':>10} {'count':>6} {'avg us':>8} {'share':>6}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} {n:6d} {t / n:8.2f} {100 * t / tot:5.1f}%  {k[:90]}")
    print(f"{tot:10.1f} total")
    for s in seq[-tail:] if tail else []:
        print(f"{s[1]:8.2f} {s[2]:>14} {s[3]:>12} {s[0][:70]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
