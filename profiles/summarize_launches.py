"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: count, total, share, average."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"{'kernel':62s} {'n':>5s} {'total_us':>11s} {'share':>7s} {'avg_us':>9s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:62]:62s} {n:5d} {t:11.1f} {100 * t / tot:6.1f}% {t / n:9.1f}")
    print(f"{'total':62s} {'':5s} {tot:11.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
