"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, mean, share."""
import collections
import csv
import re
import sys


def main(path, per=1):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"]
        m = re.search(r"(\w+_kernel)", name)
        key = "conv3x3_pair_kernel" if "conv3x3_pair" in name else (m.group(1) if m else name[:40])
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else v * 1000 if row["Metric Unit"] == "ms" else v
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'mean us':>9s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:44]:44s} {v[0]:8d} {v[1] / 1000:10.3f} {v[1] / v[0]:9.1f} {100 * v[1] / tot:6.1f}%")
    print(f"{'total':44s} {sum(v[0] for v in agg.values()):8d} {tot / 1000:10.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
