"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean us, share."""
import collections
import csv
import sys


def main(path, skip_probe=True):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0]
        if skip_probe and "probe" in k:
            continue
        agg.setdefault(k, []).append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {sum(len(v) for v in agg.values())} launches, {tot / 1e6:.3f} ms total (serialised, cold cache)")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:44s} n={len(v):4d} mean={sum(v) / len(v) / 1e3:10.1f} us  total={sum(v) / 1e6:8.3f} ms  share={sum(v) / tot:6.1%}")


if __name__ == "__main__":
    main(sys.argv[1])
