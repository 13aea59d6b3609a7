"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals over the last
N launches (one Euler iteration of bench.py).  python profiles/summarize_launches.py <csv> [launches_per_step]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 203
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in rows[-per_step:]:
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
        tot[name] += ms
        cnt[name] += 1
    total = sum(tot.values())
    print(f"{path}: last {per_step} launches = one Euler iteration; {total:.2f} ms serialised under ncu")
    print(f"{'kernel':42s} {'n':>4s} {'ms':>9s} {'share':>7s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k:42s} {cnt[k]:4d} {v:9.3f} {100 * v / total:6.1f}%")


if __name__ == "__main__":
    main()
