"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals over the last Euler
iteration of bench.py (delimited by euler_step_kernel launches).  python profiles/summarize_launches.py <csv>"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    names = [re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("pf::", "") for r in rows]
    marks = [i for i, n in enumerate(names) if n.startswith("euler_step_kernel")]
    a, b = marks[-2] + 1, marks[-1] + 1
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row, name in list(zip(rows, names))[a:b]:
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
        tot[name] += ms
        cnt[name] += 1
    total = sum(tot.values())
    print(f"{path}: {b - a} launches = one Euler iteration; {total:.2f} ms serialised under ncu (cold cache)")
    print(f"{'kernel':42s} {'n':>4s} {'ms':>9s} {'share':>7s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k:42s} {cnt[k]:4d} {v:9.3f} {100 * v / total:6.1f}%")


if __name__ == "__main__":
    main()
