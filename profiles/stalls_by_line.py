"""Join an `ncu --page source --csv` dump (per-SASS-instruction stall samples) with `nvdisasm -g` line info and
print warp-stall samples per source line and per stall reason.

  cuobjdump -xelf all pepflowww_b200/libpepflow_b200.so          # -> pf_*.sm_100a.cubin
  nvdisasm -g pf_ipa_v2.sm_100a.cubin > all.sass
  ncu -i prof.ncu-rep --page source --csv > prof_source.csv
  python profiles/stalls_by_line.py prof_source.csv all.sass <mangled kernel name> [top N]
"""
import collections
import csv
import re
import sys


def line_map(sass_path, kernel):
    """offset (int) -> (file, line) for one .text section of an nvdisasm -g listing."""
    out, cur, on = {}, None, False
    for ln in open(sass_path):
        if ln.startswith(".text."):
            on = ln.strip().rstrip(":") == ".text." + kernel
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            out[int(m.group(1), 16)] = cur
    return out


def main():
    src_csv, sass, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lm = line_map(sass, kernel)
    rows = list(csv.reader(open(src_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    per_line = collections.Counter()
    per_line_reason = collections.defaultdict(collections.Counter)
    per_line_inst = collections.Counter()
    reasons = collections.Counter()
    total = 0
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[0], 16)
        base = addr if base is None else base
        key = lm.get(addr - base, ("?", 0))
        n = int(r[col["# Samples"]] or 0)
        per_line[key] += n
        per_line_inst[key] += int(r[col["Instructions Executed"]] or 0)
        total += n
        for s in stall_cols:
            v = int(r[col[s]] or 0)
            if v:
                per_line_reason[key][s[6:]] += v
                reasons[s[6:]] += v
    print(f"total samples {total}")
    print("by reason: " + "  ".join(f"{k} {100 * v / total:.1f}%" for k, v in reasons.most_common(10)))
    print(f"{'file:line':28s} {'samples':>8s} {'share':>6s} {'inst':>10s}  top reasons")
    for key, n in per_line.most_common(top):
        rs = "  ".join(f"{k} {v}" for k, v in per_line_reason[key].most_common(4))
        print(f"{key[0] + ':' + str(key[1]):28s} {n:8d} {100 * n / total:5.1f}% {per_line_inst[key]:10d}  {rs}")


if __name__ == "__main__":
    main()
