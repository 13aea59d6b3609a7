"""Static counts of the Blackwell-specific / tensor-core SASS instructions per kernel of libpepflow_b200.so (cuobjdump -sass;
no GPU needed).   python profiles/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pepflowww_b200", "libpepflow_b200.so")
KEEP = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTCBAR", "UTCATOMSWS", "HMMA", "LDSM", "CCTL")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, cur, i = collections.OrderedDict(), None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", names[i]); i += 1
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            base = op.split(".")[0]
            if base in KEEP:
                key = base + (".2CTA" if ".2CTA" in op else "")
                counts[cur][key] += 1
    print("cuobjdump -sass pepflowww_b200/libpepflow_b200.so (round 2 build, sm_100a): static counts of the Blackwell-specific / "
          "tensor-core instructions per kernel")
    print("UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG / UTMASTG = TMA "
          "tensor load / store,\nUBLKCP = cp.async.bulk, SYNCS = mbarrier ops, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / "
          "dealloc, HMMA = legacy mma.sync, LDSM = ldmatrix.\n")
    for k, c in counts.items():
        if c:
            print(f"{k:60s} " + "  ".join(f"{n} {v}" for n, v in sorted(c.items())))


if __name__ == "__main__":
    sys.exit(main())
