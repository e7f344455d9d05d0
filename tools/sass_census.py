"""Instruction census of libavt_b200.so per kernel (cuobjdump -sass): the Blackwell-native mnemonics that prove tcgen05 / TMEM / TMA.

    python tools/sass_census.py > profiles/r02_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "avt_b200", "libavt_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "MUFU.EX2",
        "FFMA2", "HMMA", "RED.E.ADD", "ATOMG"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                per[cur][k] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            per[cur]["UTCHMMA.2CTA"] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: SASS instruction census per kernel (sm_100a)")
print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UTCBAR = tcgen05.commit, SYNCS = mbarrier\n")
cols = [k for k in KEYS if any(c[k] for c in per.values())]
print("| kernel | SASS instr | " + " | ".join(cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
tot = collections.Counter()
for name, c in per.items():
    print(f"| `{name[:80]}` | {c['total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
    tot.update(c)
print(f"| **all {len(per)} kernels** | {tot['total']} | " + " | ".join(str(tot[k]) for k in cols) + " |")
