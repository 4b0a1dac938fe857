#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-native SASS instructions in the product library (cuobjdump -sass): the evidence table of
B200_PROFILING.md ("What proves a Blackwell-native kernel").  usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dinov2.cpp_b200", "lib", "libdinov2_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "SYNCS", "MUFU.EX2", "FFMA2", "HMMA", "total"]
rows, cur, arch = collections.OrderedDict(), None, ""
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("dino::", "")
        cur = rows.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for c in cols:
            if c == "total":
                continue
            if c == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    cur[c] += 1
            elif c == "UTCHMMA":
                if op.startswith("UTCHMMA"):
                    cur[c] += 1
            elif op.startswith(c):
                cur[c] += 1
print(f"# SASS evidence: {os.path.basename(lib)} ({arch}, {os.path.getsize(lib)} bytes)\n")
print("`cuobjdump -sass` of the PRODUCT library, instruction counts per kernel (tools/sass_summary.py).  UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2),")
print("UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / .st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add, SYNCS = mbarrier ops,")
print("HMMA would be the legacy mma.sync path (absent).\n")
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---:|" * len(cols))
tot = collections.Counter()
for name, c in rows.items():
    print(f"| `{name}` | " + " | ".join(str(c[k]) for k in cols) + " |")
    tot.update(c)
print("| **all kernels** | " + " | ".join(str(tot[k]) for k in cols) + " |")
print(f"\n{len(rows)} kernels in the library.")
