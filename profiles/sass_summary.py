#!/usr/bin/env python
"""Per-kernel SASS summary of the built library (runs on CPU): cuobjdump -sass pl-viwo_b200/libplviwo_fe.so, then for every
kernel the instruction count and the mnemonics that show how it moves data (UBLKCP = cp.async.bulk through the TMA unit,
SYNCS = mbarrier, LDG.E.128 / LDG.E.64 vector loads, REDUX warp reductions, LDS.128, VIMNMX / VABSDIFF4-style byte SIMD).
    python profiles/sass_summary.py > profiles/sass_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "pl-viwo_b200", "libplviwo_fe.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = {"UBLKCP": r"\bUBLKCP", "SYNCS": r"\bSYNCS", "LDG.128": r"\bLDG\.E(\.\w+)*\.128", "LDG.64": r"\bLDG\.E(\.\w+)*\.64",
       "LDG.U8": r"\bLDG\.E\.U8", "LDS.128": r"\bLDS\.128", "STG.128": r"\bSTG\.E(\.\w+)*\.128", "REDUX": r"\bREDUX",
       "SHFL": r"\bSHFL", "ATOMS": r"\bATOMS", "ATOMG/RED": r"\b(ATOMG|RED)\b", "PRMT": r"\bPRMT", "VIADDMNMX/SIMD4": r"\b(VABSDIFF4|VIMNMX|VIADDMNMX)",
       "DFMA/DMUL": r"\b(DFMA|DMUL|DADD)"}
cur, rows, excerpts = None, collections.OrderedDict(), {}
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("plviwo::", "")
        rows[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in ln:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", ln)
    if not m:
        continue
    ins = m.group(1)
    rows[cur]["instr"] += 1
    for k, p in pat.items():
        if re.search(p, ins):
            rows[cur][k] += 1
            if k in ("UBLKCP", "SYNCS") and len(excerpts.setdefault(cur, [])) < 6:
                excerpts[cur].append(ins.strip())
print("SASS summary of pl-viwo_b200/libplviwo_fe.so (sm_100a), %d kernels" % len(rows))
cols = list(pat)
print("%-44s %7s " % ("kernel", "instr") + " ".join("%9s" % c[:9] for c in cols))
for k, c in rows.items():
    print("%-44s %7d " % (k[:44], c["instr"]) + " ".join("%9d" % c[x] for x in cols))
print("\nbulk-copy (TMA unit) and mbarrier instructions, verbatim:")
for k, v in excerpts.items():
    print(" ", k)
    for ins in v:
        print("     ", ins)
