#!/usr/bin/env python
"""Static SASS instruction mix of libottocov.so per kernel: python tools/sass_summary.py [out.md]
(cuobjdump -sass, mnemonic counts; shows which data-movement / ranking instructions the kernels are made of)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "otto_recommender_b200", "libottocov.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_instruction_mix.md")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
MN = ["VOTE", "R2P", "MATCH", "ATOMS", "ATOMG", "RED", "SHFL", "LDS", "STS", "LDG", "STG", "BAR", "UBLKCP", "UTMALDG", "UTMASTG",
      "SYNCS", "LDGSTS", "POPC", "BREV", "FLO", "IMAD", "LOP3", "NANOSLEEP"]


def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(.*", "", r).replace("void ", "")


rows, tot = [], collections.Counter()
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ins = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M)
    if len(ins) < 40:
        continue
    c = collections.Counter(i.split(".")[0] for i in ins)
    rows.append((demangle(name), len(ins), c))
    tot.update(c)
rows.sort(key=lambda r: -r[1])
out = ["# SASS instruction mix of the shipped libottocov.so (sm_100a cubins), round 2",
       "",
       "`python tools/sass_summary.py`: `cuobjdump -sass`, static mnemonic counts per kernel (not executed counts).",
       "",
       "No `UBLKCP` / `UTMALDG` / `UTMASTG` / `SYNCS` anywhere: the library uses no TMA / bulk-copy / mbarrier data movement.  The",
       "tile-movement kernels keep their keys in registers between the coalesced load and the shared-memory staging (DESIGN.md 5.2",
       "explains why a bulk copy into shared memory does not fit: the staging buffer already fills the SM, and per-digit output runs",
       "start at 8-byte, not 16-byte, boundaries); tile ranking = `VOTE` ballots + `R2P` (rs_onesweep, stable) or `ATOMS`",
       "(expand_scatter, hash_reduce, unstable), look-back = `NANOSLEEP` polling on volatile status words.",
       "",
       "| kernel | total | " + " | ".join(MN) + " |", "|" + "---|" * (len(MN) + 2)]
for d, n, c in rows:
    out.append(f"| `{d[:72]}` | {n} | " + " | ".join(str(c.get(m, 0)) for m in MN) + " |")
out.append(f"| **all {len(rows)} kernels** | {sum(r[1] for r in rows)} | " + " | ".join(str(tot.get(m, 0)) for m in MN) + " |")
open(out_path, "w").write("\n".join(out) + "\n")
print(f"wrote {out_path}: {len(rows)} kernels, {sum(r[1] for r in rows)} instructions")
