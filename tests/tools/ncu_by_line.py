"""Aggregate an `ncu --page source --csv` SASS listing by CUDA source line, using nvdisasm --print-line-info of the
same cubin (the CSV of a report captured without source import has no line correlation).

  python tests/tools/ncu_by_line.py <src.csv> <all.sass> <mangled kernel name> [top]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# 1. line info per instruction offset
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":"))
off2line = {}
cur = ("?", 0)
inl = []
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2))
# 2. samples per address
rows = list(csv.reader(open(src_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, isamp, iinst, ithr = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base = None
agg = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[h + 1:]:
    if r and r[0] in ("Address", "Kernel Name"):
        break  # next kernel's table
    if len(r) <= ithr or not r[ia]:
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    (f, ln), _ = off2line.get(a - base, (("?", 0), ""))
    v = [int(float(r[isamp] or 0)), int(float(r[iinst] or 0)), int(float(r[ithr] or 0))]
    for k in range(3):
        agg[(f, ln)][k] += v[k]
        tot[k] += v[k]
print(f"total samples {tot[0]}  warp-instructions {tot[1]}  thread-instructions {tot[2]}  (avg active {tot[2] / max(tot[1], 1):.1f})")
print(f"{'file:line':28s} {'samples%':>8s} {'inst%':>7s} {'act':>5s}")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f + ':' + str(ln):28s} {100 * v[0] / tot[0]:8.2f} {100 * v[1] / tot[1]:7.2f} {v[2] / max(v[1], 1):5.1f}")
