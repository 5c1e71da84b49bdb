"""Aggregate `ncu -i rep --page source --csv --print-source cuda,sass` by CUDA source line.
  python tests/tools/ncu_src_lines.py <csv> [top]
Per (file:line): stall samples, share of warp instructions, average active lanes, dominant stall reasons."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = defaultdict(lambda: defaultdict(float))
cur_file, cur_line, hdr = "?", 0, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        saddr = r.index("Address")
        continue
    if r[0] in ("Function Name", "File Name", "Kernel Name"):
        continue
    if hdr is None:
        continue
    if r[0].strip().isdigit():
        cur_line = int(r[0])
    if len(r) <= saddr or not r[saddr].startswith("0x"):
        continue
    key = (cur_file, cur_line)
    def g(name):
        try:
            return float(r[hdr[name]])
        except Exception:
            return 0.0
    a = agg[key]
    a["samples"] += g("# Samples")
    a["inst"] += g("Instructions Executed")
    a["thr"] += g("Thread Instructions Executed")
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            a[h] += g(h)
tot_s = sum(a["samples"] for a in agg.values())
tot_i = sum(a["inst"] for a in agg.values())
tot_t = sum(a["thr"] for a in agg.values())
print(f"total samples {tot_s:.0f}  warp-instructions {tot_i:.0f}  thread-instructions {tot_t:.0f}  (avg active {tot_t / max(tot_i, 1):.1f})")
st = defaultdict(float)
for a in agg.values():
    for h, v in a.items():
        if h.startswith("stall_"):
            st[h] += v
print("stalls: " + ", ".join(f"{h[6:]} {100 * v / max(1, sum(st.values())):.1f}%" for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
print(f"{'file:line':28s} {'samples%':>8s} {'inst%':>7s} {'act':>5s}  top stalls")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    ss = sorted(((h[6:], v) for h, v in a.items() if h.startswith("stall_") and v > 0), key=lambda kv: -kv[1])[:3]
    print(f"{f + ':' + str(ln):28s} {100 * a['samples'] / tot_s:8.2f} {100 * a['inst'] / tot_i:7.2f} {a['thr'] / max(a['inst'], 1):5.1f}  " + " ".join(f"{h}:{v:.0f}" for h, v in ss))

# share of warp instructions and samples per file and per 50-line bucket
buckets = defaultdict(lambda: [0.0, 0.0, 0.0])
for (f, ln), a in agg.items():
    b = buckets[(f, 50 * (ln // 50))]
    b[0] += a["samples"]; b[1] += a["inst"]; b[2] += a["thr"]
print("\nper 50-line bucket (samples%, inst%, lanes):")
for (f, l0), b in sorted(buckets.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{f}:{l0}-{l0 + 49:<6d} {100 * b[0] / tot_s:7.2f} {100 * b[1] / tot_i:7.2f} {b[2] / max(b[1], 1):6.1f}")
