"""Markdown summary of a round's ncu evidence.
  python tests/tools/profile_summary.py <launches.csv> <raw_kernelA.csv> [<raw_kernelB.csv> ...]
launches.csv : ncu --metrics gpu__time_duration.sum --csv log of `python bench.py --steps K --warmup W`
raw_*.csv    : ncu -i <rep> --page raw --csv of a --set full capture (one launch each)"""
import csv
import re
import sys
from collections import OrderedDict


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    names = [re.sub(r"\(.*", "", r[4]).replace("void ", "") for r in rows]
    us = [float(r[-1]) / 1e3 for r in rows]
    starts = [i for i, n in enumerate(names) if n.startswith("hbpp_stage2_kernel<0") or n.startswith("hbpp_stage_kernel<0")]
    print(f"## Launch list ({len(rows)} launches; {len(starts)} iterations seen)\n")
    if len(starts) < 2:
        return
    lo, hi = starts[-2], starts[-1]          # the last complete iteration
    it = OrderedDict()
    for n, t in zip(names[lo:hi], us[lo:hi]):
        if n.startswith("at::"):
            n = "(torch fill: the bench's L2 flush, not part of the step)"
        it.setdefault(n, [0, 0.0])
        it[n][0] += 1
        it[n][1] += t
    tot = sum(v[1] for k, v in it.items() if not k.startswith("("))
    print("Last complete iteration (times under ncu are cold-cache and serialised: the SHARE is what counts):\n")
    print("| kernel | launches | us | share |\n|---|---|---|---|")
    for k, (c, t) in it.items():
        share = "" if k.startswith("(") else f"{100 * t / tot:.1f} %"
        print(f"| `{k}` | {c} | {t:.1f} | {share} |")
    print(f"| **step total (repo kernels)** | {sum(v[0] for k, v in it.items() if not k.startswith('('))} | {tot:.1f} | |\n")


WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__shared_mem_per_block_static", "static smem/block"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM written"), ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "barrier stall / issue"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long-scoreboard stall / issue"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "no-instruction stall / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "wait stall / issue"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate"), ("sass__inst_executed_register_spilling", "spill instructions")]


def full(path):
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        print(f"## `--set full`: `{v[h.index('Kernel Name')]}`\n\n| metric | value |\n|---|---|")
        for key, label in WANT:
            if key in h:
                i = h.index(key)
                print(f"| {label} (`{key}`) | {v[i]} {u[i]} |")
        print()


if __name__ == "__main__":
    launches(sys.argv[1])
    for p in sys.argv[2:]:
        full(p)
