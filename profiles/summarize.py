#!/usr/bin/env python
"""Summarise ncu captures (run here, no GPU needed):
     python profiles/summarize.py <rep.ncu-rep | launches.csv> [...] > profiles/<name>.md
A .csv argument is a launch list (gpu__time_duration.sum per launch)."""
import collections
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction (of 32)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("lts__t_bytes.sum.per_second", "L2 throughput"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed_op_shared_ld.sum", "LDS instructions"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads"), ("smsp__inst_executed_op_local_st.sum", "local stores"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (per issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}
    print(f"## {path}\n")
    kn = col.get("Kernel Name")
    for r in rows[2:]:
        print(f"### launch id {r[col['ID']]}: `{r[kn][:90]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for key, label in KEYS:
            if key in col and r[col[key]] != "":
                print(f"| {label} (`{key}`) | {r[col[key]]} | {units[col[key]]} |")
        print()


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) > mv:
            name = r[kn].split("(")[0]
            agg[name][0] += 1
            agg[name][1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"## {path} — launch list (gpu__time_duration.sum; cold-cache, serialised: compare SHARES)\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.1f} | {v[1] / tot:.3f} |")
    print()


for a in sys.argv[1:]:
    launches(a) if a.endswith(".csv") else rep(a)
