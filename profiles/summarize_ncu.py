"""Summarise an `ncu --set full` report for profiles/: the roofline-relevant raw metrics plus the
SASS opcode mix.  Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rN/x.md"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_write.sum.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    print(f"# ncu --set full summary of `{rep}`\n")
    for k, row in enumerate(raw[2:]):
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print(f"## launch {k}: `{d.get('Kernel Name', '?')}`  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for key in KEYS:
            if key in d:
                print(f"| {key} | {d[key]} | {u[key]} |")
        print()
    src = page(rep, "source")
    h = src[1]
    iS, iI, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    mix, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in src[2:]:
        if len(r) <= iI or r[0] == "Address" or not r[iI].replace(".", "").isdigit():
            continue
        toks = r[iS].split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        if not toks:
            continue
        op = toks[0].split(".")[0]
        n = int(float(r[iI] or 0))
        mix[op] += n
        samp[op] += int(float(r[iN] or 0))
        tot += n
    ts = max(sum(samp.values()), 1)
    print(f"## SASS opcode mix (first kernel in the report): {tot:,} warp-instructions, {ts:,} stall samples\n")
    print("| opcode | warp-instructions | share | stall samples |\n|---|---|---|---|")
    for op, n in mix.most_common(24):
        print(f"| {op} | {n:,} | {100 * n / tot:.2f}% | {100 * samp[op] / ts:.2f}% |")


if __name__ == "__main__":
    main()
