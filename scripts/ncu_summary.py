"""Summarise an .ncu-rep (ncu --set full --import-source on) as markdown: selected raw metrics per launch and the
SASS opcode mix of the first kernel.  usage: python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep > profiles/.../X.md"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
WANT = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__occupancy_limit_registers sm__warps_active.avg.pct_of_peak_sustained_active dram__bytes_read.sum
dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed dram__bytes_write.sum.per_second
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
smsp__sass_thread_inst_executed_op_dadd_pred_on.sum sass__inst_executed_local_loads sass__inst_executed_local_stores
lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio""".split()

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
print(f"# ncu --set full summary of `{rep}`\n")
for k, r in enumerate(data):
    name = r[col["Kernel Name"]]
    print(f"## launch {k}: `{name}`  grid {r[col['Grid Size']]} block {r[col['Block Size']]}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for m in WANT:
        if m in col:
            print(f"| {m} | {r[col[m]]} | {units[col[m]]} |")
    print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-kernel-base", "function"],
                     capture_output=True, text=True).stdout
lines = [l for l in src.splitlines() if l.startswith('"') and not l.startswith('"Kernel Name"')]
if lines:
    rd = list(csv.reader(io.StringIO("\n".join(lines))))
    h = rd[0]
    ci = {n: i for i, n in enumerate(h)}
    s_col = ci.get("Source")
    e_col = ci.get("Instructions Executed")
    st_col = ci.get("Warp Stall Sampling (All Samples)", ci.get("Warp Stall Sampling (All Cycles)"))
    ops, stalls = collections.Counter(), collections.Counter()
    for r in rd[1:]:
        if len(r) <= max(s_col, e_col):
            continue
        ins = re.sub(r"^@!?U?P\w+\s+", "", r[s_col].strip())
        op = ins.split()[0].split(".")[0] if ins else "?"
        try:
            ops[op] += int(float(r[e_col]))
            if st_col is not None:
                stalls[op] += int(float(r[st_col]))
        except ValueError:
            pass
    tot, stot = sum(ops.values()), max(1, sum(stalls.values()))
    print(f"## SASS opcode mix (first kernel in the report): {tot:,} warp-instructions, {sum(stalls.values()):,} stall samples\n")
    print("| opcode | warp-instructions | share | stall samples |\n|---|---|---|---|")
    for op, n in ops.most_common(24):
        print(f"| {op} | {n:,} | {100.0 * n / tot:.2f}% | {100.0 * stalls[op] / stot:.2f}% |")
