"""Summarise an .ncu-rep (one kernel launch, --set full) into a small text file for profiles/.
usage: ncu_summary.py report.ncu-rep n_events out.txt"""
import csv, io, subprocess, sys, collections, re

rep, n_events, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = """Kernel Name
gpu__time_duration.sum
sm__cycles_elapsed.avg.per_second
launch__grid_size
launch__block_size
launch__registers_per_thread
launch__shared_mem_per_block_static
launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem
sm__warps_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum
smsp__thread_inst_executed_per_inst_executed.ratio
smsp__issue_active.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
dram__bytes_read.sum
dram__bytes_write.sum
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio""".split("\n")
lines = [f"# ncu summary of {rep.split('/')[-1]} ({n_events:.4g} events in the profiled launch)"]
for k in keys:
    if k in m:
        lines.append(f"{k} = {m[k][0]} {m[k][1]}")
t_ms = float(m["gpu__time_duration.sum"][0]) * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[m["gpu__time_duration.sum"][1]]
lines.append(f"events/s under ncu (not a bench value) = {n_events / (t_ms * 1e-3):.4g}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
iS, iI, iT = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
by = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) <= iT or not r[iI]:
        continue
    s = re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip())
    op = s.split()[0] if s else "?"
    fam = op.split(".")[0]
    key = op if fam in ("MUFU", "I2F", "F2I", "F2F") else fam
    by[key] += int(r[iI]); tot += int(r[iI])
fp64 = sum(n for k, n in by.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
lines.append(f"warp instructions = {tot}; lane-slots per event = {tot * 32 / n_events:.1f}; FP64-pipe share of issue slots = {100 * fp64 / tot:.1f}%")
lines.append("opcode mix (lane-slots per event, % of issued):")
for k, n in by.most_common(28):
    lines.append(f"  {k:14} {n * 32 / n_events:8.1f} {100 * n / tot:5.1f}%")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
