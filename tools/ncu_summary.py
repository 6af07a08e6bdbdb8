"""Summarise an .ncu-rep (first profiled kernel): key counters + stall/opcode breakdown.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
def P(*a): print(*a, file=out)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
P("kernel:", m.get("Kernel Name", ("?",))[0], " grid", m.get("Grid Size", ("?",))[0], " block", m.get("Block Size", ("?",))[0])
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]
for k in keys:
    if k in m: P(f"  {k:78s} {m[k][0]:>16s} {m[k][1]}")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        v = float(m[h][0])
        if v >= 0.1: P(f"  stall {h[34:-23]:40s} {v:6.2f} warps/issue")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
i_src, i_s, i_ex = h2.index("Source"), h2.index("Warp Stall Sampling (All Samples)"), h2.index("Instructions Executed")
by, byex = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= i_ex: continue
    op = re.sub(r"^@!?U?P\d+\s+", "", r[i_src].strip()).split()[0].split(".")[0]
    by[op] += int(r[i_s] or 0); byex[op] += int(r[i_ex] or 0)
tot, totex = sum(by.values()), sum(byex.values())
P(f"  opcode mix (warp instructions executed, total {totex}) and stall samples (total {tot}):")
for op, e in byex.most_common(16):
    P(f"    {op:8s} executed {100*e/totex:5.1f}%   samples {100*by[op]/max(tot,1):5.1f}%")
