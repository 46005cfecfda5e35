"""Summarise an ncu CSV export (details + raw pages): python tools/ncu_summary.py gpurun_out/TAG/KERNEL_WL"""
import csv, sys
base = sys.argv[1]
rows = list(csv.reader(open(base + ".details.csv")))
hdr = rows[0]
want = ["Duration", "Elapsed Cycles", "Memory Throughput", "DRAM Throughput", "L1/TEX Cache Throughput", "L2 Cache Throughput",
        "Compute (SM) Throughput", "Issue Slots Busy", "Executed Ipc Active", "L1/TEX Hit Rate", "L2 Hit Rate", "Mem Pipes Busy",
        "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy", "Active Warps Per Scheduler",
        "Eligible Warps Per Scheduler", "Executed Instructions", "Dynamic Shared Memory Per Block", "Grid Size", "Block Size"]
for r in rows[1:]:
    d = dict(zip(hdr, r))
    if d.get("Metric Name") in want:
        print(f"{d['Metric Name'][:40]:40s} {d['Metric Unit'][:14]:14s} {d['Metric Value']}")
rows = list(csv.reader(open(base + ".raw.csv")))
hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
d = dict(zip(hdr, vals))
def f(k):
    try: return float(d[k].replace(",", ""))
    except Exception: return float("nan")
st = sorted(((f(k), k) for k in d if "issue_stalled" in k and k.endswith("per_issue_active.ratio")), reverse=True)
print("stalls:", ", ".join(f"{k.split('issue_stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, k in st[:7]))
for k in ["dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed",
          "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum",
          "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active"]:
    if k in d: print(k, d[k], hdr and "")
for k in ["smsp__inst_executed.sum", "gpu__time_duration.sum"]:
    if k in d: print(k, d[k])
# the element count of the captured launch, printed by tools/profile_workload.py into the ncu log
import os, re
log = os.path.join(os.path.dirname(base), "ncu_" + os.path.basename(base).rsplit("_", 1)[0] + ".log")
if os.path.exists(log):
    m = re.search(r"launch_matrix_elements (\d+)", open(log, errors="replace").read())
    if m: print("launch_matrix_elements", m.group(1))
