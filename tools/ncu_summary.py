"""Condense `ncu -i X.ncu-rep --page raw --csv` into one row per kernel with the metrics DESIGN.md / profiles cite.
usage: ncu -i rep --page raw --csv | python tools/ncu_summary.py > summary.csv"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_umma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
cols = [w for w in WANT if w in hdr] + [h for h in hdr if "shared" in h and "throughput" in h][:4] + stall
out = csv.writer(sys.stdout)
out.writerow(["id", "kernel"] + [f"{c} [{units[hdr.index(c)]}]" for c in cols])
for r in data:
    if len(r) < len(hdr):
        continue
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void dexb::", "")
    out.writerow([r[0], name] + [r[hdr.index(c)] for c in cols])
