"""profiles/<tag>_ncu_summary.csv (tools/ncu_summary.py) -> profiles/r02_ncu_traffic.json: per kernel class, the DRAM bytes per launch and the
tensor-pipe activity bench.py quotes in `roofline.traffic` / `ncu_tensor_pipe_active_pct`.
usage: python tools/ncu_traffic.py profiles/r02_final_ncu_summary.csv "<how it was captured>" > profiles/r02_ncu_traffic.json"""
import csv
import json
import sys

rows = list(csv.DictReader(open(sys.argv[1])))


def col(prefix):
    return [k for k in rows[0] if k.startswith(prefix)][0]


T, RD, WR = col("gpu__time_duration.sum"), col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
TP, DT = col("sm__pipe_tensor_cycles_active"), col("gpu__dram_throughput")
unit = 1e6 if "Mbyte" in RD else (1e3 if "Kbyte" in RD else 1.0)
out = {"source": f"{sys.argv[1]} ({sys.argv[2]})"}


def sel(pred):
    return [r for r in rows if pred(r["kernel"])]


g = sel(lambda k: "gemm_tc_kernel" in k or "conv_pair_kernel" in k)      # the GEMM class: engine + CTA-pair convolutions
tt = sum(float(r[T]) for r in g)
out["gemm_tc_kernel"] = {"launches": len(g), "avg_dram_bytes_per_launch": sum((float(r[RD]) + float(r[WR])) * unit for r in g) / len(g),
                         "time_weighted_tensor_pipe_active_pct": sum(float(r[TP]) * float(r[T]) for r in g) / tt, "sum_time_us": tt}
a = sel(lambda k: "attn_fwd_kernel" in k)
a = sorted(a, key=lambda r: -float(r[T]))[:4]                  # the four DiT-block launches (the TV / linear-attention ones are shorter)
out["attn_fwd_kernel(dit)"] = {"launches": len(a), "avg_dram_bytes_per_launch": sum((float(r[RD]) + float(r[WR])) * unit for r in a) / len(a),
                               "tensor_pipe_active_pct": sum(float(r[TP]) for r in a) / len(a),
                               "avg_time_us": sum(float(r[T]) for r in a) / len(a)}
p = sel(lambda k: "posconv_kernel" in k)
if p:
    out["posconv_kernel"] = {"launches": len(p), "tensor_pipe_active_pct": sum(float(r[TP]) for r in p) / len(p),
                             "avg_time_us": sum(float(r[T]) for r in p) / len(p)}
n = sel(lambda k: "k_gn_apply" in k)
if n:
    out["k_gn_apply"] = {"launches": len(n), "sum_time_us": sum(float(r[T]) for r in n),
                         "avg_dram_bytes_per_launch": sum((float(r[RD]) + float(r[WR])) * unit for r in n) / len(n),
                         "avg_dram_throughput_pct": sum(float(r[DT]) for r in n) / len(n)}
json.dump(out, sys.stdout, indent=1)
