#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:"k_la_weff|k_la_combine|k_attn_tail_merge|k_dw_patch|k_tok_assemble|k_unpatchify|k_chan_stats|k_freq_mean|k_tv_fold|k_ln_mod|k_posconv_pack_in|k_fill_zero|k_gn_final|k_conv_in" -s 40 -c 40 --csv --log-file gpurun_out/r02z_small.csv python tools/prof_net_call.py C2 3 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02z_small.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg={}
for r in rows[1:]:
    n=r[ki].split("(")[0][-30:]; v=float(r[vi].replace(",",""))*{"ns":1e-3,"us":1,"ms":1e3}.get(r[ui],1)
    agg.setdefault(n,[]).append(v)
for n,v in agg.items(): print(n.ljust(32), len(v), " ".join(f"{x:.1f}" for x in v[:8]))
PY
