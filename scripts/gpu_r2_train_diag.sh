#!/bin/bash
mkdir -p gpurun_out
for s in 0 1; do
PN_TRAIN_SKIP=$s PN_TIME_ONLY=rays timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_issue_stalled_long_scoreboard.pct --clock-control none -k regex:train_count -c 2 --csv --log-file gpurun_out/r2_train_diag_$s.csv python scripts/time_training.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_train_diag_$s.csv")) if len(r)>10]
hdr=rows[0]; mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault(r[ii],{})[r[mi]]=r[vi]
print("skip=$s", d)
PY
done
