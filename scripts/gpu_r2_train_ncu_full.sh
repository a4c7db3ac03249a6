#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"grid_backward_kernel|train_count_kernel|train_write_kernel|train_composite_bwd" -c 4 -o gpurun_out/r2_train_full -f python scripts/time_training.py > gpurun_out/r2_train_full.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/r2_train_full.ncu-rep --page raw --csv > gpurun_out/r2_train_full_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r2_train_full_raw.csv")))
hdr=rows[0]
want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_sectors.sum","lts__t_sectors_op_red.sum","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_active","sm__throughput.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","dram__throughput.avg.pct_of_peak_sustained_elapsed","smsp__thread_inst_executed_per_inst_executed.ratio"]
idx=[hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print({hdr[i]:r[i][:60] for i in idx})
PY
