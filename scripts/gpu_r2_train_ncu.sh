#!/bin/bash
mkdir -p gpurun_out
PN_TIME_ONLY=rays timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_train_launches.csv python scripts/time_training.py > gpurun_out/r2_train_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2_train_ncu.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_train_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:70]),{})[r[mi]]=r[vi]
for (i,k),m in d.items():
    if "train" in k or "march" in k or "composite" in k:
        print(i,k,m)
PY
