#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_training.py tests/test_gpu_ext.py -q --timeout 300 > gpurun_out/r2_train_ff.log 2>&1; tail -6 gpurun_out/r2_train_ff.log | cut -c1-300
for w in 0 1; do
  PN_TRAIN_WRITE=$w PN_TIME_ONLY=rays_timed timeout 200 python scripts/time_training.py 2>&1 | grep -E "march_rays_train_800x800_ms"
done
