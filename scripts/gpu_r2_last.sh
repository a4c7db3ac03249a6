#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_training.py tests/test_gpu_ext.py tests/test_gpu_encoders.py -q --timeout 150 2>&1 | tail -3 | cut -c1-200
timeout 60 python scripts/time_training.py gpurun_out/r2_train_timing_final.json 2>&1 | grep -E "sh_forward|march_rays_train_800|composite"
