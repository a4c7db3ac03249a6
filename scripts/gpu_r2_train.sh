#!/bin/bash
# Training-side kernels on the B200 box: golden vectors from the reference's own kernels, parity tests, timing.
mkdir -p gpurun_out
timeout 300 python tests/golden/make_golden_train.py gpurun_out/ref_train.npz > gpurun_out/r2_train_golden.log 2>&1
echo "golden rc=$?" >> gpurun_out/r2_train_golden.log
cp -f gpurun_out/ref_train.npz tests/golden/ref_train.npz 2>/dev/null
timeout 600 python -m pytest tests/test_gpu_training.py tests/test_train_oracle.py tests/test_gpu_ext.py tests/test_gpu_encoders.py -q --timeout 300 > gpurun_out/r2_train_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_train_tests.log
timeout 200 python scripts/time_training.py gpurun_out/r2_train_timing.json > gpurun_out/r2_train_timing.log 2>&1
echo "timing rc=$?" >> gpurun_out/r2_train_timing.log
tail -5 gpurun_out/r2_train_golden.log; tail -30 gpurun_out/r2_train_tests.log; tail -40 gpurun_out/r2_train_timing.log
