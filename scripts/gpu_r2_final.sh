#!/bin/bash
# Final visit of round 2: whole GPU suite, training-kernel timing, a quick headline bench (no microbenches / CPU baseline).
mkdir -p gpurun_out/r2z
timeout 600 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/r2z/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2z/pytest_gpu.log
tail -6 gpurun_out/r2z/pytest_gpu.log | cut -c1-250
timeout 120 python scripts/time_training.py gpurun_out/r2z/train_timing.json > gpurun_out/r2z/train_timing.log 2>&1; echo "timing rc=$?"
timeout 300 python bench.py --quick > gpurun_out/r2z/bench_quick.json 2> gpurun_out/r2z/bench_quick.err; echo "bench rc=$?"
cat gpurun_out/r2z/bench_quick.json | cut -c1-900
cat gpurun_out/r2z/train_timing.json
