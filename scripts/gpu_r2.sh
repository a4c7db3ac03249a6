#!/bin/bash
# One GPU-box visit of round 2.  usage: scripts/gpu_r2.sh <tag> [steps...]
# steps: newtests tests bench ref launches ncu_wave pipe2 bench2 (default: newtests tests bench)
set -u
TAG=${1:-r2}; shift || true
STEPS=${*:-newtests tests bench}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
has() { [[ " $STEPS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
if has newtests; then
    timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_sim_golden.py -m gpu -x -q > "$OUT/pytest_new.log" 2>&1; echo "newtests rc=$?" | tee -a "$OUT/pytest_new.log"
    tail -25 "$OUT/pytest_new.log"
fi
if has tests; then
    timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
    tail -15 "$OUT/pytest_gpu.log"
fi
if has bench; then
    timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"
    cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
fi
if has ref; then
    timeout 900 python bench.py --impl reference --steps 5 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "ref rc=$?"
    cat "$OUT/bench_ref.json"; tail -3 "$OUT/bench_ref.err"
fi
if has launches; then
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 900 --csv --log-file "$OUT/launches.csv" \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$OUT/launches.log" 2>&1
    python scripts/launch_shares.py "$OUT/launches.csv" 30 > "$OUT/launch_shares.txt" 2>&1; head -30 "$OUT/launch_shares.txt"
fi
if has ncu_wave; then
    timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wave_ -c 3 -f -o "$OUT/wave" \
        python scripts/mode_compare.py 3 1.0 > "$OUT/ncu_wave.log" 2>&1
fi
if has bench2; then
    for n in 2; do
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 \
            > "$OUT/bench_n$n.json" 2> "$OUT/bench_n$n.err"; echo "bench n=$n rc=$?"
        cat "$OUT/bench_n$n.json"; tail -5 "$OUT/bench_n$n.err"
    done
fi
ls -la "$OUT"
