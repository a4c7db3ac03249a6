#!/bin/bash
# One GPU-box visit: parity tests, bench (ours + reference arm), the ncu launch list of the bench command and one
# `--set full` capture of each hot kernel.  Everything lands in gpurun_out/<tag>/.   usage: scripts/gpu_round.sh <tag> [steps...]
# steps: tests bench ref launches ncu_frame ncu_grid ncu_sim ncu_field (default: all)
set -u
TAG=${1:-r1}; shift || true
STEPS=${*:-tests bench ref launches ncu_frame ncu_grid ncu_sim ncu_field}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
has() { [[ " $STEPS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
if has tests; then
    timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
    tail -5 "$OUT/pytest_gpu.log"
fi
if has bench; then
    timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"
    cat "$OUT/bench.json"
fi
if has ref; then
    timeout 900 python bench.py --impl reference --steps 5 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "ref rc=$?"
    cat "$OUT/bench_ref.json"
fi
if has launches; then
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file "$OUT/launches.csv" \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$OUT/launches.log" 2>&1
    python scripts/launch_shares.py "$OUT/launches.csv" 30 > "$OUT/launch_shares.txt" 2>&1; head -20 "$OUT/launch_shares.txt"
fi
if has ncu_frame; then
    # the three wavefront kernels of pass 0 (field = the roofline kernel), timed frames only
    timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wave_ -c 3 -f -o "$OUT/wave" \
        python scripts/mode_compare.py 3 1.0 > "$OUT/ncu_wave.log" 2>&1
fi
if has ncu_grid; then
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:grid_forward_d3c2 -s 3 -c 2 -f -o "$OUT/grid" \
        python scripts/grid_bench.py > "$OUT/ncu_grid.log" 2>&1
fi
if has ncu_sim; then
    timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ip_stress_kernel|rhs_partial_kernel|matvec3_kernel|ip_info_kernel" -s 40 -c 8 -f -o "$OUT/sim" \
        python scripts/profile_frame.py 4 1.0 0 > "$OUT/ncu_sim.log" 2>&1
fi
if has ncu_field; then
    timeout 900 ncu --set full --clock-control none --import-source on -k "regex:field_forward" -c 2 -f -o "$OUT/field" \
        python scripts/tc_debug.py 2097152 > "$OUT/ncu_field.log" 2>&1
fi
ls -la "$OUT"
