#!/bin/bash
# Round-2 evidence run on ONE GPU: full bench, reference arm, ncu launch list, ncu --set full of the wavefront kernels (default build
# and the level-0-TMA-staging variant), march/field timing of both.  Everything lands in gpurun_out/r2p/.
set -u
OUT=gpurun_out/r2p; mkdir -p $OUT
PN_BENCH_WATCHDOG_S=400 timeout 450 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
PN_BENCH_WATCHDOG_S=200 timeout 250 python bench.py --impl reference --steps 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --graph-profiling node -c 1200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --quick > $OUT/launches.log 2>&1
python scripts/launch_shares.py $OUT/launches.csv 30 > $OUT/launch_shares.txt 2>&1; head -12 $OUT/launch_shares.txt
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wave_ -c 3 -f -o $OUT/wave \
    python scripts/mode_compare.py 3 1.0 > $OUT/ncu_wave.log 2>&1; tail -1 $OUT/ncu_wave.log | cut -c1-160
PN_LIB=$PWD/pienerf_b200/lib/libpienerf_b200_lvl0.so timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wave_field -c 1 -f -o $OUT/wave_lvl0 \
    python scripts/mode_compare.py 3 1.0 > $OUT/ncu_wave_lvl0.log 2>&1; tail -1 $OUT/ncu_wave_lvl0.log | cut -c1-160
echo "== timing default vs lvl0 (no profiler)"
timeout 100 python scripts/mode_compare.py 3 1.0 2>&1 | tail -1 | cut -c1-100
PN_LIB=$PWD/pienerf_b200/lib/libpienerf_b200_lvl0.so timeout 100 python scripts/mode_compare.py 3 1.0 2>&1 | tail -1 | cut -c1-100
ls -la $OUT
