#!/bin/bash
# Strong-scaling evidence on one 8-GPU box: chair at N = 8 and 4 (N = 1, 2 are measured on smaller boxes), 1080p config at N = 8, and the
# per-rank pipeline timeline at N = 8.  Everything lands in gpurun_out/r2s/.
set -u
OUT=gpurun_out/r2s; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() {  # name nproc extra-args...
    local name=$1 n=$2; shift 2
    PN_BENCH_WATCHDOG_S=150 timeout 200 $TR --nproc-per-node $n --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n --steps 30 --warmup 3 --quick "$@" > $OUT/$name.json 2> $OUT/$name.err
    python - <<PY
import json
try:
    d = json.load(open("$OUT/$name.json"))
    print("$name", round(d["value"], 1), round(d["e2e"]["value"], 1), d["frame_checksum"]["sha1_image"][:8], d["details"]["parallelism"][-160:])
except Exception as e:
    print("$name failed", e); print(open("$OUT/$name.err").read()[-1500:])
PY
}
run chair_n8 8
run chair_n4 4
run synth1080_n8 8 --config synth1080 --steps 15
timeout 120 $TR --nproc-per-node 8 --master-port 29811 scripts/pipe_timeline.py chair 9 3 > $OUT/timeline8.txt 2>&1
grep -E "^rank 0|^frame  [5-8]" $OUT/timeline8.txt | head -8
