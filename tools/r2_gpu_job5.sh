#!/bin/bash
# N = 8 job: peer-memory exchange at world 8 (correctness, timeline) + weak / strong bench lines.
mkdir -p gpurun_out/r2x
cd "$(dirname "$0")/.."
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29511 tests/multi_gpu_exchange.py > gpurun_out/r2x/multi_exchange_n$N.log 2>&1; echo "multi_exchange rc=$?"
tail -2 gpurun_out/r2x/multi_exchange_n$N.log
timeout 240 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2x/bench_n${N}_weak_p2p.json 2> gpurun_out/r2x/bench_n${N}_weak_p2p.err; echo "bench weak rc=$?"
timeout 240 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --scaling strong --ctx 131072 --no-cpu-baseline --no-e2e > gpurun_out/r2x/bench_n${N}_strong128k_p2p.json 2> gpurun_out/r2x/bench_n${N}_strong128k_p2p.err; echo "bench strong rc=$?"
timeout 240 $TR --master-port 29515 bench.py --gpus $N --steps 20 --warmup 3 --scaling strong --no-cpu-baseline --no-e2e > gpurun_out/r2x/bench_n${N}_strong32k_p2p.json 2> gpurun_out/r2x/bench_n${N}_strong32k_p2p.err; echo "bench strong32k rc=$?"
python - <<PY
import json
for f in ("weak_p2p", "strong128k_p2p", "strong32k_p2p"):
    try:
        d = json.loads(open(f"gpurun_out/r2x/bench_n${N}_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["config"].get("exchange"), d["launches_per_step"])
    except Exception as e:
        print(f, "no line", e)
PY
