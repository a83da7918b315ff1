#!/bin/bash
# N = 2 job: the peer-memory exchange between two real GPUs + the multi-GPU bench lines (p2p vs NCCL gather).
mkdir -p gpurun_out/r2x
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/multi_gpu_exchange.py > gpurun_out/r2x/multi_exchange.log 2>&1; echo "multi_exchange rc=$?"
tail -3 gpurun_out/r2x/multi_exchange.log
FKV_PEER_BACKEND=ipc timeout 200 $TR --master-port 29512 tests/multi_gpu_exchange.py > gpurun_out/r2x/multi_exchange_ipc.log 2>&1; echo "multi_exchange ipc rc=$?"
tail -2 gpurun_out/r2x/multi_exchange_ipc.log
for ex in auto gather; do
  timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --exchange $ex --no-cpu-baseline --no-e2e > gpurun_out/r2x/bench_n2_$ex.json 2> gpurun_out/r2x/bench_n2_$ex.err; echo "bench n2 $ex rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2x/bench_n2_$ex.json").read().strip().splitlines()[-1])
    print("$ex", d["ms_per_step"], d["value"], d["config"].get("exchange"), d.get("timed_with"))
except Exception as e:
    print("$ex: no line", e)
PY
done
for ex in auto gather; do
  timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong --ctx 131072 --exchange $ex --no-cpu-baseline --no-e2e > gpurun_out/r2x/bench_n2_strong128k_$ex.json 2> gpurun_out/r2x/bench_n2_strong128k_$ex.err; echo "bench n2 strong $ex rc=$?"
  tail -c 600 gpurun_out/r2x/bench_n2_strong128k_$ex.json | head -c 300; echo
done
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2x/bench_n1_ref.json 2> gpurun_out/r2x/bench_n1_ref.err; echo "bench n1 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2x/bench_n1_ref.json').read().strip().splitlines()[-1]); print('n1', d['ms_per_step'], d['value'])"
