#!/bin/bash
# N GPUs: weak-scaling bench with the peer-memory exchange + tensor-parallel model TTFT (128k) with FKV_EXCHANGE=p2p
N=${1:-4}
mkdir -p gpurun_out/r2x
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2x/bench_n${N}_weak_p2p.json 2> gpurun_out/r2x/bench_n${N}_weak_p2p.err; echo "bench weak rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2x/bench_n${N}_weak_p2p.json').read().strip().splitlines()[-1]); print('weak', d['ms_per_step'], d['value'], d['config']['exchange'])"
for ex in p2p gather; do
FKV_EXCHANGE=$ex timeout 400 $TR --master-port 29516 tools/model_bench.py --ctx 131072 --genlen 16 --runs 2 > gpurun_out/r2x/model_llama_128k_tp${N}_$ex.json 2> gpurun_out/r2x/model_tp${N}_$ex.err; echo "model tp$N $ex rc=$?"
tail -1 gpurun_out/r2x/model_llama_128k_tp${N}_$ex.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['ttft_ms'], d['decode_tok_s'])"
done
