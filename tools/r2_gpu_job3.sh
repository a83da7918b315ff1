set -x
mkdir -p gpurun_out/r2f
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/pipe_time.py > gpurun_out/r2f/pipe_time.txt 2>&1; cat gpurun_out/r2f/pipe_time.txt
timeout 900 python bench.py --sweep gpurun_out/r2f/sweep.md > gpurun_out/r2f/sweep.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f/bench_full.json 2> gpurun_out/r2f/bench_full.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-overlap --no-e2e --no-cpu-baseline --no-model --no-sweep > gpurun_out/r2f/bench_serial.json 2>&1
ncu --set full --clock-control none -k regex:score_tc -s 6 -c 1 -o gpurun_out/r2f/r2_score_32k python tools/prof_score.py 32768 10 tensor > /dev/null 2>&1
ncu --set full --clock-control none -k regex:score_tc -s 4 -c 1 -o gpurun_out/r2f/r2_score_128k python tools/prof_score.py 131072 6 tensor > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 140 --csv --log-file gpurun_out/r2f/r2_launches_bench_32k.csv python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-model --no-sweep --no-overlap > /dev/null 2>&1
timeout 400 python tools/model_bench.py --ctx 131072 --genlen 256 --graph --runs 1 > gpurun_out/r2f/model_llama_128k.json 2>&1
timeout 400 python tools/model_bench.py --model ministral --batch 8 --ctx 32768 --capacity 512 --genlen 128 --graph --runs 1 > gpurun_out/r2f/model_ministral_b8.json 2>&1
FKV_BUILD_TRACE=1 python fastkv_b200/_build.py >/dev/null 2>&1
FKV_TC_TRACE=1 timeout 100 python tools/prof_score.py 32768 6 tensor 2>&1 | tail -10 | head -9 > gpurun_out/r2f/score_trace_32k.txt
FKV_TC_TRACE=1 timeout 100 python tools/prof_score.py 131072 4 tensor 2>&1 | tail -10 | head -9 > gpurun_out/r2f/score_trace_128k.txt
ls gpurun_out/r2f
