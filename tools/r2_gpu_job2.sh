set -x
mkdir -p gpurun_out/r2
timeout 200 python tools/pipe_time.py > gpurun_out/r2/pipe_time.txt 2>&1; cat gpurun_out/r2/pipe_time.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q -x --timeout 600 2>&1 | tail -4
FKV_SELECT_TRACE=1 timeout 100 python tools/select_trace.py 32768 2048 1 2>&1 | tail -9 > gpurun_out/r2/select_trace_32k.txt
FKV_SELECT_TRACE=1 timeout 100 python tools/select_trace.py 131072 2048 1 2>&1 | tail -9 > gpurun_out/r2/select_trace_128k.txt
FKV_SELECT_TRACE=1 timeout 100 python tools/select_trace.py 2048 2048 1 2>&1 | tail -9 > gpurun_out/r2/select_trace_2k.txt
FKV_DEC_TRACE=1 timeout 100 python tools/prof_decode.py 1 2304 2304 3 2>&1 | grep "decode trace" | tail -7 > gpurun_out/r2/decode_trace_b1.txt
FKV_DEC_TRACE=1 timeout 100 python tools/prof_decode.py 8 640 768 3 2>&1 | grep "decode trace" | tail -7 > gpurun_out/r2/decode_trace_b8.txt
ncu --set full --import-source on --clock-control none -k regex:select_compact -s 2 -c 1 -o gpurun_out/r2/r2_select_32k python tools/select_trace.py 32768 2048 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:select_compact -s 2 -c 1 -o gpurun_out/r2/r2_select_128k python tools/select_trace.py 131072 2048 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:decode_attn -s 3 -c 1 -o gpurun_out/r2/r2_decode_b1 python tools/prof_decode.py 1 2304 2304 4 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:decode_attn -s 3 -c 1 -o gpurun_out/r2/r2_decode_b8 python tools/prof_decode.py 8 640 768 4 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:score_tc -s 6 -c 1 -o gpurun_out/r2/r2_score_32k python tools/prof_score.py 32768 10 tensor > /dev/null 2>&1
ncu --set full --clock-control none -k regex:score_tc -s 4 -c 1 -o gpurun_out/r2/r2_score_128k python tools/prof_score.py 131072 6 tensor > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 140 --csv --log-file gpurun_out/r2/r2_launches_bench_32k.csv python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-model --no-sweep --no-overlap > /dev/null 2>&1
ls -la gpurun_out/r2 | tail -20
