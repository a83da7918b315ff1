set -x
mkdir -p gpurun_out/r2
timeout 900 python bench.py --sweep gpurun_out/r2/sweep.md > gpurun_out/r2/sweep.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_full.json 2> gpurun_out/r2/bench_full.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2/bench_ref.json 2>&1
timeout 400 python tools/model_bench.py --model ministral --batch 8 --ctx 32768 --capacity 512 --genlen 128 --graph --runs 1 > gpurun_out/r2/model_ministral_b8.json 2>&1
timeout 400 python tools/model_bench.py --ctx 131072 --genlen 256 --graph --runs 1 > gpurun_out/r2/model_llama_128k.json 2>&1
timeout 300 python tools/model_bench.py --ctx 131072 --genlen 32 --runs 1 > gpurun_out/r2/model_llama_128k_eager.json 2>&1
for tool in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_run.py > gpurun_out/r2/sanitizer_$tool.log 2>&1
  tail -3 gpurun_out/r2/sanitizer_$tool.log
done
ls -la gpurun_out/r2
