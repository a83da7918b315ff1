#!/bin/bash
# final validation of the round-2 state on one B200: full GPU test suite, smoke, sanitizers, default bench, launch list
mkdir -p gpurun_out/r2g
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2g/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2g/smoke.txt
for tool in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_run.py > gpurun_out/r2g/sanitizer_$tool.log 2>&1
  echo "$tool: $(grep -c 'ERROR SUMMARY: 0 errors' gpurun_out/r2g/sanitizer_$tool.log) $(tail -1 gpurun_out/r2g/sanitizer_$tool.log)"
done
timeout 900 python bench.py > gpurun_out/r2g/bench_full.json 2> gpurun_out/r2g/bench_full.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2g/bench_full.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "roofline", d["roofline"]["frac"], d["roofline"].get("us_per_launch"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"],
      "cpu", d["cpu_baseline"]["value"], "launches", d["launches_per_step"])
for r in d.get("context_sweep") or []:
    print("  ctx", r)
print("  k5", d.get("decode_attention"))
print("  model", d.get("model"))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2g/bench_reference_arm.json 2>/dev/null; tail -c 400 gpurun_out/r2g/bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g/launches_bench_32k.csv python bench.py --steps 2 --warmup 3 --no-overlap --no-cpu-baseline --no-e2e --no-sweep --no-model > gpurun_out/r2g/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
