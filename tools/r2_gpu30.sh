#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest30.log 2>&1
tail -15 gpurun_out/r2_pytest30.log
rm -f gpurun_out/r2_modes30.jsonl
for args in "--config 2 --reps 8" "--config 2 --reps 8 --parents few" "--config 3 --reps 4" "--config 5 --reps 4"; do
  timeout 200 python tools/run_modes.py $args 2>&1 | tail -1 | tee -a gpurun_out/r2_modes30.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config'], d['parents'], d['loglikes']['avg_ms'], d['update']['avg_ms'])"
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench30.json 2> gpurun_out/r2_bench30.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench30.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["filter"]["frames_per_s"], d["roofline"]["plan_kernel_avg_ms"], d["roofline"]["avg_launch_ms"])
PY
