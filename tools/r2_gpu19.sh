#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest19.log 2>&1
tail -3 gpurun_out/r2_pytest19.log
rm -f gpurun_out/r2_modes19.jsonl
for args in "--config 2 --reps 8" "--config 3 --reps 4" "--config 5 --reps 4"; do
  timeout 200 python tools/run_modes.py $args 2>&1 | tail -1 | tee -a gpurun_out/r2_modes19.jsonl | cut -c1-400
done
DBOT_B200_LIB=build/variants/lib_timeline.so timeout 200 python tools/run_modes.py --config 2 --reps 1 2>&1 | grep -E "^mode" | head -8 > gpurun_out/r2_timeline19.txt
cat gpurun_out/r2_timeline19.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench19.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["filter"])
PY
