#!/bin/bash
# the driver's own command line for N = 1 with few steps: does the line come out and parse?
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["repeats"]["count"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["gpu_launches"], d["roofline"]["frac_dram"])
PY
tail -2 gpurun_out/bench_quick.err
