#!/bin/bash
# the two arms of bench.py with the final tree (profiles/traffic.json from the same kernels) + smoke
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2_bench35.json 2> gpurun_out/r2_bench35.err
timeout 300 python bench.py --impl reference --steps 20 > gpurun_out/r2_bench35_reference.json 2>> gpurun_out/r2_bench35.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench35.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["filter"]["frames_per_s"], d["roofline"]["frac"], d["roofline"]["frac_dram"], d["cpu_baseline"]["value"])
r=json.loads(open("gpurun_out/r2_bench35_reference.json").read().strip().splitlines()[-1]); print(r["value"])
PY
