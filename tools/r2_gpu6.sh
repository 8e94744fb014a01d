#!/bin/bash
# round 2: free-running copy stream, natural order, occupancy variants
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest6.log 2>&1
tail -4 gpurun_out/r2_pytest6.log
rm -f gpurun_out/r2_modes6.jsonl
for args in "--config 2 --reps 8" "--config 2 --reps 8 --parents identity" "--config 2 --reps 8 --parents few" "--config 3 --reps 4" "--config 3 --reps 4 --parents few" "--config 5 --reps 4"; do
  timeout 300 python tools/run_modes.py $args 2>&1 | tail -1 | tee -a gpurun_out/r2_modes6.jsonl
done
for v in cfirst cw20 cw24 cw28; do
  echo "variant $v" | tee -a gpurun_out/r2_modes6.jsonl
  DBOT_B200_LIB=build/variants/lib_$v.so timeout 300 python tools/run_modes.py --config 2 --reps 8 2>&1 | tail -1 | tee -a gpurun_out/r2_modes6.jsonl
  DBOT_B200_LIB=build/variants/lib_$v.so timeout 300 python tools/run_modes.py --config 3 --reps 4 2>&1 | tail -1 | tee -a gpurun_out/r2_modes6.jsonl
  DBOT_B200_LIB=build/variants/lib_$v.so timeout 300 python tools/run_modes.py --config 5 --reps 4 2>&1 | tail -1 | tee -a gpurun_out/r2_modes6.jsonl
done
DBOT_B200_LIB=build/variants/lib_timeline.so timeout 300 python tools/run_modes.py --config 2 --reps 1 2>&1 | grep -E "^mode" | head -12 > gpurun_out/r2_timeline6.txt
