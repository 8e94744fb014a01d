#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest16.log 2>&1
tail -3 gpurun_out/r2_pytest16.log
rm -f gpurun_out/r2_modes16.jsonl
for args in "--config 2 --reps 8" "--config 3 --reps 4" "--config 5 --reps 4"; do
  timeout 200 python tools/run_modes.py $args 2>&1 | tail -1 | tee -a gpurun_out/r2_modes16.jsonl | cut -c1-400
done
DBOT_B200_LIB=build/variants/lib_timeline.so timeout 200 python tools/run_modes.py --config 2 --reps 1 2>&1 | grep -E "^mode" | head -8 > gpurun_out/r2_timeline16.txt
cat gpurun_out/r2_timeline16.txt
