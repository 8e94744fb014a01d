#!/bin/bash
# round 2, first GPU check of the span rasteriser: parity tests, then per-mode timings
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1
tail -5 gpurun_out/r2_pytest1.log
for args in "--config 2 --reps 8" "--config 2 --reps 8 --nocull" "--config 2 --reps 8 --parents identity" "--config 3 --reps 4" "--config 5 --reps 4"; do
  timeout 300 python tools/run_modes.py $args 2>&1 | tail -1 | tee -a gpurun_out/r2_modes1.jsonl
done
