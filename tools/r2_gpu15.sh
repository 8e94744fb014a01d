#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest15.log 2>&1
tail -3 gpurun_out/r2_pytest15.log
rm -f gpurun_out/r2_modes15.jsonl
for args in "--config 2 --reps 8" "--config 3 --reps 4" "--config 5 --reps 4"; do
  timeout 200 python tools/run_modes.py $args 2>&1 | tail -1 | tee -a gpurun_out/r2_modes15.jsonl
done
timeout 600 python bench.py > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err
tail -c 300 gpurun_out/r2_bench15.json
