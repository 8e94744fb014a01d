#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest29.log 2>&1
tail -15 gpurun_out/r2_pytest29.log
rm -f gpurun_out/r2_modes29.jsonl
for args in "--config 2 --reps 8" "--config 2 --reps 8 --parents few" "--config 3 --reps 4" "--config 5 --reps 4"; do
  timeout 200 python tools/run_modes.py $args 2>&1 | tail -1 | tee -a gpurun_out/r2_modes29.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config'], d['parents'], d['loglikes']['avg_ms'], d['update']['avg_ms'])"
done
