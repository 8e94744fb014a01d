#!/bin/bash
# quick GPU check of the current build: parity tests, then per-mode timings of configs 2 and 3
python -m pytest tests -m gpu -x -q > gpurun_out/quick_pytest.log 2>&1
python tools/run_modes.py --config 2 --reps 8 2>&1 | tail -1
python tools/run_modes.py --config 3 --reps 4 2>&1 | tail -1
tail -1 gpurun_out/quick_pytest.log
