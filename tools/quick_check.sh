#!/bin/bash
# quick GPU check of a kernel change: parity tests, then per-mode timings of config 2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/run_modes.py --config 2 --reps 8 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('config 2 eval / update ms', d['loglikes']['avg_ms'], d['update']['avg_ms'])"
