#!/bin/bash
for n in 148 296 592 888 1000 1036 2072 4144; do python tools/run_modes.py --config 2 --n $n --reps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['n'], 'eval_ms=%.4f upd_ms=%.4f eval_us_per_wave=%.1f upd_us_per_wave=%.1f'%(d['loglikes']['avg_ms'], d['update']['avg_ms'], 1e3*d['loglikes']['avg_ms']/-(-d['n']//148), 1e3*d['update']['avg_ms']/-(-d['n']//148)))"; done
