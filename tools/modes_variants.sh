#!/bin/bash
# per-mode timings (config 2) of the shipped library and of every build/variants/lib_*.so
for f in "" build/variants/lib_*.so; do
  if [ -z "$f" ]; then name=shipped; unset DBOT_B200_LIB; else name=$(basename $f .so); export DBOT_B200_LIB=$PWD/$f; fi
  python tools/run_modes.py --config ${CFG:-2} --reps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%-16s'%'$name', 'eval_ms=%.4f update_ms=%.4f'%(d['loglikes']['avg_ms'], d['update']['avg_ms']))"
done
