#!/bin/bash
# time bench.py's device-resident step with every kernel variant under build/variants (phase-skip
# builds for profiling; their results are wrong by construction and never shipped)
for f in "" build/variants/lib_*.so; do
  if [ -z "$f" ]; then name=full; else name=$(basename $f .so); export DBOT_B200_LIB=$PWD/$f; fi
  python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$name', 'ms_per_step=%.3f'%d['ms_per_step'], 'kernel_ms=%.3f'%d['roofline']['avg_launch_ms'], 'e2e=%.0f'%d['e2e']['value'])"
done
