#!/bin/bash
# variants of the eval fetch / prepass, config 2 and 5 (per-mode CUDA-event timings)
for v in default branchy prepass2 both; do
  lib=build/variants/lib_$v.so; [ $v = default ] && lib=dbot_ros_b200/libdbot_b200.so
  for args in "--config 2 --reps 8" "--config 5 --reps 3"; do
    echo -n "$v $args: "
    DBOT_B200_LIB=$lib timeout 200 python tools/run_modes.py $args 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['loglikes']['avg_ms'], d['update']['avg_ms'])"
  done
done
