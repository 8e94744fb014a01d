#!/bin/bash
for v in default pu2 pu4; do
  lib=build/variants/lib_$v.so; [ $v = default ] && lib=dbot_ros_b200/libdbot_b200.so
  for args in "--config 2 --reps 8"; do
    echo -n "$v $args: "
    DBOT_B200_LIB=$lib timeout 200 python tools/run_modes.py $args 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['loglikes']['avg_ms'], d['update']['avg_ms'])"
  done
done
