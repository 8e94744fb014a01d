#!/bin/bash
# round 2: convergent span units + lean mixture: parity, timings (config 2/3/5 + variants), timeline, ncu
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1
tail -4 gpurun_out/r2_pytest3.log
rm -f gpurun_out/r2_modes3.jsonl
for args in "--config 2 --reps 8" "--config 2 --reps 8 --nocull" "--config 2 --reps 8 --parents identity" "--config 3 --reps 4" "--config 5 --reps 4"; do
  timeout 300 python tools/run_modes.py $args 2>&1 | tail -1 | tee -a gpurun_out/r2_modes3.jsonl
done
for v in px2 px8; do
  echo "variant $v" | tee -a gpurun_out/r2_modes3.jsonl
  DBOT_B200_LIB=build/variants/lib_$v.so timeout 300 python tools/run_modes.py --config 2 --reps 8 2>&1 | tail -1 | tee -a gpurun_out/r2_modes3.jsonl
done
DBOT_B200_LIB=build/variants/lib_timeline.so timeout 300 python tools/run_modes.py --config 2 --reps 1 2>&1 | grep -E "^mode" | head -12 > gpurun_out/r2_timeline3.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dbot_particle_kernel -s 5 -c 2 -f -o gpurun_out/r2_prof3 python tools/run_modes.py --config 2 --reps 2 > gpurun_out/r2_ncu_run3.log 2>&1
head -4 gpurun_out/r2_timeline3.txt
