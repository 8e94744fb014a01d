#!/bin/bash
# round 2: register records + flat mixture: parity, timings, ncu, bench line
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1
tail -4 gpurun_out/r2_pytest4.log
rm -f gpurun_out/r2_modes4.jsonl
for args in "--config 2 --reps 8" "--config 2 --reps 8 --nocull" "--config 2 --reps 8 --parents identity" "--config 3 --reps 4" "--config 3 --reps 4 --parents identity" "--config 5 --reps 4"; do
  timeout 300 python tools/run_modes.py $args 2>&1 | tail -1 | tee -a gpurun_out/r2_modes4.jsonl
done
for v in px8; do
  echo "variant $v" | tee -a gpurun_out/r2_modes4.jsonl
  DBOT_B200_LIB=build/variants/lib_$v.so timeout 300 python tools/run_modes.py --config 2 --reps 8 2>&1 | tail -1 | tee -a gpurun_out/r2_modes4.jsonl
done
DBOT_B200_LIB=build/variants/lib_timeline.so timeout 300 python tools/run_modes.py --config 2 --reps 1 2>&1 | grep -E "^mode" | head -12 > gpurun_out/r2_timeline4.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dbot_particle_kernel -s 5 -c 2 -f -o gpurun_out/r2_prof4 python tools/run_modes.py --config 2 --reps 2 > gpurun_out/r2_ncu_run4.log 2>&1
timeout 600 python bench.py > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
tail -c 1500 gpurun_out/r2_bench4.json; tail -3 gpurun_out/r2_bench4.err
head -4 gpurun_out/r2_timeline4.txt
