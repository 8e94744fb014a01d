#!/bin/bash
# one gpurun call: launch list + ncu --set full of the update and eval kernels (config 2), mode timings
mkdir -p gpurun_out
python tools/run_modes.py --config 2 > gpurun_out/modes_c2.json 2> gpurun_out/modes.err
python tools/run_modes.py --config 3 > gpurun_out/modes_c3.json 2>> gpurun_out/modes.err
python tools/run_modes.py --config 5 > gpurun_out/modes_c5.json 2>> gpurun_out/modes.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
# particle kernels of run_modes: 2 renders, setup update, then (eval, update) x reps -> skip 5: eval, update
ncu --set full --clock-control none --import-source on -k regex:dbot_particle_kernel -s 5 -c 2 -f -o gpurun_out/prof_modes python tools/run_modes.py --config 2 --reps 2 > gpurun_out/ncu_full_run.log 2>&1
cat gpurun_out/modes_c2.json gpurun_out/modes_c3.json gpurun_out/modes_c5.json; tail -3 gpurun_out/modes.err
