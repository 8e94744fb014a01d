#!/bin/bash
# one gpurun call: GPU parity tests, bench line, ncu launch list, one ncu --set full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dbot_particle_kernel -s 3 -c 2 -f -o gpurun_out/prof_update python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
