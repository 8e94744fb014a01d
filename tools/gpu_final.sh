#!/bin/bash
# final round measurements on one GPU: parity suite, bench (both arms), per-mode timings, ncu launch list + full capture
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_final.err
for c in 2 3 5; do timeout 200 python tools/run_modes.py --config $c > gpurun_out/modes_c$c.json 2>> gpurun_out/modes.err; done
timeout 200 python tools/run_modes.py --config 2 --parents identity > gpurun_out/modes_c2_identity.json 2>> gpurun_out/modes.err
timeout 200 python tools/run_modes.py --config 3 --parents identity > gpurun_out/modes_c3_identity.json 2>> gpurun_out/modes.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dbot_particle_kernel -s 5 -c 2 -f -o gpurun_out/prof_modes python tools/run_modes.py --config 2 --reps 2 > gpurun_out/ncu_full_run.log 2>&1
tail -2 gpurun_out/pytest_gpu.log; cat gpurun_out/modes_c2.json
