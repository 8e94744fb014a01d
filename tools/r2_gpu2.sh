#!/bin/bash
# round 2: where does the span rasteriser spend its time?  phase timeline + ncu full capture (config 2)
DBOT_B200_LIB=build/variants/libdbot_timeline.so timeout 300 python tools/run_modes.py --config 2 --reps 1 2>&1 | grep -E "^mode|config" | head -40 > gpurun_out/r2_timeline.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dbot_particle_kernel -s 5 -c 2 -f -o gpurun_out/r2_prof_modes python tools/run_modes.py --config 2 --reps 2 > gpurun_out/r2_ncu_run.log 2>&1
timeout 600 python -m pytest tests/test_multi_device_gpu.py tests/test_tracker_and_io.py -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1
tail -3 gpurun_out/r2_pytest2.log; cat gpurun_out/r2_timeline.txt | head -12
