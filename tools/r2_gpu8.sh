#!/bin/bash
# round 2: touched-row copies
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest8.log 2>&1
tail -3 gpurun_out/r2_pytest8.log
rm -f gpurun_out/r2_modes8.jsonl
for args in "--config 2 --reps 8" "--config 2 --reps 8 --parents identity" "--config 2 --reps 8 --parents few" "--config 3 --reps 4" "--config 3 --reps 4 --parents few" "--config 5 --reps 4"; do
  timeout 200 python tools/run_modes.py $args 2>&1 | tail -1 | tee -a gpurun_out/r2_modes8.jsonl
done
timeout 600 compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_parity.py::test_updates_copy_only_touched_rows_and_images_stay_exact" "tests/test_gpu_parity.py::test_multi_frame_with_resampling" -m gpu -x -q > gpurun_out/r2_memcheck8.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_memcheck8.log | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
tail -c 400 gpurun_out/r2_bench8.json; tail -2 gpurun_out/r2_bench8.err
