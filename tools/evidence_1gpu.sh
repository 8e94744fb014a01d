#!/bin/bash
# usage (through gpurun, from the repo root): bash tools/evidence_1gpu.sh [suffix]; then tools/make_profiles.sh <suffix>
N=${1:-ev}
# round 2 evidence run (1 GPU): parity, sanitizers, per-mode timings, ncu (config 2 + config 3), launch list, bench
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest${N}.log 2>&1
tail -3 gpurun_out/r2_pytest${N}.log
rm -f gpurun_out/r2_modes${N}.jsonl
for args in "--config 2 --reps 8" "--config 2 --reps 8 --nocull" "--config 2 --reps 8 --parents identity" "--config 2 --reps 8 --parents few" "--config 3 --reps 4" "--config 3 --reps 4 --parents identity" "--config 5 --reps 4"; do
  timeout 200 python tools/run_modes.py $args 2>&1 | tail -1 | tee -a gpurun_out/r2_modes${N}.jsonl
done
DBOT_B200_LIB=build/variants/lib_timeline.so timeout 200 python tools/run_modes.py --config 2 --reps 1 2>&1 | grep -E "^mode" | head -12 > gpurun_out/r2_timeline${N}.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dbot_particle_kernel -s 5 -c 2 -f -o gpurun_out/r2_prof${N}_c2 python tools/run_modes.py --config 2 --reps 2 > gpurun_out/r2_ncu_run${N}.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:dbot_particle_kernel -s 5 -c 2 -f -o gpurun_out/r2_prof${N}_c3 python tools/run_modes.py --config 3 --reps 2 >> gpurun_out/r2_ncu_run${N}.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches${N}.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch_run${N}.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_parity.py::test_loglikes_and_update_parity" "tests/test_gpu_parity.py::test_multi_frame_with_resampling" "tests/test_gpu_parity.py::test_edge_cases" "tests/test_gpu_parity.py::test_slot_adoption_patterns_and_pool_accounting" "tests/test_multi_device_gpu.py::test_shards_sharing_one_gpu_fused_gather" "tests/test_sharded_gpu.py::test_slot_export_import_between_two_handles_on_one_gpu" "tests/test_gpu_parity.py::test_updates_copy_only_touched_rows_and_images_stay_exact" "tests/test_gpu_parity.py::test_tall_and_wide_triangles_close_to_the_camera" -m gpu -x -q > gpurun_out/r2_memcheck${N}.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_memcheck${N}.log | tail -3
timeout 600 compute-sanitizer --tool racecheck python -m pytest "tests/test_gpu_parity.py::test_loglikes_and_update_parity" "tests/test_gpu_parity.py::test_multi_frame_with_resampling" "tests/test_gpu_parity.py::test_updates_copy_only_touched_rows_and_images_stay_exact" -m gpu -x -q > gpurun_out/r2_racecheck${N}.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2_racecheck${N}.log | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench${N}.json 2> gpurun_out/r2_bench${N}.err
timeout 300 python bench.py --impl reference --steps 20 > gpurun_out/r2_bench${N}_reference.json 2>> gpurun_out/r2_bench${N}.err
tail -c 600 gpurun_out/r2_bench${N}.json
