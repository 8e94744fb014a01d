#!/bin/bash
python tools/dbg_multi.py > gpurun_out/r2_dbg_multi.txt 2>&1
tail -30 gpurun_out/r2_dbg_multi.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest "tests/test_gpu_parity.py::test_loglikes_and_update_parity" -m gpu -x -q > gpurun_out/r2_racecheck5.log 2>&1
grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/r2_racecheck5.log | tail -5
rm -f gpurun_out/r2_modes5.jsonl
for v in su8 sw8 su8sw8; do
  echo "variant $v" | tee -a gpurun_out/r2_modes5.jsonl
  DBOT_B200_LIB=build/variants/lib_$v.so timeout 300 python tools/run_modes.py --config 2 --reps 8 2>&1 | tail -1 | tee -a gpurun_out/r2_modes5.jsonl
  DBOT_B200_LIB=build/variants/lib_$v.so timeout 300 python tools/run_modes.py --config 3 --reps 4 2>&1 | tail -1 | tee -a gpurun_out/r2_modes5.jsonl
done
