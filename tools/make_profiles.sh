#!/bin/bash
# profiles/ for round 2 from the files a gpurun evidence call left in gpurun_out/ (tools/evidence_1gpu.sh)
# usage: tools/make_profiles.sh <suffix of the evidence run, e.g. 14>
set -e
n=$1; o=gpurun_out; p=profiles
python tools/ncu_summary.py $o/r2_prof${n}_c2.ncu-rep --launches $o/r2_launches${n}.csv --title "round 2: particle kernel, config 2 (1000 particles, 640x480, 5120 triangles)" > $p/r02_particle_kernel_ncu_config2.md
python tools/ncu_summary.py $o/r2_prof${n}_c3.ncu-rep --title "round 2: particle kernel, config 3 (10 000 particles, 640x480, 5120 triangles; 24.6 GB of occlusion state)" > $p/r02_particle_kernel_ncu_config3.md
NCU_LAUNCH=0 python tools/ncu_lines.py $o/r2_prof${n}_c2.ncu-rep dbot_ros_b200/libdbot_b200.so "dbot_particle_kernelILi0ELb0" 40 > $p/r02_eval_kernel_hot_lines.txt
NCU_LAUNCH=1 python tools/ncu_lines.py $o/r2_prof${n}_c2.ncu-rep dbot_ros_b200/libdbot_b200.so "dbot_particle_kernelILi1ELb0" 40 > $p/r02_update_kernel_hot_lines.txt
python tools/ncu_traffic.py $o/r2_prof${n}_c2.ncu-rep "config 2: 1000 particles, 640x480, 5120 triangles, resampled parents (about 50 % distinct), touched-row copies" > $p/traffic.json
python tools/ncu_traffic.py $o/r2_prof${n}_c3.ncu-rep "config 3: 10 000 particles, 640x480, 5120 triangles, resampled parents" > $p/r02_traffic_config3.json
cp $o/r2_launches${n}.csv $p/r02_bench_launches.csv
cp $o/r2_modes${n}.jsonl $p/r02_modes.jsonl
cp $o/r2_timeline${n}.txt $p/r02_timeline.txt
cp $o/r2_bench${n}.json $p/r02_bench_1gpu.json
cp $o/r2_bench${n}_reference.json $p/r02_bench_reference.json
grep -E "ERROR SUMMARY|passed|failed|========= " $o/r2_memcheck${n}.log | grep -v "^=========$" | head -40 > $p/r02_sanitizer_memcheck.txt
grep -E "RACECHECK SUMMARY|passed|failed|========= " $o/r2_racecheck${n}.log | head -40 > $p/r02_sanitizer_racecheck.txt
tail -5 $o/r2_pytest${n}.log > $p/r02_pytest_1gpu.txt
bash tools/sass_evidence.sh > $p/r02_sass_evidence.txt
ls -la $p | tail -30
