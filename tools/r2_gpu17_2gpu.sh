#!/bin/bash
nvidia-smi -L > gpurun_out/r2_2gpu_devices17.txt
timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_multi_device_gpu.py -m gpu -x -q -s > gpurun_out/r2_pytest17_2gpu.log 2>&1
tail -5 gpurun_out/r2_pytest17_2gpu.log
export DBOT_BENCH_SKIP_SINGLE=1
for g in fused collective; do
DBOT_B200_SHARDED_GATHER=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r2_bench17_2gpu_$g.json 2> gpurun_out/r2_bench17_2gpu_$g.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench17_2gpu_$g.json").read().strip().splitlines()[-1])
print("$g", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_pinned"]["value"], d.get("gather"))
PY
done
