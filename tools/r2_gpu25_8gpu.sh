#!/bin/bash
DBOT_B200_HOST_TIMING=1 timeout 300 python tools/multi_host.py --steps 30 > gpurun_out/r2_multi_host25.log 2>&1
grep -v "^  GPU [1-6]" gpurun_out/r2_multi_host25.log | cut -c1-200
export DBOT_BENCH_SKIP_SINGLE=1
DBOT_B200_HOST_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/r2_bench25_8gpu.json 2> gpurun_out/r2_bench25_8gpu.err
grep "host timing" gpurun_out/r2_bench25_8gpu.err | head -3
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'pinned', d['e2e_pinned']['value'], d.get('gather',{}).get('mode'))" gpurun_out/r2_bench25_8gpu.json
