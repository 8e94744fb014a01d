#!/bin/bash
nvidia-smi -L > gpurun_out/r2_2gpu_devices34.txt
timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_multi_device_gpu.py tests/test_adapter_compiles.py -m gpu -v > gpurun_out/r2_pytest34_2gpu.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|passed|failed" gpurun_out/r2_pytest34_2gpu.log | tail -14
export DBOT_BENCH_SKIP_SINGLE=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r2_bench34_2gpu.json 2> gpurun_out/r2_bench34_2gpu.err
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'pinned', d['e2e_pinned']['value'], d.get('gather',{}).get('mode'))" gpurun_out/r2_bench34_2gpu.json
