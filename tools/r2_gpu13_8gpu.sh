#!/bin/bash
# round 2, 8 GPUs: multi-GPU handle tests on the whole box, bench --gpus 8 (torchrun) incl. single-process config 4
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2_8gpu_devices.txt
timeout 600 python -m pytest tests/test_multi_device_gpu.py -m gpu -v > gpurun_out/r2_pytest13_8gpu.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|passed|failed" gpurun_out/r2_pytest13_8gpu.log | tail -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/r2_bench13_8gpu.json 2> gpurun_out/r2_bench13_8gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench13_8gpu.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value']); print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in d.get('single_process',{}).items()})"
tail -3 gpurun_out/r2_bench13_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 4 --steps 50 --warmup 3 > gpurun_out/r2_bench13_4gpu.json 2> gpurun_out/r2_bench13_4gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench13_4gpu.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value']); print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in d.get('single_process',{}).items()})"
