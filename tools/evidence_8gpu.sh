#!/bin/bash
# round 2, 8 GPUs: sharded + multi-device tests on the whole box; bench --gpus 8 under torchrun with the fused
# gather (incl. the single-process section and config 4 at size) and with the NCCL all-gather; bench --gpus 4
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2_8gpu_devices28.txt
nproc >> gpurun_out/r2_8gpu_devices28.txt
timeout 600 python -m pytest tests/test_multi_device_gpu.py tests/test_sharded_gpu.py -m gpu -v > gpurun_out/r2_pytest28_8gpu.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|passed|failed" gpurun_out/r2_pytest28_8gpu.log | tail -12
show() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'pinned', d['e2e_pinned']['value'], d.get('gather',{}).get('mode')); print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in d.get('single_process',{}).items()})" $1; }
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/r2_bench28_8gpu.json 2> gpurun_out/r2_bench28_8gpu.err
show gpurun_out/r2_bench28_8gpu.json; tail -3 gpurun_out/r2_bench28_8gpu.err
export DBOT_BENCH_SKIP_SINGLE=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 4 --steps 50 --warmup 3 > gpurun_out/r2_bench28_4gpu.json 2> gpurun_out/r2_bench28_4gpu.err
show gpurun_out/r2_bench28_4gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29564 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r2_bench28_2gpu.json 2> gpurun_out/r2_bench28_2gpu.err
show gpurun_out/r2_bench28_2gpu.json
