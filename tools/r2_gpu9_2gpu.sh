#!/bin/bash
# round 2, 2 GPUs: NCCL slot migration (torchrun), multi-GPU behind the C-ABI (NCCL + fused gather), C++ adapter on 2 GPUs, bench --gpus 2
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2_2gpu_devices.txt
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_multi_device_gpu.py tests/test_adapter_compiles.py -m gpu -v > gpurun_out/r2_pytest9_2gpu.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|passed|failed|Error" gpurun_out/r2_pytest9_2gpu.log | tail -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r2_bench9_2gpu.json 2> gpurun_out/r2_bench9_2gpu.err
tail -c 1200 gpurun_out/r2_bench9_2gpu.json; tail -3 gpurun_out/r2_bench9_2gpu.err
