#!/bin/bash
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2_2gpu_devices.txt
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_multi_device_gpu.py tests/test_adapter_compiles.py -m gpu -v -s > gpurun_out/r2_pytest11_2gpu.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|passed|failed|bit-identical" gpurun_out/r2_pytest11_2gpu.log | tail -20
