#!/bin/bash
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/r2_pytest10_2gpu.log 2>&1
grep -E "passed|failed" gpurun_out/r2_pytest10_2gpu.log | tail -2
grep -E "Fatal|File \"|rank [01] frame|Error" gpurun_out/r2_pytest10_2gpu.log | head -60
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r2_bench10_2gpu.json 2> gpurun_out/r2_bench10_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench10_2gpu.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value']); print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in d.get('single_process',{}).items()})"
