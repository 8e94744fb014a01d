#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest36.log 2>&1
tail -12 gpurun_out/r2_pytest36.log
