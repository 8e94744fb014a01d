#!/usr/bin/env python
"""The single-process multi-GPU handle (dbot_b200_config.n_devices) on its own: BASELINE config 2 per GPU and
config 4 at size through ONE handle, both gather modes; with DBOT_B200_HOST_TIMING=1 the library prints where the
host time of a frame goes.  usage: tools/multi_host.py [--gpus G] [--steps K] [--config4]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dbot_ros_b200 import CameraData, ObjectModel, Parameters, RbSensorBuilder, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--config4", action="store_true")
    a = ap.parse_args()
    import torch

    world = a.gpus or torch.cuda.device_count()
    w = bench.workload()
    one = RbSensorBuilder(ObjectModel(w["V"], w["T"]), CameraData(w["K"], w["rows"], w["cols"]),
                          Parameters(sample_count=8)).build()
    _, frames, _, _ = bench.make_inputs(dict(w, n=8), one.render, 0)
    one.close()
    for name, mode in (("nccl", 0), ("fused", 1)):
        out = bench.single_process_multi_gpu(w, world, w["n"], mode, frames, a.steps, 3)
        print(json.dumps({"what": "config2_per_gpu_" + name, **out}), flush=True)
        if a.config4:
            out = bench.single_process_multi_gpu(w, world, synth.CONFIGS[4][1], mode, frames, max(4, a.steps // 4), 3)
            print(json.dumps({"what": "config4_" + name, **out}), flush=True)


if __name__ == "__main__":
    main()
