#!/usr/bin/env python
"""Per-kernel timing of the depth observation model for one BASELINE config: the update=false and
update=true launches of dbot_particle_kernel timed separately with CUDA events on the library's
stream (dbot_b200_profile_*).  Also the command the ncu captures under profiles/ are taken from.
usage: tools/run_modes.py [--config 2|3|5] [--reps 5] [--n N]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dbot_ros_b200 import CameraData, ObjectModel, Parameters, RbSensorBuilder, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--parents", default="resampled", choices=["resampled", "identity", "few"],
                    help="parent pattern of the update calls: systematic resampling every frame, or none")
    ap.add_argument("--nocull", action="store_true", help="draw every triangle (no back-face culling)")
    a = ap.parse_args()
    import torch

    bench.CONFIG_ID = a.config
    w = bench.workload(a.config)
    if a.n:
        w["n"] = a.n
    n, rows, cols = w["n"], w["rows"], w["cols"]
    P = rows * cols
    sensor = RbSensorBuilder(ObjectModel(w["V"], w["T"]), CameraData(w["K"], rows, cols),
                             Parameters(sample_count=n, cull_backfaces=not a.nocull)).build()
    rng = synth.rng_for(a.config)
    gt = synth.ground_truth_poses(len(w["V"]), rng)
    frames = [synth.make_frame(sensor.render, gt, rows, cols, rng) for _ in range(2)]
    poses = [synth.particle_poses(gt, n, rng) for _ in range(2)]
    def make_parents():
        if a.parents == "resampled":
            return bench.resampled_parents(n, rng)
        if a.parents == "identity":
            return np.arange(n, dtype=np.int32)
        # "few": what a peaked likelihood does to systematic resampling -- a handful of distinct parents
        return np.sort(rng.choice(rng.choice(n, 8, replace=False), n)).astype(np.int32)

    parents = [make_parents() for _ in range(2)]
    dev = torch.device("cuda", 0)
    d_frames = [torch.from_numpy(f).to(dev) for f in frames]
    d_poses = [torch.from_numpy(p).to(dev) for p in poses]
    d_parents = [torch.from_numpy(p).to(dev) for p in parents]
    d_idx = torch.zeros(n, dtype=torch.int32, device=dev)
    d_out = torch.zeros(n, dtype=torch.float64, device=dev)
    ext = torch.cuda.ExternalStream(sensor.stream, device=dev)
    sensor.set_observation(frames[0])
    sensor.loglikes_poses(poses[0], np.zeros(n, dtype=np.int32), True)   # n slots
    sensor.profile_enable(True)
    for k in range(a.reps + 2):
        f = k % 2
        if k == 2:
            sensor.profile_read()  # drop the warm-up
        with torch.cuda.stream(ext):
            d_idx.copy_(d_parents[f], non_blocking=True)
            sensor.set_observation_device(d_frames[f].data_ptr())
            for _ in range(len(w["V"])):  # one non-updating call per sampling block, then the update
                sensor.loglikes_device(d_poses[f].data_ptr(), d_idx.data_ptr(), n, False, d_out.data_ptr())
            sensor.loglikes_device(d_poses[f].data_ptr(), d_idx.data_ptr(), n, True, d_out.data_ptr())
        sensor.synchronize()
    prof = sensor.profile_read()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    out = {"config": a.config, "n": n, "parents": a.parents, "cull": not a.nocull, "resolution": [cols, rows], "triangles": w["tris"],
           "device_bytes": sensor.device_bytes}
    for mode, bytes_per in (("loglikes", None), ("update", 8.0 * P)):
        ms, cnt = prof[mode]
        avg = ms / max(cnt, 1)
        out[mode] = {"avg_ms": avg, "launches": cnt, "pl_per_s": n / (avg * 1e-3)}
        if bytes_per:
            out[mode]["alg_GBps"] = bytes_per * n / (avg * 1e-3) / 1e9
            out[mode]["frac_of_measured_hbm"] = out[mode]["alg_GBps"] / peak
    print(json.dumps(out))


if __name__ == "__main__":
    main()
