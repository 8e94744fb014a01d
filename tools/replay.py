#!/usr/bin/env python
"""Offline replay harness (shape of the reference's unbuilt
TrackingDataset / DataSetCameraDataProvider, ref:source/dbot_ros/util/tracking_dataset.cpp:92-101,176-285):
depth frames + ground-truth poses in, tracking error + frames/s out.

  tools/replay.py --synthetic [--config 2] [--frames 60] [--tracker device|host]
  tools/replay.py --frames-npy depth.npy --truth-npy poses.npy --mesh a.obj [b.obj ...] --K fx fy cx cy

depth.npy: [F][rows][cols] float32 metres (NaN = no data); poses.npy: [F][n_obj][6] position + rotation
vector; the first ground-truth pose (plus --init-offset) initialises the tracker, as the interactive
marker does in the reference (ref:source/dbot_ros/tracker/particle_tracker_node.cpp:226-252)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dbot_ros_b200 import CameraData, ObjectModel, Parameters, RbSensorBuilder, io, synth  # noqa: E402
from dbot_ros_b200.tracker import DeviceParticleTracker, ParticleTracker, TrackerParameters  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--particles", type=int, default=0)
    ap.add_argument("--tracker", default="device", choices=["device", "host"])
    ap.add_argument("--frames-npy")
    ap.add_argument("--truth-npy")
    ap.add_argument("--mesh", nargs="*")
    ap.add_argument("--K", nargs=4, type=float)
    ap.add_argument("--init-offset", type=float, default=0.004)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    if a.synthetic:
        mesh, n, ds = synth.CONFIGS[a.config]
        rows, cols = synth.resolution(ds)
        K = synth.camera_matrix(ds)
        V, T = synth.mesh_set(mesh)
        model = ObjectModel(V, T)
    else:
        depth = np.load(a.frames_npy)
        rows, cols = depth.shape[1:]
        K = np.array([[a.K[0], 0, a.K[2]], [0, a.K[1], a.K[3]], [0, 0, 1.0]])
        model = io.object_model_from_files("", a.mesh, center_object_frame=False)
        n = 1000
    n_obj = model.count_parts
    n = a.particles or n
    sensor = RbSensorBuilder(model, CameraData(K, rows, cols), Parameters(sample_count=n)).build()
    if a.synthetic:
        gt0 = synth.ground_truth_poses(n_obj, rng)
        start = np.array([np.concatenate([t, [0, 0, 0]]) for _, t in gt0])
        vel = rng.normal(scale=[0.0015, 0.0015, 0.0015, 0.01, 0.01, 0.01], size=(n_obj, 6))
        truth = np.stack([start + f * vel for f in range(a.frames)])
        depth = np.stack([synth.make_frame(sensor.render, [(synth.rodrigues(p[3:6]), p[0:3]) for p in truth[f]],
                                           rows, cols, rng, occluder=False) for f in range(a.frames)])
    else:
        truth = np.load(a.truth_npy)
    prm = TrackerParameters(evaluation_count=n * n_obj)
    tracker = (DeviceParticleTracker(sensor, prm, seed=a.seed) if a.tracker == "device"
               else ParticleTracker(sensor, n_obj, prm, seed=a.seed))
    init = truth[0].copy()
    init[:, :3] += a.init_offset * np.array([1.0, -1.0, 1.0]) / np.sqrt(3.0)
    tracker.initialize(init)
    est, times = [], []
    for f in range(len(depth)):
        t0 = time.perf_counter()
        est.append(tracker.track(depth[f]))
        times.append(time.perf_counter() - t0)
    est = np.array(est)
    err = np.linalg.norm(est[:, :, :3] - truth[:, :, :3], axis=2)
    warm = times[3:] if len(times) > 6 else times
    print(json.dumps({"tracker": a.tracker, "frames": len(depth), "particles_per_block": n, "objects": n_obj,
                      "resolution": [cols, rows], "position_error_mm": {"mean": 1e3 * float(err.mean()),
                      "max": 1e3 * float(err.max()), "last": 1e3 * float(err[-1].max())},
                      "ms_per_frame": 1e3 * float(np.mean(warm)), "fps": 1.0 / float(np.mean(warm)),
                      "particle_likelihoods_per_s": n * n_obj / float(np.mean(warm)),
                      "resamplings": int(tracker.resample_count)}))


if __name__ == "__main__":
    main()
