#!/usr/bin/env python
"""Offline replay harness, command line (the library part is dbot_ros_b200/replay.py).

  tools/replay.py --synthetic [--config 2] [--frames 60] [--tracker device|host] [--save-dataset DIR]
  tools/replay.py --dataset DIR --mesh a.obj [b.obj ...] [--downsampling 1] [--center-object-frame]

DIR holds ground_truth.txt in the reference's layout (ref:source/dbot_ros/util/tracking_dataset.cpp:325-357:
"<stamp> <12 state values per object>" per line) and measurements.npz (depth [F][rows][cols] float32 metres,
stamps [F], K [3][3]) standing in for measurements.bag, which needs ROS to read."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dbot_ros_b200 import CameraData, ObjectModel, Parameters, RbSensorBuilder, io, replay, synth  # noqa: E402
from dbot_ros_b200.tracker import TrackerParameters  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--particles", type=int, default=0)
    ap.add_argument("--tracker", default="device", choices=["device", "host"])
    ap.add_argument("--dataset")
    ap.add_argument("--save-dataset")
    ap.add_argument("--mesh", nargs="*")
    ap.add_argument("--center-object-frame", action="store_true")
    ap.add_argument("--moving-average-update-rate", type=float, default=1.0)
    ap.add_argument("--init-offset", type=float, default=0.004)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    if a.synthetic:
        mesh, n, ds = synth.CONFIGS[a.config]
        rows, cols = synth.resolution(ds)
        K = synth.camera_matrix(ds)
        V, T = synth.mesh_set(mesh)
        model = ObjectModel(V, T, center_object_frame=a.center_object_frame)
    else:
        d = io.load_dataset(a.dataset)
        rows, cols = d["depth"].shape[1:]
        K = d["K"]
        model = io.object_model_from_files("", a.mesh, center_object_frame=a.center_object_frame)
        n = 1000
    n_obj = model.count_parts
    n = a.particles or n
    sensor = RbSensorBuilder(model, CameraData(K, rows, cols), Parameters(sample_count=n)).build()
    prm = TrackerParameters(evaluation_count=n * n_obj, center_object_frame=a.center_object_frame,
                            moving_average_update_rate=a.moving_average_update_rate)
    kw = dict(tracker=a.tracker, params=prm, centers=model.centers, seed=a.seed, init_offset=a.init_offset)
    if a.synthetic:
        # frames are rendered with the un-centred meshes' poses: the model frame is what ground truth speaks
        plain = sensor if not a.center_object_frame else RbSensorBuilder(
            ObjectModel(V, T), CameraData(K, rows, cols), Parameters(sample_count=8)).build()
        depth, stamps, states = replay.synthetic_sequence(plain.render, n_obj, rows, cols, a.frames, rng)
        if a.save_dataset:
            io.save_dataset(a.save_dataset, depth, stamps, K, states)
        out, _ = replay.replay(sensor, depth, io.ground_truth_poses(states, n_obj), **kw)
    else:
        out, _ = replay.replay_dataset(a.dataset, sensor, n_obj, **kw)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
