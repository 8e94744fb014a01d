"""Offline replay: depth frames + ground-truth poses in, tracking error and
frames/s out -- the shape of the reference's TrackingDataset / DataSetCameraDataProvider harness
(ref:source/dbot_ros/util/tracking_dataset.cpp:92-101,176-285, ref:source/dbot_ros/util/data_set_camera_data_provider.cpp:27-61;
commented out of the reference's build, ref:CMakeLists.txt:98-101).  The ground truth comes in the reference's
own `ground_truth.txt` layout (dbot_ros_b200.io); the depth frames, which the reference keeps in a rosbag, come
as an array (there is no ROS in this image)."""
import time

import numpy as np

from . import io, synth
from .tracker import DeviceParticleTracker, ParticleTracker, TrackerParameters


def synthetic_sequence(render, n_objects, rows, cols, n_frames, rng, dt=1.0 / 30.0):
    """objects drifting at constant velocity from seeded start poses; -> (depth [F,rows,cols], stamps [F],
    states [F, n_objects*12] as ground_truth.txt holds them)"""
    gt0 = synth.ground_truth_poses(n_objects, rng)
    start = np.array([np.concatenate([t, [0.0, 0.0, 0.0]]) for _, t in gt0])
    vel = rng.normal(scale=[0.0015, 0.0015, 0.0015, 0.01, 0.01, 0.01], size=(n_objects, 6))   # per frame
    states = np.zeros((n_frames, n_objects, 12))
    depth = []
    for f in range(n_frames):
        pose = start + f * vel
        states[f, :, :6] = pose
        states[f, :, 6:] = vel / dt
        depth.append(synth.make_frame(render, [(synth.rodrigues(p[3:6]), p[0:3]) for p in pose], rows, cols, rng,
                                      occluder=False))
    stamps = 1000.0 + dt * np.arange(n_frames)
    return np.stack(depth), stamps, states.reshape(n_frames, -1)


def replay(sensor, depth, truth, tracker="device", params=None, centers=None, seed=1, init_offset=0.004):
    """truth: [F, n_obj, 6] model-frame poses (NaN rows = no ground truth for that frame).  The first pose plus
    `init_offset` metres initialises the tracker, as the interactive marker does in the reference
    (ref:source/dbot_ros/tracker/particle_tracker_node.cpp:226-252).  -> dict with error and rate figures"""
    n_obj = sensor.n_objects
    params = params or TrackerParameters(evaluation_count=sensor.max_sample_count * n_obj)
    trk = (DeviceParticleTracker(sensor, params, seed=seed, centers=centers) if tracker == "device"
           else ParticleTracker(sensor, n_obj, params, seed=seed, centers=centers))
    init = np.array(truth[0], dtype=np.float64, copy=True)
    init[:, :3] += init_offset * np.array([1.0, -1.0, 1.0]) / np.sqrt(3.0)
    trk.initialize(init)
    est, times = [], []
    for f in range(len(depth)):
        t0 = time.perf_counter()
        est.append(trk.track(depth[f]))
        times.append(time.perf_counter() - t0)
    est = np.array(est)
    have = np.isfinite(truth[:, :, 0]).all(axis=1)
    err = np.linalg.norm(est[have][:, :, :3] - truth[have][:, :, :3], axis=2)
    warm = times[3:] if len(times) > 6 else times
    n = trk.n
    out = {"tracker": tracker, "frames": int(len(depth)), "frames_with_ground_truth": int(have.sum()),
           "particles_per_block": int(n), "objects": int(n_obj), "resolution": [int(depth.shape[2]), int(depth.shape[1])],
           "position_error_mm": {"mean": 1e3 * float(err.mean()), "max": 1e3 * float(err.max()),
                                 "last": 1e3 * float(err[-1].max())},
           "ms_per_frame": 1e3 * float(np.mean(warm)), "fps": 1.0 / float(np.mean(warm)),
           "particle_likelihoods_per_s": n * n_obj / float(np.mean(warm)), "resamplings": int(trk.resample_count)}
    if hasattr(trk, "close"):
        trk.close()
    return out, est


def replay_dataset(directory, sensor, n_objects, **kw):
    """a dataset directory (io.load_dataset): frames without a ground-truth line within the admissible time
    difference are tracked but not scored"""
    d = io.load_dataset(directory)
    poses = io.ground_truth_poses(d["gt_states"], n_objects)
    truth = np.full((len(d["depth"]), n_objects, 6), np.nan)
    ok = d["gt_index"] >= 0
    truth[ok] = poses[d["gt_index"][ok]]
    if not ok[0]:
        raise ValueError("%s: the first frame has no ground truth to initialise the tracker from" % directory)
    return replay(sensor, d["depth"], truth, **kw)
