"""The callers and data formats either side of the hot path: the
host-side particle-filter loop that drives the sensor, and the .obj reader.  The CPU test runs the
filter on the oracle (tests may use it); the GPU test runs it on the CUDA sensor and checks that
the log-weights actually steer the filter onto a moving object."""
import numpy as np
import pytest

import oracle_lib
from dbot_ros_b200 import CameraData, ObjectModel, Parameters, RbSensorBuilder, io, synth
from dbot_ros_b200.rb_sensor import compose_poses
from dbot_ros_b200.tracker import (ParticleTracker, TrackerParameters, rotation_vector, rotation_vectors,
                                   systematic_resample)


class OracleSensor:
    """RbSensor interface over the CPU oracle (test infrastructure)"""

    def __init__(self, orc):
        self.orc = orc
        self.integrated_poses = np.zeros((orc.n_objects, 6))

    def reset(self):
        self.orc.reset()

    def set_observation(self, image):
        self.orc.set_observation(image)

    def loglikes(self, deltas, indices, update=False):
        return self.orc.loglikes_poses(compose_poses(deltas, self.integrated_poses), indices, update)


def moving_scene(render, rows, cols, frames, rng, start, step):
    """object translating by `step` per frame from `start`; frames rendered with `render`"""
    out, truth = [], []
    for f in range(frames):
        t = np.asarray(start) + f * np.asarray(step)
        gt = [(np.eye(3), t)]
        out.append(synth.make_frame(render, gt, rows, cols, rng, occluder=False))
        truth.append(t)
    return out, np.array(truth)


def run_tracker(sensor, render, rows, cols, n_particles, frames, seed):
    rng = np.random.default_rng(seed)
    start, step = np.array([0.0, 0.0, 0.8]), np.array([0.002, -0.001, 0.0015])
    images, truth = moving_scene(render, rows, cols, frames, rng, start, step)
    tr = ParticleTracker(sensor, 1, TrackerParameters(evaluation_count=n_particles), seed=seed)
    tr.initialize([[start[0] + 0.004, start[1] - 0.003, start[2] + 0.004, 0, 0, 0]])  # 6 mm off
    est = np.array([tr.track(img)[0, :3] for img in images])
    return tr, est, truth


def test_resampling_and_rotation_helpers():
    w = np.array([0.1, 0.0, 0.6, 0.3])
    p = systematic_resample(w, 10, 0.5)
    assert np.all(np.diff(p) >= 0) and list(np.bincount(p, minlength=4)) == [1, 0, 6, 3]
    rng = np.random.default_rng(2)
    rv = rng.normal(size=(6, 3)) * 0.7
    R = np.stack([synth.rodrigues(v) for v in rv])
    assert np.allclose(rotation_vectors(R), rv, atol=1e-12)
    assert np.allclose(rotation_vector(np.eye(3)), 0.0)


def test_filter_loop_on_the_oracle_tracks_a_moving_object():
    rows, cols = synth.resolution(8)
    K = synth.camera_matrix(8)
    V, T = synth.mesh_set("ico80")
    orc = oracle_lib.Oracle(rows, cols, K, V, T, synth.default_parameters())
    tr, est, truth = run_tracker(OracleSensor(orc), orc.render, rows, cols, 80, 10, seed=3)
    err = np.linalg.norm(est - truth, axis=1)
    assert err.max() < 0.004                                       # pulled in from the 6.4 mm offset and kept
    assert orc.n_slots == tr.n                                     # update ran on the last block
    assert tr.resample_count >= 1                                  # indices were permuted at least once


def test_obj_reader_roundtrip_and_polygons(tmp_path):
    V, T = synth.mesh_set("ico80")
    path = tmp_path / "ico.obj"
    io.save_obj(path, V[0], T[0])
    v, t = io.load_obj(path)
    assert np.allclose(v, V[0], atol=1e-7) and np.array_equal(t, T[0])
    quad = tmp_path / "quad.obj"
    quad.write_text("# quad with texture/normal indices and a relative face\n"
                    "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvn 0 0 1\n"
                    "f 1/1/1 2/1/1 3/1/1 4/1/1\nf -4//1 -3//1 -2//1\n")
    v, t = io.load_obj(quad)
    assert v.shape == (4, 3) and t.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]]
    m = io.object_model_from_files(str(tmp_path), ["ico.obj", "quad.obj"], center_object_frame=True)
    assert m.count_parts == 2 and np.allclose(m.vertices[1].mean(axis=0), 0, atol=1e-7)
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(ValueError):
        io.load_obj(bad)


@pytest.mark.gpu
def test_tracker_on_the_cuda_sensor_follows_the_object():
    rows, cols = synth.resolution(4)
    K = synth.camera_matrix(4)
    V, T = synth.mesh_set("icosphere5k")
    sensor = RbSensorBuilder(ObjectModel(V, T), CameraData(K, rows, cols), Parameters(sample_count=400)).build()
    tr, est, truth = run_tracker(sensor, sensor.render, rows, cols, 400, 30, seed=5)
    err = np.linalg.norm(est - truth, axis=1)
    assert err[-10:].max() < 0.003, err                            # within 3 mm while it moves 2.7 mm/frame
    assert sensor.n_slots == 400 and tr.resample_count >= 1
