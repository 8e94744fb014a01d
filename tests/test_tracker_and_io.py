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


def test_counter_noise_is_standard_normal_and_reproducible():
    from dbot_ros_b200.tracker import counter_normals, counter_uniform

    z = counter_normals(7, 2, 1, 50000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert np.abs(np.corrcoef(z.T) - np.eye(6)).max() < 0.02
    assert np.array_equal(z[:100], counter_normals(7, 2, 1, 100))          # a function of its counters only
    assert not np.allclose(z[:100], counter_normals(7, 3, 1, 100))
    assert 0.0 <= counter_uniform(7, 2, 1) < 1.0


@pytest.mark.gpu
def test_device_resident_filter_matches_the_host_loop_and_tracks():
    """dbot_b200_filter_*: the whole filter step on the GPU.  With the same counter-based noise the
    host loop (numpy, driving an identical sensor) must produce the same track; two objects so that
    there are two sampling blocks per frame."""
    from dbot_ros_b200.tracker import DeviceParticleTracker

    rows, cols = synth.resolution(4)
    K = synth.camera_matrix(4)
    a = synth.icosphere(2, radius=0.05, seed=4)
    b = synth.torus(24, 12, R=0.05, r=0.02)
    V, T = [a[0], b[0]], [a[1], b[1]]
    n_eval = 600                                                  # 300 particles per object block
    mk = lambda: RbSensorBuilder(ObjectModel(V, T), CameraData(K, rows, cols), Parameters(sample_count=300)).build()  # noqa: E731
    s_dev, s_host = mk(), mk()
    rng = np.random.default_rng(12)
    start = np.array([[-0.08, 0.0, 0.75, 0.0, 0.0, 0.0], [0.09, 0.01, 0.85, 0.3, -0.2, 0.1]])
    step = np.array([[0.0015, -0.001, 0.001, 0, 0, 0.01], [-0.001, 0.001, -0.0015, 0, 0.01, 0]])
    frames, truth = [], []
    for f in range(12):
        pose = start + f * step
        gt = [(synth.rodrigues(pose[o, 3:6]), pose[o, 0:3]) for o in range(2)]
        frames.append(synth.make_frame(s_dev.render, gt, rows, cols, rng, occluder=False))
        truth.append(pose[:, :3].copy())
    prm = TrackerParameters(evaluation_count=n_eval)
    init = start.copy()
    init[:, :3] += [[0.004, -0.003, 0.004], [-0.003, 0.004, -0.004]]
    dev = DeviceParticleTracker(s_dev, prm, seed=99)
    host = ParticleTracker(s_host, 2, prm, seed=99, counter_noise=True)
    dev.initialize(init)
    host.initialize(init)
    assert dev.n == host.n == 300
    est_d, est_h = [], []
    for img in frames:
        est_d.append(dev.track(img))
        est_h.append(host.track(img))
    est_d, est_h, truth = np.array(est_d), np.array(est_h), np.array(truth)
    # same noise, same sensor arithmetic: the two tracks agree far below the tracking error
    assert np.max(np.abs(est_d[:4] - est_h[:4])) < 1e-6, np.max(np.abs(est_d - est_h), axis=(1, 2))
    assert np.max(np.abs(est_d - est_h)) < 5e-4, np.max(np.abs(est_d - est_h), axis=(1, 2))
    err = np.linalg.norm(est_d[:, :, :3] - truth, axis=2)
    assert err[-6:].max() < 0.004, err
    assert dev.resample_count >= 1 and abs(dev.resample_count - host.resample_count) <= 2
    assert s_dev.n_slots == 300


@pytest.mark.gpu
def test_device_filter_argument_errors():
    from dbot_ros_b200 import DbotB200Error
    from dbot_ros_b200.tracker import DeviceParticleTracker

    rows, cols = synth.resolution(8)
    V, T = synth.mesh_set("ico80")
    sensor = RbSensorBuilder(ObjectModel(V, T), CameraData(synth.camera_matrix(8), rows, cols),
                             Parameters(sample_count=16)).build()
    with pytest.raises(DbotB200Error):                      # more particles than the sensor was built for
        DeviceParticleTracker(sensor, TrackerParameters(evaluation_count=17))
    trk = DeviceParticleTracker(sensor, TrackerParameters(evaluation_count=16), seed=3)
    trk.initialize([[0.0, 0.0, 0.8, 0.0, 0.0, 0.0]])
    frame = synth.make_frame(sensor.render, [(np.eye(3), np.array([0.0, 0.0, 0.8]))], rows, cols,
                             np.random.default_rng(0), occluder=False)
    a = trk.track(frame)
    trk.initialize([[0.0, 0.0, 0.8, 0.0, 0.0, 0.0]])        # same seed, same frame counter: same result
    b = trk.track(frame)
    assert np.array_equal(a, b) and np.linalg.norm(a[0, :3] - [0, 0, 0.8]) < 0.01
    trk.close()
