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


def test_center_object_frame_poses_cross_the_tracker_in_the_model_frame():
    """center_object_frame (ref:config/particle_tracker.yaml:30, ref:source/dbot_ros/tracker/particle_tracker_node.cpp:
    92-97,212-214): the mesh is shifted so that its centroid is the origin; dbot's Tracker converts the initial
    poses into that frame (t += R c) and the published ones back (t -= R c).  The same scene tracked with an
    off-centre mesh, once centred and once not, must give the same model-frame track."""
    rows, cols = synth.resolution(8)
    K = synth.camera_matrix(8)
    V, T = synth.mesh_set("ico80")
    V = [V[0] + np.array([0.05, -0.03, 0.02], np.float32)]            # model origin far from the centroid
    plain = ObjectModel(V, T, center_object_frame=False)
    centred = ObjectModel(V, T, center_object_frame=True)
    assert np.linalg.norm(centred.centers[0]) > 0.05
    o_plain = oracle_lib.Oracle(rows, cols, K, plain.vertices, plain.triangles, synth.default_parameters())
    o_centred = oracle_lib.Oracle(rows, cols, K, centred.vertices, centred.triangles, synth.default_parameters())
    rng = np.random.default_rng(5)
    start = np.array([0.01, 0.0, 0.8, 0.2, -0.4, 0.3])
    step = np.array([0.002, -0.001, 0.001, 0.0, 0.02, 0.0])
    truth = np.array([start + f * step for f in range(8)])
    frames = [synth.make_frame(o_plain.render, [(synth.rodrigues(p[3:6]), p[0:3])], rows, cols, rng, occluder=False)
              for p in truth]
    tracks = []
    for orc, model, flag in ((o_plain, plain, False), (o_centred, centred, True)):
        prm = TrackerParameters(evaluation_count=150, center_object_frame=flag)
        tr = ParticleTracker(OracleSensor(orc), 1, prm, seed=4, centers=model.centers)
        tr.initialize([truth[0] + [0.003, -0.002, 0.003, 0, 0, 0]])
        tracks.append(np.array([tr.track(img)[0] for img in frames]))
        assert tracks[-1].shape == (8, 12)                            # pose and velocity, as the reference's State
    err_plain = np.linalg.norm(tracks[0][:, :3] - truth[:, :3], axis=1)
    err_centred = np.linalg.norm(tracks[1][:, :3] - truth[:, :3], axis=1)
    # both tracks are in the model frame; the centred one is the tighter of the two (rotation noise about an
    # origin 6 cm off the object swings the whole object -- the reason the option exists); without the
    # conversion the centred track would be off by |c| = 6 cm
    assert err_centred[-4:].max() < 0.006 and err_plain[-4:].max() < 0.03, (err_plain, err_centred)
    with pytest.raises(ValueError):
        ParticleTracker(OracleSensor(o_centred), 1, TrackerParameters(center_object_frame=True))
    # the mean velocity the tracker reports has the sign and size of the motion (2 mm / frame at 30 Hz = 6 cm/s)
    assert 0.01 < np.mean(tracks[0][3:, 6]) < 0.2


def test_moving_average_blends_orientations_as_quaternions():
    from dbot_ros_b200.tracker import moving_average

    a = np.zeros((1, 12))
    b = np.zeros((1, 12))
    a[0, :3], b[0, :3] = [0.0, 0.0, 1.0], [0.1, 0.0, 1.0]
    a[0, 3:6], b[0, 3:6] = [0.0, 0.0, 0.2], [0.0, 0.0, 0.6]          # same axis: the blend is the mean angle
    a[0, 6:], b[0, 6:] = 1.0, 3.0
    m = moving_average(a, b, 0.5)
    assert np.allclose(m[0, :3], [0.05, 0.0, 1.0]) and np.allclose(m[0, 6:], 2.0)
    assert np.allclose(m[0, 3:6], [0.0, 0.0, 0.4], atol=2e-3)
    assert np.array_equal(moving_average(None, b, 0.3), b) and np.array_equal(moving_average(a, b, 1.0), b)
    # rotation vectors near +pi and -pi about the same axis are neighbours: the blend stays near pi instead of
    # passing through zero as a linear blend of the vectors would
    a[0, 3:6], b[0, 3:6] = [0.0, 0.0, 3.1], [0.0, 0.0, -3.1]
    m = moving_average(a, b, 0.5)
    assert abs(np.linalg.norm(m[0, 3:6]) - np.pi) < 0.05


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
    assert est_d.shape == est_h.shape == (12, 2, 12)                # pose + velocity per object
    # same noise, same sensor arithmetic: the two tracks agree far below the tracking error
    assert np.max(np.abs(est_d[:4] - est_h[:4])) < 1e-6, np.max(np.abs(est_d - est_h), axis=(1, 2))
    assert np.max(np.abs(est_d[:, :, :6] - est_h[:, :, :6])) < 5e-4, np.max(np.abs(est_d - est_h), axis=(1, 2))
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


# ---- the reference's offline dataset layout ----------------------------------------
def test_ground_truth_txt_layout_roundtrip_and_stamp_matching(tmp_path):
    """ground_truth.txt as TrackingDataset::StoreTextFile writes it (ref:source/dbot_ros/util/tracking_dataset.cpp:
    325-357): '<stamp> <state...>' per line; LoadTextFile's rule attaches a line to the frames within 0.02 s"""
    rng = np.random.default_rng(1)
    stamps = 1431.5 + np.arange(5) / 30.0
    states = rng.normal(size=(5, 24))                                 # two objects x 12
    io.save_ground_truth(tmp_path, stamps, states)
    text = (tmp_path / "ground_truth.txt").read_text().splitlines()
    assert len(text) == 5 and len(text[0].split()) == 25 and abs(float(text[2].split()[0]) - stamps[2]) < 1e-6
    s2, x2 = io.load_ground_truth(tmp_path)
    assert np.allclose(s2, stamps, atol=1e-9) and np.array_equal(x2, states)
    assert io.ground_truth_poses(x2, 2).shape == (5, 2, 6)
    assert np.array_equal(io.ground_truth_poses(x2, 2)[3, 1], states[3, 12:18])
    assert np.array_equal(io.ground_truth_poses(x2[:, :12].reshape(5, 12), 2)[3, 1], states[3, 6:12])   # 6 per object
    with pytest.raises(ValueError):
        io.ground_truth_poses(x2, 5)
    # frames at the ground-truth stamps +- jitter; one frame without ground truth (dropped line)
    frame_stamps = np.concatenate([stamps + 0.004, [stamps[-1] + 1.0 / 30.0]])
    idx = io.match_ground_truth(frame_stamps, stamps)
    assert idx.tolist() == [0, 1, 2, 3, 4, -1]
    # 21 ms after a line is beyond the 0.02 s bound, but 12 ms before the next one: that line is attached
    assert io.match_ground_truth(stamps + 0.021, stamps).tolist() == [1, 2, 3, 4, -1]
    (tmp_path / "bad.txt").write_text("12.0 1 2 3\n")
    with pytest.raises(ValueError):
        io.load_ground_truth(str(tmp_path / "bad.txt"))


def test_replay_of_a_dataset_directory_on_the_oracle(tmp_path):
    """save a synthetic sequence in the dataset layout, read it back, replay it through the host tracker on the
    oracle: frames whose ground-truth line is missing are tracked but not scored"""
    from dbot_ros_b200 import replay

    rows, cols = synth.resolution(8)
    K = synth.camera_matrix(8)
    V, T = synth.mesh_set("ico80")
    orc = oracle_lib.Oracle(rows, cols, K, V, T, synth.default_parameters())
    depth, stamps, states = replay.synthetic_sequence(orc.render, 1, rows, cols, 10, np.random.default_rng(3))
    io.save_dataset(tmp_path, depth, stamps, K, states)
    lines = (tmp_path / "ground_truth.txt").read_text().splitlines()
    (tmp_path / "ground_truth.txt").write_text("\n".join(lines[:4] + lines[5:]) + "\n")      # frame 4 unscored
    sensor = OracleSensor(orc)
    sensor.n_objects, sensor.max_sample_count = 1, 100
    out, est = replay.replay_dataset(str(tmp_path), sensor, 1, tracker="host",
                                     params=TrackerParameters(evaluation_count=100), seed=2)
    assert out["frames"] == 10 and out["frames_with_ground_truth"] == 9 and est.shape == (10, 1, 12)
    assert out["position_error_mm"]["last"] < 4.0 and out["fps"] > 0


@pytest.mark.gpu
def test_replay_synthetic_sequence_on_the_device_tracker():
    """20 frames of BASELINE config 2's shape (1000 particles, 640x480, 5120 triangles) through
    dbot_b200_filter_track: the estimate stays on the object and the loop runs far above the sensor's 30 Hz"""
    from dbot_ros_b200 import replay

    rows, cols = synth.resolution(1)
    K = synth.camera_matrix(1)
    V, T = synth.mesh_set("icosphere5k")
    model = ObjectModel(V, T, center_object_frame=True)
    plain = RbSensorBuilder(ObjectModel(V, T), CameraData(K, rows, cols), Parameters(sample_count=8)).build()
    sensor = RbSensorBuilder(model, CameraData(K, rows, cols), Parameters(sample_count=1000)).build()
    depth, stamps, states = replay.synthetic_sequence(plain.render, 1, rows, cols, 20, np.random.default_rng(1))
    prm = TrackerParameters(evaluation_count=1000, center_object_frame=True, moving_average_update_rate=0.8)
    out, est = replay.replay(sensor, depth, io.ground_truth_poses(states, 1), tracker="device", params=prm,
                             centers=model.centers, seed=1)
    assert out["position_error_mm"]["max"] < 6.0 and out["position_error_mm"]["last"] < 3.0, out
    assert out["fps"] > 200.0, out                                   # the reference's real-time bar is 30 Hz
    assert est.shape == (20, 1, 12)
