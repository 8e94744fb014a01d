"""Host-side mirror of the reference's sensor-construction interface (dbot_ros_b200/rb_sensor.py):
the YAML key surface, K / resolution scaling, depth-image decimation and pose composition -- the
logic that sits between the reference's node (ref:source/dbot_ros/tracker/particle_tracker_node.cpp:
84-203) and the C-ABI.  No GPU needed."""
import os

import numpy as np
import pytest
import yaml

from dbot_ros_b200 import CameraData, ObjectModel, Parameters, RbSensorBuilder, synth
from dbot_ros_b200.rb_sensor import B200KinectImageModel, rotation_matrices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parameters_from_the_reference_yaml_surface():
    p = Parameters.from_yaml(os.path.join(ROOT, "config", "particle_tracker.yaml"))
    assert p.use_gpu is True and p.sample_count == 2000          # gpu/sample_count when use_gpu
    assert p.delta_time == pytest.approx(1.0 / 30.0)             # fixed in the node, not in the YAML
    assert (p.occlusion.p_occluded_visible, p.occlusion.p_occluded_occluded,
            p.occlusion.initial_occlusion_prob) == (0.1, 0.7, 0.1)
    assert (p.kinect.tail_weight, p.kinect.model_sigma, p.kinect.sigma_factor) == (0.01, 0.003, 0.0014247)
    assert p.geometry_shader_file == "none" and p.use_custom_shaders is False  # accepted, ignored
    tree = yaml.safe_load(open(os.path.join(ROOT, "config", "particle_tracker.yaml")))
    tree["particle_filter"]["use_gpu"] = False
    q = Parameters.from_dict(tree)
    assert q.use_gpu is False and q.sample_count == 100          # cpu/sample_count otherwise
    del tree["particle_filter"]["observation"]["kinect"]["tail_weight"]
    with pytest.raises(KeyError):                                # checked read, like ri::read<T>
        Parameters.from_dict(tree)


def test_builder_refuses_the_cpu_branch():
    V, T = synth.mesh_set("ico80")
    b = RbSensorBuilder(ObjectModel(V, T), CameraData(synth.camera_matrix(8), 60, 80),
                        Parameters(use_gpu=False, sample_count=4))
    with pytest.raises(RuntimeError, match="no CPU sensor"):
        b.build()


def test_camera_data_downsampling():
    cam = yaml.safe_load(open(os.path.join(ROOT, "config", "camera.yaml")))
    ds = cam["downsampling_factor"]
    c = CameraData.from_native(synth.K_VGA, cam["resolution"]["width"], cam["resolution"]["height"], ds)
    assert (c.rows, c.cols, c.pixels) == (60, 80, 4800)
    assert np.allclose(c.camera_matrix, [[72.5, 0, 40.0], [0, 72.5, 30.0], [0, 0, 1]])  # rows 0-1 / 8
    img = np.arange(480 * 640, dtype=np.float32).reshape(480, 640)
    v = c.depth_image_vector(img)
    assert v.dtype == np.float32 and v.shape == (4800,)
    for row, col in ((0, 0), (3, 7), (59, 79)):                  # v[row*n_cols+col] = img(row*ds, col*ds)
        assert v[row * 80 + col] == img[row * ds, col * ds]


def test_object_model_centering():
    V = np.array([[0, 0, 0], [2, 0, 0], [0, 4, 0], [0, 0, 6]], dtype=np.float64) + 10.0
    T = np.array([[0, 1, 2], [0, 1, 3]], dtype=np.int32)
    m = ObjectModel([V], [T], center_object_frame=True)
    assert np.allclose(m.centers[0], V.mean(axis=0)) and np.allclose(m.vertices[0].mean(axis=0), 0, atol=1e-6)
    m2 = ObjectModel([V], [T], center_object_frame=False)
    assert np.allclose(m2.vertices[0], V) and m2.count_parts == 1
    with pytest.raises(ValueError):
        ObjectModel([V], [T, T])


def test_rotation_matrices_and_compose():
    rng = np.random.default_rng(0)
    rv = rng.normal(size=(7, 3))
    rv[3] = 0.0
    R = rotation_matrices(rv)
    for k in range(7):
        assert np.allclose(R[k], synth.rodrigues(rv[k]), atol=1e-14)
        assert np.allclose(R[k] @ R[k].T, np.eye(3), atol=1e-14)
    # compose(): pose = delta (x) default, without touching the GPU library
    s = B200KinectImageModel.__new__(B200KinectImageModel)
    s.n_objects = 2
    s._default_poses = np.array([[0.1, 0.0, 0.8, 0.2, -0.1, 0.3], [-0.1, 0.05, 0.9, 0.0, 0.0, 0.0]])
    deltas = np.zeros((3, 24))                                   # FreeFloatingRigidBodiesState: 12 per body
    deltas[1, 0:3] = (0.01, 0.02, 0.03)
    deltas[2, 12 + 3:12 + 6] = (0.0, 0.0, np.pi / 2)
    out = s.compose(deltas)
    assert out.shape == (3, 2, 12) and out.dtype == np.float32
    R0 = synth.rodrigues(s._default_poses[0, 3:6])
    assert np.allclose(out[0, 0, :9].reshape(3, 3), R0, atol=1e-7) and np.allclose(out[0, 0, 9:], (0.1, 0.0, 0.8))
    assert np.allclose(out[1, 0, 9:], (0.11, 0.02, 0.83), atol=1e-7)
    Rz = synth.rodrigues([0, 0, np.pi / 2])
    assert np.allclose(out[2, 1, :9].reshape(3, 3), Rz, atol=1e-7)
    assert np.allclose(s.compose(deltas.reshape(3, 2, 12)[:, :, :6].reshape(3, 12)), out)  # 6-number pose vectors
    with pytest.raises(ValueError):
        s.compose(np.zeros((3, 14)))
    s._h = None  # nothing to destroy


def test_shard_scan_agrees_with_the_python_partition():
    """dbot_b200_shard_scan (the one pass over a sharded call's parent vector, host code in the library) against
    shard_bounds / shard_owner: same owners, first remote child, range errors"""
    import ctypes

    from dbot_ros_b200 import abi
    from dbot_ros_b200.sharded import shard_bounds, shard_owner

    lib = abi.load_library()
    rng = np.random.default_rng(3)
    for n, n_slots, world in [(10, 10, 2), (501, 501, 2), (7, 3, 8), (1000, 997, 8), (5, 1, 4), (8000, 8000, 8)]:
        child_owner = np.array([shard_owner(i, n, world) for i in range(n)])
        for r in range(world):
            lo, hi = shard_bounds(n, world, r)
            assert np.all(child_owner[lo:hi] == r)
        # every parent on its child's rank
        idx = np.empty(n, dtype=np.int32)
        ok = True
        for i in range(n):
            lo, hi = shard_bounds(n_slots, world, child_owner[i])
            if hi == lo:
                ok = False
                idx[i] = 0
            else:
                idx[i] = rng.integers(lo, hi)
        first = ctypes.c_int32(-2)
        rc = lib.dbot_b200_shard_scan(abi.ptr(idx), n, n_slots, world, ctypes.byref(first))
        want = [i for i in range(n) if n_slots > 1 and shard_owner(int(idx[i]), n_slots, world) != child_owner[i]]
        assert rc == (1 if want else 0) and first.value == (want[0] if want else -1)
        if ok and n_slots > 1:
            assert rc == 0
        # one child with a parent somewhere else
        idx2 = rng.integers(0, n_slots, size=n).astype(np.int32)
        rc = lib.dbot_b200_shard_scan(abi.ptr(idx2), n, n_slots, world, ctypes.byref(first))
        want = [i for i in range(n) if n_slots > 1 and shard_owner(int(idx2[i]), n_slots, world) != child_owner[i]]
        assert rc == (1 if want else 0) and first.value == (want[0] if want else -1)
        idx2[n // 2] = n_slots
        assert lib.dbot_b200_shard_scan(abi.ptr(idx2), n, n_slots, world, None) == -7
        idx2[n // 2] = -1
        assert lib.dbot_b200_shard_scan(abi.ptr(idx2), n, n_slots, world, None) == -7
