"""include/dbot_b200_adapter.hpp (the C++ dbot::RbSensor<State> a maintainer would add)
compiles against stand-ins of dbot's headers and links with libdbot_b200.so.  Without a GPU it turns the
C-ABI's error code into a C++ exception instead of falling back to anything; with one (-m gpu) a C++ driver
calls reset / set_observation / loglikes through the dbot::RbSensor base class exactly as dbot's filter does
and its log-weights are compared with the oracle's."""
import os
import struct
import subprocess

import numpy as np
import pytest

from dbot_ros_b200 import abi, synth
from dbot_ros_b200.rb_sensor import rotation_matrices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path, source):
    exe = str(tmp_path / os.path.splitext(source)[0])
    libdir = os.path.dirname(abi.LIB_PATH)
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "tests", "stubs"), os.path.join(ROOT, "tests", "stubs", source),
                    "-L", libdir, "-ldbot_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


def test_adapter_compiles_links_and_reports_errors(tmp_path):
    exe = build(tmp_path, "adapter_check.cpp")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "created" in r.stdout or "no CPU fallback" in r.stdout, r.stdout
    build(tmp_path, "adapter_run.cpp")          # the GPU driver must at least compile and link everywhere


def write_scene(path, rows, cols, K, V, T, default, frames, deltas, parents):
    """binary scene for tests/stubs/adapter_run.cpp; poses as position + rotation matrix in double"""
    def pose_bytes(p6):
        return np.asarray(p6[:3], np.float64).tobytes() + rotation_matrices(np.asarray(p6[3:6])).astype(np.float64).tobytes()

    n_obj, n = len(V), deltas[0].shape[0]
    with open(path, "wb") as f:
        f.write(struct.pack("<5i", rows, cols, n_obj, n, len(frames)))
        f.write(np.asarray(K, np.float64).tobytes())
        for v, t in zip(V, T):
            f.write(struct.pack("<2i", len(v), len(t)))
            f.write(np.asarray(v, np.float64).tobytes())
            f.write(np.asarray(t, np.int32).tobytes())
        for o in range(n_obj):
            f.write(pose_bytes(default[o]))
        for fr, d, par in zip(frames, deltas, parents):
            f.write(np.asarray(fr, np.float32).tobytes())
            for i in range(n):
                for o in range(n_obj):
                    f.write(pose_bytes(d[i, o]))
            f.write(np.asarray(par, np.int32).tobytes())


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [[], [0, 1]])
def test_adapter_drives_the_library_like_dbots_filter(tmp_path, devices):
    """set_observation + loglikes(update=false/true) over two frames with resampled parents, through the C++
    adapter; single GPU, and particle-sharded over two GPUs behind the same C++ object when the box has them"""
    import torch

    import oracle_lib
    from dbot_ros_b200.rb_sensor import compose_poses

    if len(devices) > torch.cuda.device_count():
        pytest.skip("needs %d GPUs" % len(devices))
    exe = build(tmp_path, "adapter_run.cpp")
    rows, cols = synth.resolution(4)
    K = synth.camera_matrix(4)
    V, T = synth.mesh_set("three20k")
    n, n_obj = 40, len(V)
    orc = oracle_lib.Oracle(rows, cols, K, V, T, synth.default_parameters(), cull=1)
    rng = synth.rng_for(6)
    gt = synth.ground_truth_poses(n_obj, rng)
    from dbot_ros_b200.tracker import rotation_vector

    default = np.array([np.concatenate([t, rotation_vector(R)]) for R, t in gt])
    frames = [synth.make_frame(orc.render, gt, rows, cols, rng) for _ in range(2)]
    deltas = [np.concatenate([rng.normal(scale=0.004, size=(n, n_obj, 3)), rng.normal(scale=0.04, size=(n, n_obj, 3))],
                             axis=2) for _ in range(2)]
    parents = [np.zeros(n, np.int32), np.sort(rng.integers(0, n, size=n)).astype(np.int32)]
    scene, out = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(scene, rows, cols, K, V, T, default, frames, deltas, parents)
    r = subprocess.run([exe, scene, out] + [str(d) for d in devices], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr
    got = np.fromfile(out, dtype=np.float64).reshape(2, 2, n)
    for f in range(2):
        poses = compose_poses(deltas[f].reshape(n, -1), default)
        orc.set_observation(frames[f])
        idx = parents[f].copy()
        want_ev = orc.loglikes_poses(poses, idx, False)
        want_up = orc.loglikes_poses(poses, idx, True)
        for g, w in ((got[f, 0], want_ev), (got[f, 1], want_up)):
            assert np.max(np.abs(g - w) / np.maximum(np.abs(w), 1.0)) < 1e-4
