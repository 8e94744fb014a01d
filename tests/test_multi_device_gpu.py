"""One handle, several GPUs (include/dbot_b200.h: n_devices / device_ids / gather_mode; SURVEY.md section 8e and
BASELINE config 4): sharding particles over GPUs inside the library must not change a single bit -- log-weights,
occlusion images and `indices` semantics are compared with a single-GPU handle, with resampled parents that
force occlusion slots to migrate between the shards.  On a one-GPU box the shards share GPU 0 (fused gather
only; NCCL wants one rank per device); with two or more GPUs both gather modes run across real devices."""
import ctypes as C

import numpy as np
import pytest

from dbot_ros_b200 import CameraData, DbotB200Error, ObjectModel, Parameters, RbSensorBuilder, synth

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch

    return torch.cuda.device_count()


def build(n, device_ids=None, gather_mode=0, mesh="uvsphere480", ds=4):
    rows, cols = synth.resolution(ds)
    K = synth.camera_matrix(ds)
    V, T = synth.mesh_set(mesh)
    b = RbSensorBuilder(ObjectModel(V, T), CameraData(K, rows, cols), Parameters(sample_count=n),
                        device_ids=device_ids, gather_mode=gather_mode)
    return b.build(), rows, cols, len(V)


def run_sequence(sensor, frames, poses, parents):
    out, occ = [], []
    for f in range(len(frames)):
        idx = parents[f].copy()
        sensor.set_observation(frames[f])
        ev = sensor.loglikes_poses(poses[f], idx, False)
        assert np.array_equal(idx, parents[f])              # update=false leaves indices alone
        up = sensor.loglikes_poses(poses[f], idx, True)
        assert np.array_equal(idx, np.arange(len(idx)))
        out.append((ev, up))
        occ.append([sensor.occlusion(s) for s in (0, len(idx) // 2, len(idx) - 1)])
    return out, occ


def scenario(n, n_frames=4, mesh="uvsphere480", ds=4):
    ref, rows, cols, n_obj = build(n, mesh=mesh, ds=ds)
    rng = synth.rng_for(4)
    gt = synth.ground_truth_poses(n_obj, rng)
    frames = [synth.make_frame(ref.render, gt, rows, cols, rng) for _ in range(n_frames)]
    poses = [synth.particle_poses(gt, n, rng) for _ in range(n_frames)]
    parents = [np.zeros(n, np.int32)] + [rng.integers(0, n, size=n).astype(np.int32) for _ in range(n_frames - 1)]
    parents[-1] = np.sort(parents[-1])                      # systematic resampling gives sorted parents
    return ref, frames, poses, parents


def check_identical(device_ids, gather_mode, n=61):
    ref, frames, poses, parents = scenario(n)
    want, want_occ = run_sequence(ref, frames, poses, parents)
    sharded, *_ = build(n, device_ids=device_ids, gather_mode=gather_mode)
    assert sharded.device_count == len(device_ids)
    got, got_occ = run_sequence(sharded, frames, poses, parents)
    for (a, b), (c, d) in zip(want, got):
        assert np.array_equal(a, c) and np.array_equal(b, d)
    for a, b in zip(want_occ, got_occ):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert sharded.n_slots == n and sharded.migrated_slots > 0   # random parents crossed the shard boundaries
    assert sharded.free_slots + n == 2 * len(device_ids) * ((n + len(device_ids) - 1) // len(device_ids))
    sharded.reset()
    assert sharded.n_slots == 1 and sharded.observation_time == 0.0
    sharded.close()
    ref.close()


def test_shards_sharing_one_gpu_fused_gather():
    """3 shards on GPU 0, log-weights stored by the kernel epilogue into every shard's gathered vector"""
    check_identical([0, 0, 0], gather_mode=1)


def test_fewer_particles_than_shards_and_errors():
    sensor, rows, cols, n_obj = build(8, device_ids=[0, 0, 0, 0], gather_mode=1)
    ref, *_ = build(8)
    rng = synth.rng_for(5)
    gt = synth.ground_truth_poses(n_obj, rng)
    frame = synth.make_frame(ref.render, gt, rows, cols, rng)
    poses = synth.particle_poses(gt, 8, rng)
    for s in (sensor, ref):
        s.set_observation(frame)
    for n in (2, 8, 1, 3):          # shards without particles must stay in step with the others
        ia = np.zeros(n, np.int32) if n == 2 else np.arange(n, dtype=np.int32) % sensor.n_slots
        ib = ia.copy()
        assert np.array_equal(sensor.loglikes_poses(poses[:n], ia, True), ref.loglikes_poses(poses[:n], ib, True))
        assert sensor.n_slots == n == ref.n_slots
        assert np.array_equal(sensor.occlusion(n - 1), ref.occlusion(n - 1))
        for s in (sensor, ref):
            s.set_observation(frame)
    with pytest.raises(DbotB200Error) as e:           # device-pointer entry points want one GPU
        sensor.synchronize()
    assert e.value.code == -12
    with pytest.raises(DbotB200Error):
        sensor.loglikes_poses(poses[:3], np.full(3, 3, np.int32), False)
    with pytest.raises(DbotB200Error):                 # NCCL wants one rank per device
        build(8, device_ids=[0, 0], gather_mode=0)


@pytest.mark.parametrize("gather_mode", [0, 1])
def test_two_gpus_bit_identical_to_one(gather_mode):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    check_identical([0, 1], gather_mode)


def test_all_gpus_of_the_box_config4_shape():
    """BASELINE config 4's shape at a size the test can afford: every GPU of the box, 500 particles each at
    640x480 with the 5120-triangle mesh, shard-local parents; shard-invariance against a single GPU"""
    g = n_gpus()
    if g < 2:
        pytest.skip("needs 2 GPUs")
    n = 500 * g
    ref, frames, poses, _ = scenario(n, n_frames=2, mesh="icosphere5k", ds=1)
    sharded, *_ = build(n, device_ids=list(range(g)), gather_mode=0, mesh="icosphere5k", ds=1)
    per = (n + g - 1) // g
    local = (np.arange(n) // per * per + np.random.default_rng(1).integers(0, per, size=n)).clip(max=n - 1).astype(np.int32)
    parents = [np.zeros(n, np.int32), local]
    want, _ = run_sequence(ref, frames, poses, parents)
    got, _ = run_sequence(sharded, frames, poses, parents)
    for (a, b), (c, d) in zip(want, got):
        assert np.array_equal(a, c) and np.array_equal(b, d)
    assert sharded.migrated_slots == 0
