"""GPU parity: libdbot_b200.so (through the C-ABI / host mirror) against the CPU oracle on the same
seeded inputs.  Bars: rendered depth bit-exact (integer coverage + identical fp32 op order);
log-weights within 1e-4 relative (BASELINE.json north_star); occlusion images within 2e-5."""
import numpy as np
import pytest

import oracle_lib
from dbot_ros_b200 import (CameraData, DbotB200Error, ObjectModel, Parameters, RbSensorBuilder,
                           synth)

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4   # north_star: log-weights within 1e-4 relative of the CPU reference
OCC_TOL = 2e-5  # fp32 posterior: ex2.approx argument rounding at |t|^2 ~ 20


def rel_err(got, want):
    return np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) if len(want) else 0.0


def make_pair(mesh, ds, max_n, K=None, cull=True):
    """cull=True: the library's default (back faces of meshes proven closed are skipped) against the
    oracle with the same rule; cull=False: every triangle drawn on both sides, as upstream dbot."""
    rows, cols = synth.resolution(ds)
    K = synth.camera_matrix(ds) if K is None else K
    V, T = synth.mesh_set(mesh) if isinstance(mesh, str) else mesh
    sensor = RbSensorBuilder(ObjectModel(V, T), CameraData(K, rows, cols),
                             Parameters(sample_count=max_n, cull_backfaces=cull)).build()
    orc = oracle_lib.Oracle(rows, cols, K, V, T, synth.default_parameters(), cull=1 if cull else 0)
    for o in range(len(V)):
        assert sensor.cull_sign(o) == (orc.cull_sign(o) if cull else 0)
    return sensor, orc, rows, cols, len(V)


@pytest.mark.parametrize("cull", [True, False])
@pytest.mark.parametrize("mesh,ds", [("uvsphere480", 4), ("icosphere5k", 1), ("three20k", 2),
                                     ("ico80", 8)])
def test_render_bit_exact(mesh, ds, cull):
    sensor, orc, rows, cols, n_obj = make_pair(mesh, ds, 8, cull=cull)
    rng = np.random.default_rng(5)
    gt = synth.ground_truth_poses(n_obj, rng)
    for pose in synth.particle_poses(gt, 6, rng, sigma_t=0.03, sigma_r=0.5):
        got = sensor.render(pose)
        want = orc.render(pose)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert np.isfinite(got).sum() > 0


@pytest.mark.parametrize("mesh,ds,n", [("uvsphere480", 4, 200), ("icosphere5k", 1, 48),
                                       ("three20k", 1, 24), ("three20k", 4, 64)])
def test_loglikes_and_update_parity(mesh, ds, n):
    """BASELINE configs 1, 2 and 5 at oracle-friendly particle counts"""
    sensor, orc, rows, cols, n_obj = make_pair(mesh, ds, n)
    rng = synth.rng_for(2)
    gt = synth.ground_truth_poses(n_obj, rng)
    frame = synth.make_frame(orc.render, gt, rows, cols, rng)
    poses = synth.particle_poses(gt, n, rng)
    sensor.set_observation(frame)
    orc.set_observation(frame)
    ia = np.zeros(n, dtype=np.int32)
    ib = np.zeros(n, dtype=np.int32)
    # update=False must not touch state
    got = sensor.loglikes_poses(poses, ia, update=False)
    want = orc.loglikes_poses(poses, ib, update=False)
    assert rel_err(got, want) < REL_TOL
    assert sensor.n_slots == 1 and np.all(ia == 0)
    again = sensor.loglikes_poses(poses, ia, update=False)
    assert np.array_equal(got, again)  # deterministic reduction
    got = sensor.loglikes_poses(poses, ia, update=True)
    want = orc.loglikes_poses(poses, ib, update=True)
    assert rel_err(got, want) < REL_TOL
    assert np.array_equal(ia, np.arange(n)) and np.array_equal(ib, ia)
    assert sensor.n_slots == n == orc.n_slots
    for slot in (0, n // 2, n - 1):
        assert np.max(np.abs(sensor.occlusion(slot) - orc.occlusion(slot))) < OCC_TOL


def test_multi_frame_with_resampling():
    """5 frames; parents drawn at random each frame (what the filter's resampling does to
    `indices`); log-weights and occlusion state must track the oracle frame after frame."""
    n = 96
    sensor, orc, rows, cols, n_obj = make_pair("uvsphere480", 4, n)
    rng = synth.rng_for(1)
    gt = synth.ground_truth_poses(n_obj, rng)
    ia = np.zeros(n, dtype=np.int32)
    for f in range(5):
        frame = synth.make_frame(orc.render, gt, rows, cols, rng)
        poses = synth.particle_poses(gt, n, rng)
        sensor.set_observation(frame)
        orc.set_observation(frame)
        assert abs(sensor.observation_time - orc.observation_time) < 1e-12
        if f > 0:
            ia = rng.integers(0, n, size=n).astype(np.int32)
        ib = ia.copy()
        if f == 2:  # a frame with an extra non-updating call first (second sampling block)
            g0 = sensor.loglikes_poses(poses, ia, update=False)
            w0 = orc.loglikes_poses(poses, ib, update=False)
            assert rel_err(g0, w0) < REL_TOL
        got = sensor.loglikes_poses(poses, ia, update=True)
        want = orc.loglikes_poses(poses, ib, update=True)
        assert rel_err(got, want) < REL_TOL, "frame %d" % f
        for slot in (0, n - 1):
            assert np.max(np.abs(sensor.occlusion(slot) - orc.occlusion(slot))) < 2 * OCC_TOL
    # occlusion read-back between frames is "as of the current observation time"
    frame = synth.make_frame(orc.render, gt, rows, cols, rng)
    sensor.set_observation(frame)
    orc.set_observation(frame)
    assert np.max(np.abs(sensor.occlusion(3) - orc.occlusion(3))) < 2 * OCC_TOL
    sensor.reset()
    orc.reset()
    assert sensor.n_slots == 1 and sensor.observation_time == 0.0
    assert np.allclose(sensor.occlusion(0), 0.1)


def test_multi_chunk_mesh_and_big_window():
    """icosphere level 5 (10242 vertices / 20480 triangles -> several TMA-staged chunks) held
    25 cm from the camera: the window covers most of the frame -> several z-tile bands."""
    V, T = synth.icosphere(5, radius=0.1, seed=9)
    sensor, orc, rows, cols, _ = make_pair(([V], [T]), 1, 8)
    rng = np.random.default_rng(11)
    gt = [(synth.rodrigues([0.3, -0.2, 0.9]), np.array([0.02, -0.01, 0.25]))]
    pose = synth.pack_pose(*gt[0])
    got = sensor.render(pose)
    want = orc.render(pose)
    assert np.isfinite(want).sum() > 0.4 * rows * cols
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    frame = synth.make_frame(orc.render, gt, rows, cols, rng)
    poses = synth.particle_poses(gt, 8, rng)
    sensor.set_observation(frame)
    orc.set_observation(frame)
    ia = np.zeros(8, dtype=np.int32)
    ib = np.zeros(8, dtype=np.int32)
    assert rel_err(sensor.loglikes_poses(poses, ia, True), orc.loglikes_poses(poses, ib, True)) < REL_TOL
    assert np.max(np.abs(sensor.occlusion(5) - orc.occlusion(5))) < OCC_TOL


def test_edge_cases():
    n = 16
    sensor, orc, rows, cols, _ = make_pair("ico80", 4, n)
    rng = np.random.default_rng(3)
    gt = [(np.eye(3), np.array([0.0, 0.0, 0.8]))]
    frame = synth.make_frame(orc.render, gt, rows, cols, rng)
    sensor.set_observation(frame)
    orc.set_observation(frame)
    poses = synth.particle_poses(gt, n, rng)
    # off-screen, behind the camera, straddling the image border, straddling z = 0
    poses[0, 0, 9:] = (5.0, 0.0, 0.8)
    poses[1, 0, 9:] = (0.0, 0.0, -0.8)
    poses[2, 0, 9:] = (-0.42, 0.3, 0.8)
    poses[3, 0, 9:] = (0.0, 0.0, 0.05)
    # far away: the whole object falls on a few pixels, then on a single column / a single pixel
    poses[4, 0, 9:] = (0.0, 0.0, 12.0)
    poses[5, 0, 9:] = (0.3, 0.1, 40.0)
    poses[6, 0, 9:] = (0.0, 0.0, 150.0)
    ia = np.zeros(n, dtype=np.int32)
    ib = np.zeros(n, dtype=np.int32)
    got = sensor.loglikes_poses(poses, ia, False)
    want = orc.loglikes_poses(poses, ib, False)
    assert got[0] == 0.0 and got[1] == 0.0
    assert rel_err(got, want) < REL_TOL
    # empty call
    e = np.zeros(0, dtype=np.int32)
    assert sensor.loglikes_poses(np.zeros((0, 1, 12), np.float32), e, False).size == 0
    # all-NaN frame: every pixel is "no data" -> log-weight 0, occlusion only propagates
    sensor.set_observation(np.full((rows, cols), np.nan, np.float32))
    orc.set_observation(np.full((rows, cols), np.nan, np.float32))
    got = sensor.loglikes_poses(poses, ia, True)
    want = orc.loglikes_poses(poses, ib, True)
    assert np.all(got == 0.0) and np.all(want == 0.0)
    assert np.max(np.abs(sensor.occlusion(4) - orc.occlusion(4))) < OCC_TOL
    # errors: too many particles, parent slot out of range (mirrors dbot's GPU model exceptions)
    with pytest.raises(DbotB200Error):
        sensor.loglikes_poses(np.zeros((n + 1, 1, 12), np.float32), np.zeros(n + 1, np.int32), False)
    bad = np.full(n, n, dtype=np.int32)
    with pytest.raises(DbotB200Error):
        sensor.loglikes_poses(poses, bad, False)


def test_odd_resolution_pitch():
    """cols not a multiple of 4 -> padded state pitch"""
    rows, cols = 61, 83
    K = np.array([[70.0, 0.3, 41.0], [0, 71.0, 30.0], [0, 0, 1]])
    V, T = synth.mesh_set("ico80")
    sensor = RbSensorBuilder(ObjectModel(V, T), CameraData(K, rows, cols), Parameters(sample_count=8)).build()
    orc = oracle_lib.Oracle(rows, cols, K, V, T, synth.default_parameters(), cull=1)
    rng = np.random.default_rng(8)
    gt = [(synth.rodrigues([0.1, 0.2, 0.3]), np.array([0.01, 0.0, 0.6]))]
    frame = synth.make_frame(orc.render, gt, rows, cols, rng)
    poses = synth.particle_poses(gt, 8, rng)
    for s in (sensor, orc):
        s.set_observation(frame)
    ia = np.zeros(8, dtype=np.int32)
    ib = np.zeros(8, dtype=np.int32)
    assert np.array_equal(sensor.render(poses[0]).view(np.uint32), orc.render(poses[0]).view(np.uint32))
    assert rel_err(sensor.loglikes_poses(poses, ia, True), orc.loglikes_poses(poses, ib, True)) < REL_TOL
    assert np.max(np.abs(sensor.occlusion(7) - orc.occlusion(7))) < OCC_TOL


def test_full_size_properties():
    """BASELINE config 2 at full size (1000 particles, 640x480, 5120 triangles): properties that
    need no oracle -- shard invariance (what the multi-GPU split relies on), permutation
    equivariance, determinism -- plus an oracle spot check on 16 of the particles."""
    n = 1000
    sensor, orc, rows, cols, n_obj = make_pair("icosphere5k", 1, n)
    rng = synth.rng_for(2)
    gt = synth.ground_truth_poses(n_obj, rng)
    frame = synth.make_frame(sensor.render, gt, rows, cols, rng)
    poses = synth.particle_poses(gt, n, rng)
    sensor.set_observation(frame)
    idx = np.zeros(n, dtype=np.int32)
    whole = sensor.loglikes_poses(poses, idx, False)
    halves = np.concatenate([sensor.loglikes_poses(poses[:500], idx[:500].copy(), False),
                             sensor.loglikes_poses(poses[500:], idx[500:].copy(), False)])
    assert np.array_equal(whole, halves)
    perm = rng.permutation(n)
    assert np.array_equal(sensor.loglikes_poses(poses[perm], idx, False), whole[perm])
    upd = sensor.loglikes_poses(poses, idx, True)
    assert np.array_equal(upd, whole)            # update does not change the weights of its call
    assert np.array_equal(idx, np.arange(n))
    # after the update with dt = 0 the second evaluation uses the posterior occlusion
    orc.set_observation(frame)
    pick = rng.choice(n, 16, replace=False)
    io = np.zeros(16, dtype=np.int32)
    want = orc.loglikes_poses(poses[pick], io, True)
    assert rel_err(whole[pick], want) < REL_TOL
    second = sensor.loglikes_poses(poses[pick], pick.astype(np.int32), False)
    want2 = orc.loglikes_poses(poses[pick], np.arange(16, dtype=np.int32), False)
    assert rel_err(second, want2) < REL_TOL


def test_config3_roofline_size_properties():
    """BASELINE config 3 (10 000 particles, 640x480, 5120 triangles; 24.6 GB of occlusion state):
    too big for the oracle, so size-independent properties -- update leaves the weights of its own
    call unchanged, chunked evaluation equals whole evaluation bit for bit, identity parents after
    the update -- plus an oracle spot check on 8 particles drawn from all over the index range."""
    n = 10000
    sensor, orc, rows, cols, n_obj = make_pair("icosphere5k", 1, n)
    rng = synth.rng_for(3)
    gt = synth.ground_truth_poses(n_obj, rng)
    frame = synth.make_frame(sensor.render, gt, rows, cols, rng)
    poses = synth.particle_poses(gt, n, rng)
    sensor.set_observation(frame)
    idx = np.zeros(n, dtype=np.int32)
    whole = sensor.loglikes_poses(poses, idx, False)
    parts = np.concatenate([sensor.loglikes_poses(poses[a:a + 2500], idx[a:a + 2500].copy(), False)
                            for a in range(0, n, 2500)])
    assert np.array_equal(whole, parts)
    upd = sensor.loglikes_poses(poses, idx, True)
    assert np.array_equal(upd, whole) and np.array_equal(idx, np.arange(n)) and sensor.n_slots == n
    pick = np.sort(rng.choice(n, 8, replace=False))
    orc.set_observation(frame)
    io = np.zeros(8, dtype=np.int32)
    want = orc.loglikes_poses(poses[pick], io, True)
    assert rel_err(whole[pick], want) < REL_TOL
    # second frame: parents = a resampling of all 10 000 slots; spot check against the oracle,
    # whose slots 0..7 now hold the picked particles' images
    frame2 = synth.make_frame(sensor.render, gt, rows, cols, rng)
    sensor.set_observation(frame2)
    orc.set_observation(frame2)
    parents = rng.integers(0, n, size=n).astype(np.int32)
    parents[pick] = pick                      # the picked particles keep their own history
    got2 = sensor.loglikes_poses(poses, parents, True)
    want2 = orc.loglikes_poses(poses[pick], np.arange(8, dtype=np.int32), True)
    assert rel_err(got2[pick], want2) < REL_TOL
    for k in (0, 7):
        assert np.max(np.abs(sensor.occlusion(int(pick[k])) - orc.occlusion(k))) < 2 * OCC_TOL


def test_observation_paths_pinned_and_staged():
    """set_observation copies straight from page-locked caller memory and stages pageable memory
    asynchronously; both must behave like a synchronous copy (the caller may reuse its buffer)."""
    import torch

    n = 32
    sensor, orc, rows, cols, n_obj = make_pair("uvsphere480", 4, n)
    rng = synth.rng_for(1)
    gt = synth.ground_truth_poses(n_obj, rng)
    poses = synth.particle_poses(gt, n, rng)
    f1 = synth.make_frame(orc.render, gt, rows, cols, rng)
    f2 = synth.make_frame(orc.render, gt, rows, cols, rng)
    orc.set_observation(f1)
    orc.set_observation(f2)                          # two frames in a row: time advances twice
    want = orc.loglikes_poses(poses, np.zeros(n, np.int32), False)
    buf = f1.copy()                                  # pageable
    sensor.set_observation(buf)
    buf[:] = np.nan                                  # reuse at once: the staged copy must be unaffected
    pinned = torch.from_numpy(f2.copy()).pin_memory().numpy()
    sensor.set_observation(pinned)
    pinned[:] = np.nan
    got = sensor.loglikes_poses(poses, np.zeros(n, np.int32), False)
    assert abs(sensor.observation_time - orc.observation_time) < 1e-12
    assert rel_err(got, want) < REL_TOL
    sensor.set_observation(f1)
    sensor.set_observation(f2)                       # staged twice in a row
    orc.set_observation(f1)
    orc.set_observation(f2)
    assert rel_err(sensor.loglikes_poses(poses, np.zeros(n, np.int32), False),
                   orc.loglikes_poses(poses, np.zeros(n, np.int32), False)) < REL_TOL


def test_slot_adoption_patterns_and_pool_accounting():
    """The update plan (last child adopts the parent's slot, the others copy into free slots) must
    be invisible: identity parents, one parent for everybody, shrinking and growing particle sets,
    repeated and unused parents -- log-weights and occlusion images track the oracle, and no
    physical slot is ever lost (free + live == pool)."""
    cap = 48
    sensor, orc, rows, cols, n_obj = make_pair("ico80", 8, cap)
    rng = np.random.default_rng(21)
    gt = synth.ground_truth_poses(n_obj, rng)
    n_prev = 1
    schedule = [("same", 32), ("identity", 32), ("identity", 32), ("random", 32), ("shrink", 7),
                ("random", 48), ("reverse", 48), ("half", 48), ("identity", 48), ("same", 5)]
    for f, (kind, n) in enumerate(schedule):
        frame = synth.make_frame(orc.render, gt, rows, cols, rng)
        poses = synth.particle_poses(gt, n, rng, sigma_t=0.01, sigma_r=0.1)
        if kind == "same":
            par = np.full(n, rng.integers(0, n_prev), dtype=np.int32)
        elif kind == "identity":
            par = np.arange(n, dtype=np.int32)
        elif kind == "reverse":
            par = np.arange(n, dtype=np.int32)[::-1].copy()
        elif kind == "half":
            par = (np.arange(n, dtype=np.int32) // 2) * 2
        else:
            par = rng.integers(0, n_prev, size=n).astype(np.int32)
        ia, ib = par.copy(), par.copy()
        sensor.set_observation(frame)
        orc.set_observation(frame)
        got = sensor.loglikes_poses(poses, ia, True)
        want = orc.loglikes_poses(poses, ib, True)
        assert rel_err(got, want) < REL_TOL, (f, kind)
        assert sensor.n_slots == n and sensor.free_slots + n == 2 * cap, (f, kind)
        for slot in {0, n // 2, n - 1}:
            assert np.max(np.abs(sensor.occlusion(slot) - orc.occlusion(slot))) < 2 * OCC_TOL, (f, kind, slot)
        n_prev = n
    sensor.reset()
    assert sensor.n_slots == 1 and sensor.free_slots == 2 * cap - 1


def test_lazy_propagation_rebase():
    """With 5 s between frames the composed propagation factor falls below 2^-60 after 17 updates
    and the stored values are rebased; the read-back and the log-weights must not notice."""
    n = 12
    rows, cols = synth.resolution(8)
    K = synth.camera_matrix(8)
    V, T = synth.mesh_set("ico80")
    prm = dict(synth.default_parameters(), delta_time=5.0)
    sensor = RbSensorBuilder(ObjectModel(V, T), CameraData(K, rows, cols),
                             Parameters(sample_count=n, delta_time=5.0)).build()
    orc = oracle_lib.Oracle(rows, cols, K, V, T, prm, cull=1)
    rng = np.random.default_rng(33)
    gt = synth.ground_truth_poses(1, rng)
    idx = np.zeros(n, dtype=np.int32)
    for f in range(40):
        frame = synth.make_frame(orc.render, gt, rows, cols, rng)
        poses = synth.particle_poses(gt, n, rng, sigma_t=0.01, sigma_r=0.1)
        par = idx.copy() if f == 0 else rng.integers(0, n, size=n).astype(np.int32)
        ia, ib = par.copy(), par.copy()
        sensor.set_observation(frame)
        orc.set_observation(frame)
        got = sensor.loglikes_poses(poses, ia, True)
        want = orc.loglikes_poses(poses, ib, True)
        assert rel_err(got, want) < REL_TOL, f
        assert np.max(np.abs(sensor.occlusion(n - 1) - orc.occlusion(n - 1))) < 2 * OCC_TOL, f


def test_maximum_sizes_and_ragged_inputs():
    """Edges of the configuration space: the maximum number of objects (8) with ragged mesh sizes
    including an object without triangles; a 1280x960 frame with the object close enough that the
    window needs several z-tile bands; a single particle."""
    rng = np.random.default_rng(17)
    # -- 8 objects, ragged meshes, one of them empty ------------------------------------------
    parts = [synth.icosphere(1, radius=0.03, seed=k) for k in range(6)]
    parts.append(synth.uv_ellipsoid(6, 7, radii=(0.03, 0.02, 0.02)))
    parts.append((np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)))
    V, T = [p[0] for p in parts], [p[1] for p in parts]
    sensor, orc, rows, cols, n_obj = make_pair((V, T), 4, 6)
    assert n_obj == 8
    gt = [(synth.rodrigues(rng.normal(size=3)), np.array([-0.21 + 0.06 * o, 0.02 * (o % 3 - 1), 0.7 + 0.03 * o]))
          for o in range(8)]
    frame = synth.make_frame(orc.render, gt, rows, cols, rng)
    poses = synth.particle_poses(gt, 6, rng)
    for s in (sensor, orc):
        s.set_observation(frame)
    ia, ib = np.zeros(6, np.int32), np.zeros(6, np.int32)
    assert np.array_equal(sensor.render(poses[0]).view(np.uint32), orc.render(poses[0]).view(np.uint32))
    assert rel_err(sensor.loglikes_poses(poses, ia, True), orc.loglikes_poses(poses, ib, True)) < REL_TOL
    assert np.max(np.abs(sensor.occlusion(5) - orc.occlusion(5))) < OCC_TOL
    # -- 1280 x 960, window far larger than the z tile -----------------------------------------
    rows, cols = 960, 1280
    K = np.array([[1160.0, 0.0, 640.0], [0.0, 1160.0, 480.0], [0.0, 0.0, 1.0]])
    Vb, Tb = synth.mesh_set("icosphere5k")
    sensor = RbSensorBuilder(ObjectModel(Vb, Tb), CameraData(K, rows, cols), Parameters(sample_count=2)).build()
    orc = oracle_lib.Oracle(rows, cols, K, Vb, Tb, synth.default_parameters(), cull=1)
    gt = [(synth.rodrigues([0.2, 0.1, -0.3]), np.array([0.01, -0.02, 0.4]))]
    pose = synth.pack_pose(*gt[0])
    want = orc.render(pose)
    assert np.isfinite(want).sum() > 200000          # > 5 bands of the shared-memory tile
    assert np.array_equal(sensor.render(pose).view(np.uint32), want.view(np.uint32))
    frame = synth.make_frame(orc.render, gt, rows, cols, rng)
    poses = synth.particle_poses(gt, 1, rng)         # a single particle
    for s in (sensor, orc):
        s.set_observation(frame)
    ia, ib = np.zeros(1, np.int32), np.zeros(1, np.int32)
    assert rel_err(sensor.loglikes_poses(poses, ia, True), orc.loglikes_poses(poses, ib, True)) < REL_TOL
    assert np.max(np.abs(sensor.occlusion(0) - orc.occlusion(0))) < OCC_TOL


def test_updates_copy_only_touched_rows_and_images_stay_exact():
    """An update copies only the rectangle of the parent's image that an ancestor has ever written
    (dbot_b200_slot_rect); everything else must already be the background in the recycled destination slot.
    Two groups of particles track objects at different heights and in different columns; resampling moves
    children between the groups, so slots recycled from one patch are reused under parents with the other patch
    and the other way round.  Every occlusion image is compared with the oracle in full, frame after frame."""
    n = 40
    sensor, orc, rows, cols, _ = make_pair("ico80", 4, n)
    rng = np.random.default_rng(21)
    assert sensor.slot_rows(0) == (rows, -1)                      # reset: background everywhere
    top = [(np.eye(3), np.array([-0.15, -0.12, 0.8]))]
    low = [(np.eye(3), np.array([0.17, 0.14, 0.9]))]
    idx = np.zeros(n, dtype=np.int32)
    for f in range(6):
        frame = synth.make_frame(orc.render, top if f % 2 == 0 else low, rows, cols, rng)
        poses = np.concatenate([synth.particle_poses(top, n // 2, rng), synth.particle_poses(low, n - n // 2, rng)])
        if f == 3:
            poses[:] = synth.particle_poses(top, n, rng)         # everybody jumps to the upper object
        if f > 0:
            idx = rng.integers(0, n, size=n).astype(np.int32)     # parents from either group
            if f == 4:
                idx = np.sort(rng.choice(np.arange(n // 2, n), n)).astype(np.int32)   # only lower-group parents
        for s in (sensor, orc):
            s.set_observation(frame)
        ia, ib = idx.copy(), idx.copy()
        got = sensor.loglikes_poses(poses, ia, True)
        want = orc.loglikes_poses(poses, ib, True)
        assert rel_err(got, want) < REL_TOL
        for slot in range(n):
            assert np.max(np.abs(sensor.occlusion(slot) - orc.occlusion(slot))) < OCC_TOL, (f, slot)
        a, b = sensor.slot_rows(0)
        assert 0 <= a <= b < rows
        r0, r1, c0, c1 = sensor.slot_rect(0)
        assert (r0, r1) == (a, b) and 0 <= c0 <= c1 < cols and c0 % 8 == 0
        if f == 0:
            # after the first update a slot's rectangle is its particle's window, rounded to 8 columns
            img = np.isfinite(sensor.render(poses[0]))
            rr, cc = np.nonzero(img)
            # (the window is the box of the projected vertices: it may exceed the covered pixels by one or two)
            assert r0 <= rr.min() <= r0 + 2 and r1 - 2 <= rr.max() <= r1
            assert c0 <= cc.min() < c0 + 10 and c1 - 10 < cc.max() <= c1
    # the patches stayed patches: the upper object never touched the bottom rows and vice versa, and nobody
    # the full width
    spans = [sensor.slot_rect(s) for s in range(n)]
    assert max(r1 - r0 + 1 for r0, r1, _, _ in spans) < rows
    assert max(c1 - c0 + 1 for _, _, c0, c1 in spans) < cols
    sensor.reset()
    assert sensor.slot_rows(0) == (rows, -1)
    assert sensor.slot_rect(0) == (rows, -1, cols, -1)


def test_tall_and_wide_triangles_close_to_the_camera():
    """an 80-triangle icosphere 20-30 cm from a VGA camera: its triangles span 40-300 pixels, i.e. more rows than
    a span batch lane may own (walked by a warp alone) or more than the int32 edge functions allow (int64 block
    walk); depth images bit-exact, log-weights within tolerance, with and without culling"""
    for cull in (True, False):
        sensor, orc, rows, cols, _ = make_pair("ico80", 1, 8, cull=cull)
        rng = np.random.default_rng(17)
        gt = [(synth.rodrigues([0.2, 0.4, -0.1]), np.array([0.01, -0.02, 0.25]))]
        poses = synth.particle_poses(gt, 8, rng, sigma_t=0.02, sigma_r=0.3)
        poses[1, 0, 11] = 0.18
        poses[2, 0, 11] = 0.45
        for p in poses[:4]:
            assert np.array_equal(sensor.render(p).view(np.uint32), orc.render(p).view(np.uint32))
        frame = synth.make_frame(orc.render, gt, rows, cols, rng)
        for s in (sensor, orc):
            s.set_observation(frame)
        ia, ib = np.zeros(8, np.int32), np.zeros(8, np.int32)
        assert rel_err(sensor.loglikes_poses(poses, ia, True), orc.loglikes_poses(poses, ib, True)) < REL_TOL
        assert np.max(np.abs(sensor.occlusion(3) - orc.occlusion(3))) < OCC_TOL


def test_failed_update_on_the_device_needs_reset():
    """a parent index that only the device can see to be wrong (device-pointer entry point) fails the update
    after the slot tables were touched: the error is reported, the handle refuses further work until reset()"""
    import torch

    n = 12
    sensor, orc, rows, cols, _ = make_pair("ico80", 8, n)
    rng = np.random.default_rng(2)
    gt = [(np.eye(3), np.array([0.0, 0.0, 0.8]))]
    frame = synth.make_frame(orc.render, gt, rows, cols, rng)
    poses = synth.particle_poses(gt, n, rng)
    sensor.set_observation(frame)
    sensor.loglikes_poses(poses, np.zeros(n, np.int32), True)
    d_poses = torch.from_numpy(poses).cuda()
    d_idx = torch.arange(n, dtype=torch.int32, device="cuda")
    d_idx[5] = 4000                                           # not a slot
    d_out = torch.zeros(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    sensor.loglikes_device(d_poses.data_ptr(), d_idx.data_ptr(), n, True, d_out.data_ptr())
    with pytest.raises(DbotB200Error) as e:
        sensor.synchronize()
    assert e.value.code == -7
    with pytest.raises(DbotB200Error) as e:
        sensor.loglikes_poses(poses, np.zeros(n, np.int32), False)
    assert e.value.code == -13
    with pytest.raises(DbotB200Error) as e:
        sensor.set_observation(frame)
    assert e.value.code == -13
    sensor.reset()
    orc.reset()
    for s in (sensor, orc):
        s.set_observation(frame)
    ia, ib = np.zeros(n, np.int32), np.zeros(n, np.int32)
    assert rel_err(sensor.loglikes_poses(poses, ia, True), orc.loglikes_poses(poses, ib, True)) < REL_TOL
    # a wrong parent in a call that does not update leaves the handle usable
    d_idx[:] = torch.arange(n, dtype=torch.int32, device="cuda")
    d_idx[2] = -1
    sensor.loglikes_device(d_poses.data_ptr(), d_idx.data_ptr(), n, False, d_out.data_ptr())
    with pytest.raises(DbotB200Error):
        sensor.synchronize()
    assert sensor.loglikes_poses(poses, np.arange(n, dtype=np.int32), False).shape == (n,)


def test_pageable_frames_staged_with_and_without_helper_threads():
    """dbot_b200_set_observation stages a pageable VGA frame in pieces that the calling thread and up to two
    helper threads claim (csrc/host_workers.cuh); page-locked frames are copied as they are.  All three ways must
    put the same frame on the device: identical log-weights, frame after frame, and bad option values are refused."""
    import torch

    n = 16
    sensor, orc, rows, cols, _ = make_pair("ico80", 1, n)
    rng = np.random.default_rng(33)
    gt = [(np.eye(3), np.array([0.02, -0.01, 0.7]))]
    poses = synth.particle_poses(gt, n, rng)
    frames = [synth.make_frame(orc.render, gt, rows, cols, rng) for _ in range(6)]
    results = {}
    for mode in ("helpers", "alone", "pinned"):
        sensor.reset()
        sensor.set_staging_threads(0 if mode == "alone" else 2)
        got = []
        for f in frames:
            img = torch.from_numpy(f.copy()).pin_memory().numpy() if mode == "pinned" else f.copy()
            sensor.set_observation(img)
            img[:] = np.nan            # the caller may reuse its buffer as soon as the call has returned
            got.append(sensor.loglikes_poses(poses, np.zeros(n, np.int32) if not got else np.arange(n, dtype=np.int32), True))
        results[mode] = np.stack(got)
    assert np.array_equal(results["helpers"], results["alone"])
    assert np.array_equal(results["helpers"], results["pinned"])
    orc.reset()
    want = []
    for f in frames:
        orc.set_observation(f)
        want.append(orc.loglikes_poses(poses, np.zeros(n, np.int32) if not want else np.arange(n, dtype=np.int32), True))
    assert rel_err(results["helpers"], np.stack(want)) < REL_TOL
    with pytest.raises(DbotB200Error):
        sensor.set_staging_threads(7)
