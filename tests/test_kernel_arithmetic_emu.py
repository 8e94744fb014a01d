"""The kernel's per-element arithmetic (dbot_ros_b200/csrc/particle_core.cuh, compiled for the host by
tests/emu/) against the oracle, without a GPU: fp32 mixture algebra, integer raster set-up and
coverage, window / band logic.  What this cannot cover -- warp orchestration, shared-memory atomics,
the streaming warps -- is what the -m gpu parity tests are for."""
import numpy as np
import pytest

import oracle_lib
from dbot_ros_b200 import synth

PRM = synth.default_parameters()


def make(mesh, ds, zcap=1 << 20, cull=False):
    rows, cols = synth.resolution(ds)
    K = synth.camera_matrix(ds)
    V, T = synth.mesh_set(mesh)
    orc = oracle_lib.Oracle(rows, cols, K, V, T, PRM, cull=1 if cull else 0)
    signs = [orc.cull_sign(o) for o in range(len(V))] if cull else None
    return oracle_lib.Emu(rows, cols, K, V, T, PRM, zcap=zcap, cull_sign=signs), orc, rows, cols, len(V)


def test_pixel_term_float32_against_double_oracle():
    emu, orc, *_ = make("ico80", 8)
    lib, h = orc.lib, orc.h
    rng = np.random.default_rng(4)
    worst_t, worst_o = 0.0, 0.0
    for _ in range(4000):
        y = rng.uniform(0.4, 3.0)
        sigma = PRM["model_sigma"] + PRM["sigma_factor"] * y * y
        d = y + rng.choice([rng.normal() * sigma, rng.normal() * 6 * sigma, rng.uniform(-0.4, 1.0)])
        if d <= 0.05:
            continue
        o_in, dt = rng.uniform(0.0, 1.0), rng.choice([0.0, 1 / 30.0, 0.5])
        got, post = emu.pixel_term(np.float32(d), np.float32(y), np.float32(o_in), dt)
        yy, dd = float(np.float32(y)), float(np.float32(d))
        o = lib.oracle_propagate(h, float(np.float32(o_in)), dt)
        pv = lib.oracle_p_visible(h, yy, dd) * (1 - o)
        po = lib.oracle_p_occluded(h, yy, dd) * o
        want = np.log((pv + po) / lib.oracle_p_occluded_inf(h, yy))
        worst_t = max(worst_t, abs(got - want) / max(1.0, abs(want)))
        worst_o = max(worst_o, abs(post - po / (pv + po)))
    # a single pixel term may be off by ~1e-5 relative; the log-weight sums thousands of them with
    # errors of both signs and is held to 1e-4 by the parity tests
    assert worst_t < 5e-5 and worst_o < 2e-5


@pytest.mark.parametrize("cull", [False, True])
@pytest.mark.parametrize("mesh,ds,zcap", [("uvsphere480", 4, 1 << 20), ("three20k", 4, 1 << 20),
                                          ("ico80", 8, 160), ("icosphere5k", 2, 3000)])
def test_particle_walk_matches_oracle(mesh, ds, zcap, cull):
    """zcap smaller than the particle's window forces the multi-band path; the emulation walks the
    kernel's (triangle, row) span units, with and without back-face culling"""
    emu, orc, rows, cols, n_obj = make(mesh, ds, zcap, cull)
    rng = synth.rng_for(1)
    gt = synth.ground_truth_poses(n_obj, rng)
    frame = synth.make_frame(orc.render, gt, rows, cols, rng)
    poses = synth.particle_poses(gt, 4, rng, sigma_t=0.01, sigma_r=0.2)
    emu.set_observation(frame)
    orc.set_observation(frame)
    occ0 = np.full(rows * cols, 0.1, np.float32)
    idx = np.zeros(4, dtype=np.int32)
    want = orc.loglikes_poses(poses, idx, True)
    for i in range(4):
        L, occ, depth = emu.particle(poses[i], occ0, 1.0 / 30.0, update=True, want_depth=True)
        assert np.array_equal(depth.view(np.uint32), orc.render(poses[i]).reshape(-1).view(np.uint32))
        assert abs(L - want[i]) / max(1.0, abs(want[i])) < 1e-4
        assert np.max(np.abs(occ.reshape(rows, cols) - orc.occlusion(i))) < 2e-5


def test_wide_triangles_take_the_int64_block_walk_and_agree():
    """an 80-triangle icosphere 25 cm from a VGA camera: triangles wider than DBOT_I32_EXTENT go through
    dbot_covers / dbot_fragment_w instead of the span units; same depth image as the oracle, and the
    same again when every triangle is forced through the block walk"""
    rows, cols = synth.resolution(1)
    K = synth.camera_matrix(1)
    V, T = synth.mesh_set("ico80")
    orc = oracle_lib.Oracle(rows, cols, K, V, T, PRM)
    pose = synth.pack_pose(synth.rodrigues([0.2, 0.4, -0.1]), [0.01, -0.02, 0.25])
    want = orc.render(pose).reshape(-1).view(np.uint32)
    occ = np.full(rows * cols, 0.1, np.float32)
    for force in (0, 1):
        emu = oracle_lib.Emu(rows, cols, K, V, T, PRM, force_blocks=force)
        _, _, depth = emu.particle(pose, occ, 0.0, want_depth=True)
        assert np.array_equal(depth.view(np.uint32), want)


def test_span_units_cover_exactly_the_pixels_of_the_edge_functions():
    """dbot_row_span (float estimate + one exact int32 correction) against dbot_covers (int64 edge
    functions) on random snapped triangles: typical, maximal int32 extent, slivers, vertices and
    edges exactly on pixel sample points, horizontal edges, boxes cut by the raster window"""
    import ctypes as C

    from dbot_ros_b200 import abi

    lib = oracle_lib.emu_lib()
    rng = np.random.default_rng(7)
    n_cov = C.c_int32()
    taken = covered = 0
    for it in range(40000):
        mode = it % 6
        base = rng.integers(-2000, 170000, size=2)
        ext = (3000, 29999, 600, 256 * 8, 40000, 12000)[mode]
        pts = rng.integers(0, ext, size=(3, 2))
        if mode == 3:
            pts = (pts // 256) * 256
            if it % 12 == 3:
                pts[1, 1] = pts[0, 1]
        if mode == 2 and it % 4 == 0:
            pts[2, 1] = pts[1, 1]
        XY = (pts + base).astype(np.int32).reshape(-1)
        if it % 3 == 0:
            row0, row1, col0, col1 = 0, 479, 0, 639
        else:
            c = (XY[0::2].min() >> 8, XY[0::2].max() >> 8)
            r = (XY[1::2].min() >> 8, XY[1::2].max() >> 8)
            col0 = int(rng.integers(c[0] - 2, c[1] + 2)); col1 = int(rng.integers(col0, c[1] + 3))
            row0 = int(rng.integers(r[0] - 2, r[1] + 2)); row1 = int(rng.integers(row0, r[1] + 3))
        m = lib.emu_span_mismatches(abi.ptr(XY), row0, row1, col0, col1, C.byref(n_cov))
        if m >= 0:
            taken += 1
            covered += n_cov.value
            assert m == 0, (XY.tolist(), row0, row1, col0, col1)
    assert taken > 20000 and covered > 1000000
