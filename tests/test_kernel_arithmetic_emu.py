"""The kernel's per-element arithmetic (dbot_ros_b200/csrc/particle_core.cuh, compiled for the host by
tests/emu/) against the oracle, without a GPU: fp32 mixture algebra, integer raster set-up and
coverage, window / band logic.  What this cannot cover -- warp orchestration, shared-memory atomics,
the streaming warps -- is what the -m gpu parity tests are for."""
import numpy as np
import pytest

import oracle_lib
from dbot_ros_b200 import synth

PRM = synth.default_parameters()


def make(mesh, ds, zcap=1 << 20):
    rows, cols = synth.resolution(ds)
    K = synth.camera_matrix(ds)
    V, T = synth.mesh_set(mesh)
    return (oracle_lib.Emu(rows, cols, K, V, T, PRM, zcap=zcap), oracle_lib.Oracle(rows, cols, K, V, T, PRM),
            rows, cols, len(V))


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


@pytest.mark.parametrize("mesh,ds,zcap", [("uvsphere480", 4, 1 << 20), ("three20k", 4, 1 << 20),
                                          ("ico80", 8, 160), ("icosphere5k", 2, 3000)])
def test_particle_walk_matches_oracle(mesh, ds, zcap):
    """zcap smaller than the particle's window forces the multi-band path"""
    emu, orc, rows, cols, n_obj = make(mesh, ds, zcap)
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
