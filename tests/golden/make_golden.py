#!/usr/bin/env python
"""Generates tests/golden/*.npz.  Run from the repo root:  python tests/golden/make_golden.py

PARITY UNPINNED: the reference tree holds no tests, fixtures or golden vectors for this path and its
arithmetic lives in the un-vendored dbot library, so these files cannot come
from the reference.  They pin what CAN be pinned:

  pixel_model.npz   the per-pixel densities of SURVEY.md A.3 evaluated from their DEFINITION by
                    numerical integration (scipy.integrate.quad of truncated-exponential (x) Gaussian),
                    independent of the closed form the oracle and the kernel use; and the occlusion
                    propagation of A.4 from the 2-state Markov chain's matrix exponential.
  sensor_small.npz  one seeded 80x60 scenario (inputs + the oracle's outputs at generation time):
                    a regression pin for the oracle and the fixture the CUDA path is compared with
                    on the GPU box, where neither /root/reference nor scipy are needed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from dbot_ros_b200 import synth  # noqa: E402


def pixel_model_table():
    from scipy import integrate, linalg

    prm = synth.default_parameters()
    w, D = prm["tail_weight"], prm["max_depth"]
    ms, sf = prm["model_sigma"], prm["sigma_factor"]
    lam = np.log(2.0) / prm["half_life_depth"]
    ys = np.array([0.45, 0.62, 0.80, 1.10, 1.50, 2.40])
    ds = np.array([0.50, 0.79, 0.80, 0.81, 1.00, 1.49, 1.52, 3.00])
    pv = np.empty((len(ys), len(ds)))
    po = np.empty_like(pv)
    pinf = np.empty(len(ys))
    for i, y in enumerate(ys):
        sigma = ms + sf * y * y

        def gauss(c):
            return np.exp(-0.5 * ((y - c) / sigma) ** 2) / (np.sqrt(2 * np.pi) * sigma)

        # occluder anywhere in front of the camera: exponential prior on [0, inf)
        val, _ = integrate.quad(lambda c: lam * np.exp(-lam * c) * gauss(c), max(y - 12 * sigma, 0.0),
                                y + 12 * sigma, epsabs=0, epsrel=1e-13, limit=400)
        pinf[i] = w / D + (1 - w) * val
        for j, d in enumerate(ds):
            pv[i, j] = w / D + (1 - w) * gauss(d)
            # occluder in front of the object: exponential prior truncated to [0, d]
            lo, hi = max(y - 12 * sigma, 0.0), min(y + 12 * sigma, d)
            val = 0.0
            if hi > lo:
                val, _ = integrate.quad(lambda c: lam * np.exp(-lam * c) / (1 - np.exp(-lam * d)) * gauss(c),
                                        lo, hi, epsabs=0, epsrel=1e-13, limit=400)
            po[i, j] = w / D + (1 - w) * val
    # occlusion chain: dO/dt from the transition probabilities over one second
    p_ov, p_oo = prm["p_occluded_visible"], prm["p_occluded_occluded"]
    T1 = np.array([[1 - p_ov, p_ov], [1 - p_oo, p_oo]])  # rows: from visible / from occluded
    G = linalg.logm(T1)
    o0 = np.array([0.0, 0.1, 0.35, 0.9, 1.0])
    dts = np.array([0.0, 1.0 / 30.0, 0.1, 1.0, 2.5])
    prop = np.empty((len(o0), len(dts)))
    for j, t in enumerate(dts):
        Tt = linalg.expm(G * t).real
        for i, o in enumerate(o0):
            prop[i, j] = (1 - o) * Tt[0, 1] + o * Tt[1, 1]
    np.savez(os.path.join(HERE, "pixel_model.npz"), ys=ys, ds=ds, p_visible=pv, p_occluded=po,
             p_occluded_inf=pinf, o0=o0, dts=dts, propagated=prop)


def small_scenario():
    """inputs of sensor_small.npz, rebuilt from the seed (also used by the tests)"""
    ds, n = 8, 12
    rows, cols = synth.resolution(ds)
    K = synth.camera_matrix(ds)
    V, T = synth.mesh_set("ico80")
    return ds, n, rows, cols, K, V, T


def sensor_small():
    ds, n, rows, cols, K, V, T = small_scenario()
    orc = oracle_lib.Oracle(rows, cols, K, V, T, synth.default_parameters())
    rng = np.random.default_rng(20260117)
    gt = synth.ground_truth_poses(1, rng)
    frames, poses, parents, ll_eval, ll_upd, occ = [], [], [], [], [], []
    idx = np.zeros(n, dtype=np.int32)
    for f in range(3):
        frame = synth.make_frame(orc.render, gt, rows, cols, rng)
        p = synth.particle_poses(gt, n, rng, sigma_t=0.01, sigma_r=0.1)
        if f > 0:
            idx = rng.integers(0, n, size=n).astype(np.int32)
        frames.append(frame)
        poses.append(p)
        parents.append(idx.copy())
        orc.set_observation(frame)
        ll_eval.append(orc.loglikes_poses(p, idx, False))
        ll_upd.append(orc.loglikes_poses(p, idx, True))
        occ.append(np.stack([orc.occlusion(s) for s in (0, n - 1)]))
    render0 = orc.render(poses[0][0])
    np.savez_compressed(os.path.join(HERE, "sensor_small.npz"), frames=np.stack(frames), poses=np.stack(poses),
                        parents=np.stack(parents), loglikes_eval=np.stack(ll_eval),
                        loglikes_update=np.stack(ll_upd), occlusion_slots=np.stack(occ), render0=render0,
                        vertices=V[0], triangles=T[0], K=K)


if __name__ == "__main__":
    pixel_model_table()
    sensor_small()
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
