"""TEST INFRASTRUCTURE: independent numpy re-derivation of SURVEY.md Appendix A (A.3-A.6), written
from the formulas rather than from oracle/dbot_oracle.c, used to cross-check the C oracle:
coverage must agree exactly (integer edge functions), depth to float32 rounding (numpy has no
fmaf), log-weights and occlusion posteriors to double rounding.  Pure numpy, small cases only."""
import numpy as np
from scipy.special import erf

SUBPIX = 256


def p_visible(y, d, w, D, ms, sf):
    s = ms + sf * y * y
    return w / D + (1 - w) * np.exp(-(d - y) ** 2 / (2 * s * s)) / (np.sqrt(2 * np.pi) * s)


def p_occluded(y, d, w, D, ms, sf, lam):
    s = ms + sf * y * y
    s2 = s * s
    return w / D + (1 - w) * lam * np.exp(0.5 * lam * (2 * d - 2 * y + lam * s2)) * (
        1 + erf((d - y + lam * s2) / (np.sqrt(2) * s))) / (2 * (np.exp(lam * d) - 1))


def p_occluded_inf(y, w, D, ms, sf, lam):
    s = ms + sf * y * y
    return w / D + (1 - w) * lam * np.exp(0.5 * lam * (-2 * y + lam * s * s))


def propagate(o, t, p_ov, p_oo):
    c = p_oo - p_ov
    ct = c ** t
    return 1 - (ct * (1 - o) + (1 - p_oo) * (ct - 1) / (c - 1))


def render(K, rows, cols, vertices, triangles, poses):
    """min-z depth image, +inf where empty; poses [n_obj][12]"""
    K = np.asarray(K, dtype=np.float64)
    Kf = K.astype(np.float32)
    Kinv = np.linalg.inv(K).astype(np.float32)
    w = np.zeros((rows, cols), dtype=np.float64)
    for V, T, pose in zip(vertices, triangles, poses):
        R = pose[:9].reshape(3, 3).astype(np.float32)
        t = pose[9:].astype(np.float32)
        # same association as the oracle: r0*x + (r1*y + (r2*z + t)), products in float32
        cam = np.empty((len(V), 3), dtype=np.float32)
        V64 = V.astype(np.float64)
        for k in range(3):
            acc = (R[k, 2].astype(np.float64) * V64[:, 2] + t[k]).astype(np.float32)
            acc = (R[k, 1].astype(np.float64) * V64[:, 1] + acc).astype(np.float32)
            acc = (R[k, 0].astype(np.float64) * V64[:, 0] + acc).astype(np.float32)
            cam[:, k] = acc  # fma emulated with one double rounding (exact product + add fits 53 bits)
        ok = cam[:, 2] > 0
        iz = (np.float32(1.0) / np.where(ok, cam[:, 2], 1)).astype(np.float32)
        xn = (cam[:, 0] * iz).astype(np.float32)
        yn = (cam[:, 1] * iz).astype(np.float32)
        ux = (Kf[0, 0].astype(np.float64) * xn +
              (Kf[0, 1].astype(np.float64) * yn + Kf[0, 2]).astype(np.float32)).astype(np.float32)
        uy = (Kf[1, 1].astype(np.float64) * yn + Kf[1, 2]).astype(np.float32)
        X = np.rint(np.clip(ux * np.float32(SUBPIX), -4194304.0, 4194304.0)).astype(np.int64)
        Y = np.rint(np.clip(uy * np.float32(SUBPIX), -4194304.0, 4194304.0)).astype(np.int64)
        for ia, ib, ic in T:
            if not (ok[ia] and ok[ib] and ok[ic]):
                continue
            xa, ya, xb, yb, xc, yc = X[ia], Y[ia], X[ib], Y[ib], X[ic], Y[ic]
            area2 = (xb - xa) * (yc - ya) - (yb - ya) * (xc - xa)
            if area2 == 0:
                continue
            sg = 1 if area2 > 0 else -1
            cmin = max(-(-min(xa, xb, xc) // SUBPIX), 0)
            cmax = min(max(xa, xb, xc) // SUBPIX, cols - 1)
            rmin = max(-(-min(ya, yb, yc) // SUBPIX), 0)
            rmax = min(max(ya, yb, yc) // SUBPIX, rows - 1)
            if cmin > cmax or rmin > rmax:
                continue
            a, b, c = cam[ia].astype(np.float64), cam[ib].astype(np.float64), cam[ic].astype(np.float64)
            nrm = np.cross(b - a, c - a)
            off = nrm @ a
            if off == 0 or not np.isfinite(off):
                continue
            nrm = nrm / off
            m = nrm @ Kinv.astype(np.float64)
            cc, rr = np.meshgrid(np.arange(cmin, cmax + 1), np.arange(rmin, rmax + 1))
            px, py = cc * SUBPIX, rr * SUBPIX
            e0 = sg * ((xb - xa) * (py - ya) - (yb - ya) * (px - xa))
            e1 = sg * ((xc - xb) * (py - yb) - (yc - yb) * (px - xb))
            e2 = sg * ((xa - xc) * (py - yc) - (ya - yc) * (px - xc))
            inside = (e0 >= 0) & (e1 >= 0) & (e2 >= 0)
            wv = m[0] * cc + m[1] * rr + m[2]
            # inverse depth clamped to the triangle's own vertex range; the largest wins
            wv = np.clip(wv, 1.0 / max(a[2], b[2], c[2]), 1.0 / min(a[2], b[2], c[2]))
            sub = w[rmin:rmax + 1, cmin:cmax + 1]
            np.maximum(sub, np.where(inside, wv, 0.0), out=sub)
    with np.errstate(divide="ignore"):
        return np.where(w > 0, 1.0 / w, np.inf).astype(np.float32)


def loglike(depth, obs, occ, dt, prm):
    """one particle: (log-weight, posterior occlusion image); occ is the stored image, dt its age"""
    w, D, ms, sf = prm["tail_weight"], prm["max_depth"], prm["model_sigma"], prm["sigma_factor"]
    lam = np.log(2.0) / prm["half_life_depth"]
    d = depth.astype(np.float64)
    y = obs.astype(np.float64)
    o = propagate(occ.astype(np.float64), dt, prm["p_occluded_visible"], prm["p_occluded_occluded"])
    use = np.isfinite(d) & ~np.isnan(y)
    dd = np.where(use, d, 1.0)
    yy = np.where(use, y, 1.0)
    pv = p_visible(yy, dd, w, D, ms, sf) * (1 - o)
    po = p_occluded(yy, dd, w, D, ms, sf, lam) * o
    pinf = p_occluded_inf(yy, w, D, ms, sf, lam)
    L = np.sum(np.where(use, np.log((pv + po) / pinf), 0.0))
    post = np.where(use, po / (pv + po), o)
    return L, post
