"""Seeded synthetic inputs for the depth observation model.

The reference ships no meshes, depth frames or bags (duck.obj / impact_battery.obj are only named
in ref:README.md:64 and ref:config/object.yaml:5), so tests and bench.py use these stand-ins:
a bumpy icosphere (level 4: 2562 vertices, 5120 triangles), a 16x16 UV sphere (480 triangles),
a torus and a UV ellipsoid for the multi-object set.  Nothing here calls the oracle or the CUDA
library; scene construction takes the renderer as an argument.
"""
import numpy as np

# native VGA intrinsics used throughout
K_VGA = np.array([[580.0, 0.0, 320.0], [0.0, 580.0, 240.0], [0.0, 0.0, 1.0]])


def camera_matrix(downsampling=1):
    """K with rows 0-1 divided by the factor (ref:source/dbot_ros/util/ros_camera_data_provider.cpp:72)."""
    K = K_VGA.copy()
    K[:2, :] /= float(downsampling)
    return K


def resolution(downsampling=1):
    return 480 // downsampling, 640 // downsampling  # rows, cols


# ----------------------------------------------------------------------------------------------
def icosphere(level, radius=0.10, bump=0.08, seed=0):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t),
         (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
             (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8),
             (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(level):
        cache = {}
        new_faces = []

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = new_faces
    V = np.array(verts)
    ids = np.arange(len(V), dtype=np.float64)
    # deterministic "sin hash" bumpiness: non-convex surface that exercises the z test
    h = np.sin(ids * 12.9898 + seed * 78.233) * 43758.5453
    h = h - np.floor(h)
    V = V * (radius * (1.0 + bump * np.sin(2.0 * np.pi * h)))[:, None]
    return V.astype(np.float32), np.array(faces, dtype=np.int32)


def uv_ellipsoid(n_lat, n_lon, radii=(0.08, 0.08, 0.08)):
    """poles + (n_lat-1) rings of n_lon vertices: 2*n_lon*(n_lat-1) triangles."""
    verts = [(0.0, 0.0, 1.0)]
    for i in range(1, n_lat):
        th = np.pi * i / n_lat
        for j in range(n_lon):
            ph = 2.0 * np.pi * j / n_lon
            verts.append((np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)))
    verts.append((0.0, 0.0, -1.0))
    south = len(verts) - 1
    tris = []
    ring = lambda i, j: 1 + (i - 1) * n_lon + (j % n_lon)  # noqa: E731
    for j in range(n_lon):
        tris.append((0, ring(1, j), ring(1, j + 1)))
        tris.append((south, ring(n_lat - 1, j + 1), ring(n_lat - 1, j)))
    for i in range(1, n_lat - 1):
        for j in range(n_lon):
            a, b, c, d = ring(i, j), ring(i, j + 1), ring(i + 1, j), ring(i + 1, j + 1)
            tris += [(a, c, b), (b, c, d)]
    V = np.array(verts) * np.array(radii)[None, :]
    return V.astype(np.float32), np.array(tris, dtype=np.int32)


def torus(n_major, n_minor, R=0.09, r=0.03):
    verts, tris = [], []
    for i in range(n_major):
        a = 2.0 * np.pi * i / n_major
        for j in range(n_minor):
            b = 2.0 * np.pi * j / n_minor
            verts.append(((R + r * np.cos(b)) * np.cos(a), (R + r * np.cos(b)) * np.sin(a), r * np.sin(b)))
    idx = lambda i, j: (i % n_major) * n_minor + (j % n_minor)  # noqa: E731
    for i in range(n_major):
        for j in range(n_minor):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i, j + 1), idx(i + 1, j + 1)
            tris += [(a, b, c), (b, d, c)]
    return np.array(verts, dtype=np.float32), np.array(tris, dtype=np.int32)


def mesh_set(name):
    """-> (list of vertex arrays, list of triangle arrays)"""
    if name == "icosphere5k":     # BASELINE configs 2-4: single ~5k-triangle mesh
        v, t = icosphere(4, seed=1)
        return [v], [t]
    if name == "uvsphere480":     # BASELINE config 1: duck.obj stand-in (~500 triangles)
        v, t = uv_ellipsoid(16, 16)
        return [v], [t]
    if name == "three20k":        # BASELINE config 5: 3 meshes, ~20k triangles
        a = icosphere(4, radius=0.08, seed=2)
        b = torus(80, 48)
        c = uv_ellipsoid(60, 60, radii=(0.09, 0.06, 0.05))
        return [a[0], b[0], c[0]], [a[1], b[1], c[1]]
    if name == "ico80":
        v, t = icosphere(1, seed=3)
        return [v], [t]
    raise KeyError(name)


# ----------------------------------------------------------------------------------------------
def rodrigues(rv):
    rv = np.asarray(rv, dtype=np.float64)
    th = np.linalg.norm(rv)
    if th == 0.0:
        return np.eye(3)
    k = rv / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1.0 - np.cos(th)) * (Kx @ Kx)


def pack_pose(R, t):
    return np.concatenate([np.asarray(R, dtype=np.float64).reshape(9), np.asarray(t, dtype=np.float64)]).astype(np.float32)


def ground_truth_poses(n_objects, rng):
    """objects 0.7-1.0 m in front of the camera, inside the frustum, side by side"""
    poses = []
    xs = np.linspace(-0.18, 0.18, n_objects) if n_objects > 1 else [0.0]
    for o in range(n_objects):
        t = np.array([xs[o] + rng.uniform(-0.02, 0.02), rng.uniform(-0.05, 0.05), rng.uniform(0.7, 1.0)])
        R = rodrigues(rng.normal(size=3))
        poses.append((R, t))
    return poses


def particle_poses(gt, n, rng, sigma_t=0.005, sigma_r=0.05):
    """ground truth (+) N(0,(5 mm)^2) translation, N(0,(0.05 rad)^2) rotation vector; float32
    [n][n_objects][12] exactly as the C-ABI takes them."""
    n_obj = len(gt)
    out = np.empty((n, n_obj, 12), dtype=np.float32)
    for o, (R, t) in enumerate(gt):
        dt = rng.normal(scale=sigma_t, size=(n, 3))
        dr = rng.normal(scale=sigma_r, size=(n, 3))
        for i in range(n):
            out[i, o] = pack_pose(rodrigues(dr[i]) @ R, t + dt[i])
    return out


def make_frame(render, gt, rows, cols, rng, *, model_sigma=0.003, sigma_factor=0.0014247,
               background=1.5, nan_fraction=0.02, occluder=True, K=None):
    """depth frame = render of the ground-truth poses, background plane, one occluder box at
    0.5 m over ~20 % of the object, sigma(z) Gaussian noise, 2 % NaN.  `render(pose[n_obj][12])`
    returns rows x cols float32 with +inf where nothing is hit."""
    gt_pose = np.stack([pack_pose(R, t) for R, t in gt])
    d = np.array(render(gt_pose), dtype=np.float64).reshape(rows, cols)
    covered = np.isfinite(d)
    d = np.where(covered, d, background)
    if occluder and covered.any():
        rr, cc = np.nonzero(covered)
        r0, r1, c0, c1 = rr.min(), rr.max(), cc.min(), cc.max()
        # left ~35 % of the bounding box, middle half of its rows
        oc1 = c0 + int(0.35 * (c1 - c0 + 1))
        or0 = r0 + (r1 - r0 + 1) // 4
        or1 = r1 - (r1 - r0 + 1) // 4
        d[or0:or1 + 1, c0:oc1 + 1] = 0.5
    sigma = model_sigma + sigma_factor * d * d
    d = d + rng.normal(size=d.shape) * sigma
    d = d.astype(np.float32)
    d[rng.random(d.shape) < nan_fraction] = np.nan
    return d


def default_parameters():
    """ref:config/particle_tracker.yaml:41-49 + dbot pixel-model defaults + 1/30 s"""
    return dict(tail_weight=0.01, model_sigma=0.003, sigma_factor=0.0014247, half_life_depth=1.0,
                max_depth=6.0, p_occluded_visible=0.1, p_occluded_occluded=0.7,
                initial_occlusion_prob=0.1, delta_time=1.0 / 30.0)


CONFIGS = {
    # BASELINE.json configs[0..4]: (mesh set, N, downsampling)
    1: ("uvsphere480", 200, 4),
    2: ("icosphere5k", 1000, 1),
    3: ("icosphere5k", 10000, 1),
    4: ("icosphere5k", 10000, 1),   # per GPU; 80000 over 8
    5: ("three20k", 2000, 1),
}


def rng_for(config_id, stream=0):
    return np.random.default_rng(20260117 + config_id + 1000 * stream)
