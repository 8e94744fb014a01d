"""TEST INFRASTRUCTURE: ctypes loaders for the CPU oracle (oracle/libdbot_oracle.so) and for the
host emulation of the kernel phases (tests/emu/libdbot_emu.so).  Never imported by the package."""
import ctypes as C
import os
import subprocess

import numpy as np

from dbot_ros_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libdbot_oracle.so")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libdbot_emu.so")
CORE_H = os.path.join(ROOT, "dbot_ros_b200", "csrc", "particle_core.cuh")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_oracle():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("dbot_oracle.c", "dbot_oracle.h", "Makefile")]
    if _stale(ORACLE_SO, srcs):
        subprocess.run(["make", "-C", ORACLE_DIR, "-B"], check=True, capture_output=True)
    return ORACLE_SO


def build_emu():
    src = os.path.join(EMU_DIR, "emu_core.cpp")
    if _stale(EMU_SO, [src, CORE_H]):
        subprocess.run(
            ["g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17",
             src, "-o", EMU_SO], check=True, capture_output=True)
    return EMU_SO


class OracleConfig(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("cols", C.c_int32), ("K", C.c_double * 9), ("n_objects", C.c_int32),
        ("vertices", C.POINTER(C.POINTER(C.c_float))), ("n_vertices", C.POINTER(C.c_int32)),
        ("triangles", C.POINTER(C.POINTER(C.c_int32))), ("n_triangles", C.POINTER(C.c_int32)),
        ("tail_weight", C.c_double), ("model_sigma", C.c_double), ("sigma_factor", C.c_double),
        ("half_life_depth", C.c_double), ("max_depth", C.c_double),
        ("p_occluded_visible", C.c_double), ("p_occluded_occluded", C.c_double),
        ("initial_occlusion_prob", C.c_double), ("delta_time", C.c_double),
        ("eager", C.c_int32), ("float_intermediates", C.c_int32), ("cull_backfaces", C.c_int32),
    ]


class EmuConfig(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("cols", C.c_int32), ("K", C.c_double * 9), ("n_objects", C.c_int32),
        ("vertices", C.POINTER(C.POINTER(C.c_float))), ("n_vertices", C.POINTER(C.c_int32)),
        ("triangles", C.POINTER(C.POINTER(C.c_int32))), ("n_triangles", C.POINTER(C.c_int32)),
        ("tail_weight", C.c_double), ("model_sigma", C.c_double), ("sigma_factor", C.c_double),
        ("half_life_depth", C.c_double), ("max_depth", C.c_double),
        ("p_occluded_visible", C.c_double), ("p_occluded_occluded", C.c_double),
        ("zcap", C.c_int32),
        ("cull_sign", C.c_int32 * 8),   # per object: 0 = keep both orientations
        ("force_blocks", C.c_int32),    # 1: every triangle through the int64 block walk (dbot_covers)
    ]


_oracle = None
_emu = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(build_oracle())
        P = C.c_void_p
        lib.oracle_create.argtypes = [C.POINTER(OracleConfig), C.POINTER(P)]
        lib.oracle_destroy.argtypes = [P]
        lib.oracle_reset.argtypes = [P]
        lib.oracle_set_observation.argtypes = [P, P]
        lib.oracle_loglikes.argtypes = [P, P, P, C.c_int32, C.c_int32, P]
        lib.oracle_render.argtypes = [P, P, P]
        lib.oracle_n_slots.argtypes = [P]
        lib.oracle_observation_time.argtypes = [P]
        lib.oracle_observation_time.restype = C.c_double
        lib.oracle_get_occlusion.argtypes = [P, C.c_int32, P]
        lib.oracle_get_occlusion_raw.argtypes = [P, C.c_int32, P, P]
        lib.oracle_import_slot.argtypes = [P, P, P]
        lib.oracle_set_num_threads.argtypes = [C.c_int]
        lib.oracle_get_max_threads.restype = C.c_int
        lib.oracle_cull_sign.argtypes = [P, C.c_int32]
        for f in ("oracle_p_visible", "oracle_p_occluded", "oracle_propagate"):
            getattr(lib, f).argtypes = [P, C.c_double, C.c_double]
            getattr(lib, f).restype = C.c_double
        lib.oracle_p_occluded_inf.argtypes = [P, C.c_double]
        lib.oracle_p_occluded_inf.restype = C.c_double
        _oracle = lib
    return _oracle


def emu_lib():
    global _emu
    if _emu is None:
        lib = C.CDLL(build_emu())
        P = C.c_void_p
        lib.emu_observation_table.argtypes = [C.POINTER(EmuConfig), P, P]
        lib.emu_particle.argtypes = [C.POINTER(EmuConfig), P, P, P, C.c_float, C.c_float, P, P]
        lib.emu_particle.restype = C.c_double
        lib.emu_pixel_term.argtypes = [C.POINTER(EmuConfig), C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.c_float, C.POINTER(C.c_float)]
        lib.emu_pixel_term.restype = C.c_float
        lib.emu_span_mismatches.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
        _emu = lib
    return _emu


class Oracle:
    """oracle_* behind the same method names as dbot_ros_b200.B200KinectImageModel"""

    def __init__(self, rows, cols, K, vertices, triangles, params, eager=1, float_intermediates=0,
                 cull=0):
        self.lib = oracle_lib()
        self.rows, self.cols, self.n_objects = rows, cols, len(vertices)
        cfg, self._keep = abi.make_config(rows, cols, K, 0, vertices, triangles, struct=OracleConfig,
                                          **params)
        cfg.eager = eager
        cfg.float_intermediates = float_intermediates
        cfg.cull_backfaces = cull
        self.h = C.c_void_p()
        rc = self.lib.oracle_create(C.byref(cfg), C.byref(self.h))
        if rc:
            raise RuntimeError("oracle_create failed: %d" % rc)

    def reset(self):
        assert self.lib.oracle_reset(self.h) == 0

    def set_observation(self, image):
        img = np.ascontiguousarray(image, dtype=np.float32).reshape(-1)
        assert img.size == self.rows * self.cols
        assert self.lib.oracle_set_observation(self.h, abi.ptr(img)) == 0

    def loglikes_poses(self, poses, indices, update=False):
        poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, self.n_objects, 12)
        n = poses.shape[0]
        assert indices.dtype == np.int32 and indices.size == n
        out = np.empty(n, dtype=np.float64)
        rc = self.lib.oracle_loglikes(self.h, abi.ptr(poses), abi.ptr(indices), n, 1 if update else 0,
                                      abi.ptr(out))
        if rc:
            raise RuntimeError("oracle_loglikes failed: %d" % rc)
        return out

    def render(self, pose):
        pose = np.ascontiguousarray(pose, dtype=np.float32).reshape(self.n_objects, 12)
        out = np.empty(self.rows * self.cols, dtype=np.float32)
        assert self.lib.oracle_render(self.h, abi.ptr(pose), abi.ptr(out)) == 0
        return out.reshape(self.rows, self.cols)

    def cull_sign(self, obj):
        return self.lib.oracle_cull_sign(self.h, obj)

    @property
    def n_slots(self):
        return self.lib.oracle_n_slots(self.h)

    @property
    def observation_time(self):
        return self.lib.oracle_observation_time(self.h)

    def occlusion(self, slot):
        out = np.empty(self.rows * self.cols, dtype=np.float32)
        assert self.lib.oracle_get_occlusion(self.h, slot, abi.ptr(out)) == 0
        return out.reshape(self.rows, self.cols)

    def occlusion_raw(self, slot):
        o = np.empty(self.rows * self.cols, dtype=np.float32)
        t = np.empty(self.rows * self.cols, dtype=np.float64)
        assert self.lib.oracle_get_occlusion_raw(self.h, slot, abi.ptr(o), abi.ptr(t)) == 0
        return o.reshape(self.rows, self.cols), t.reshape(self.rows, self.cols)

    # -- slot migration, as the sharded evaluator drives it (payload: float32 vector) -----------
    @property
    def slot_floats(self):
        return 3 * self.rows * self.cols           # P floats + P doubles

    def export_slot(self, slot):
        o, t = self.occlusion_raw(slot)
        return np.concatenate([o.reshape(-1), t.reshape(-1).view(np.float32)])

    def import_slot(self, payload):
        P = self.rows * self.cols
        payload = np.ascontiguousarray(payload, dtype=np.float32)
        o = np.ascontiguousarray(payload[:P])
        t = np.ascontiguousarray(payload[P:]).view(np.float64)
        idx = self.lib.oracle_import_slot(self.h, abi.ptr(o), abi.ptr(t))
        assert idx >= 0
        return idx

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Emu:
    """one-particle host walk through the kernel's phase functions (tests/emu/emu_core.cpp)"""

    def __init__(self, rows, cols, K, vertices, triangles, params, zcap=1 << 20, cull_sign=None,
                 force_blocks=0):
        self.lib = emu_lib()
        self.rows, self.cols, self.n_objects = rows, cols, len(vertices)
        p = {k: v for k, v in params.items()}
        cfg, self._keep = abi.make_config(rows, cols, K, 0, vertices, triangles, struct=EmuConfig, **p)
        cfg.zcap = zcap
        for o, sg in enumerate(cull_sign or []):
            cfg.cull_sign[o] = sg
        cfg.force_blocks = force_blocks
        self.cfg = cfg
        self.params = params
        self.tab = None

    def set_observation(self, image):
        img = np.ascontiguousarray(image, dtype=np.float32).reshape(-1)
        self.tab = np.empty((self.rows * self.cols, 4), dtype=np.float32)
        self.lib.emu_observation_table(C.byref(self.cfg), abi.ptr(img), abi.ptr(self.tab))

    def prop_affine(self, dt):
        c = self.params["p_occluded_occluded"] - self.params["p_occluded_visible"]
        ct = np.exp(dt * np.log(c))
        k = (1.0 - self.params["p_occluded_occluded"]) * (ct - 1.0) / (c - 1.0)
        return np.float32(ct), np.float32(1.0 - ct - k)

    def particle(self, pose, occ_in, dt, update=False, want_depth=False):
        pose = np.ascontiguousarray(pose, dtype=np.float32).reshape(self.n_objects, 12)
        occ_in = np.ascontiguousarray(occ_in, dtype=np.float32).reshape(-1)
        a, b = self.prop_affine(dt)
        occ_out = np.empty_like(occ_in) if update else None
        depth = np.empty(self.rows * self.cols, dtype=np.float32) if want_depth else None
        L = self.lib.emu_particle(C.byref(self.cfg), abi.ptr(pose),
                                  abi.ptr(self.tab) if self.tab is not None else None,
                                  abi.ptr(occ_in), a, b,
                                  abi.ptr(occ_out) if update else None,
                                  abi.ptr(depth) if want_depth else None)
        return L, occ_out, depth

    def pixel_term(self, d, y, o_in, dt=0.0):
        a, b = self.prop_affine(dt)
        post = C.c_float()
        t = self.lib.emu_pixel_term(C.byref(self.cfg), d, y, o_in, a, b, C.byref(post))
        return t, post.value
