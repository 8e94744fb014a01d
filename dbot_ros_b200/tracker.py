"""The caller of the hot path: a host-side restatement of dbot's Rao-Blackwellised coordinate particle
filter and of ParticleTracker::on_track, just enough to drive the sensor the way dbot_ros does
(SURVEY.md 3.2, A.7, section 8f row N1).  It exists so that the observation model can be exercised in
its real loop -- set_observation once per frame, one loglikes per sampling block with update on the
last, KL-triggered resampling that permutes `indices` -- and so that a synthetic sequence can show
that the log-weights actually track an object.  All O(N) work here is numpy on the host; the
per-particle heavy lifting is the sensor's.

PARITY UNPINNED: dbot's filter/tracker sources are not in the reference tree (un-vendored dbot).  The
call pattern (the part the sensor depends on) follows SURVEY.md A.7; the motion model is the
"random motion with Gaussian noise injected into the velocities" that
ref:config/particle_tracker.yaml:51-63 parameterises (sigmas, velocity_factor) and
ref:source/dbot_ros/tracker/particle_tracker_node.cpp:138-159 builds:
    v' = velocity_factor * v + sigma * w / dt,   pose' = pose + v' * dt,   w ~ N(0, I).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class ObjectTransitionParameters:
    """ref:config/particle_tracker.yaml:55-63"""

    linear_sigma: tuple = (0.0025, 0.0025, 0.0025)
    angular_sigma: tuple = (0.02, 0.02, 0.02)
    velocity_factor: float = 0.8

    @classmethod
    def from_dict(cls, tree):
        t = tree.get("particle_filter", tree)["object_transition"]
        return cls((float(t["linear_sigma_x"]), float(t["linear_sigma_y"]), float(t["linear_sigma_z"])),
                   (float(t["angular_sigma_x"]), float(t["angular_sigma_y"]), float(t["angular_sigma_z"])),
                   float(t["velocity_factor"]))


@dataclass
class TrackerParameters:
    """ParticleTrackerBuilder parameters (ref:...particle_tracker_node.cpp:208-214)"""

    evaluation_count: int = 2000
    moving_average_update_rate: float = 1.0
    max_kl_divergence: float = 2.0
    delta_time: float = 1.0 / 30.0
    transition: ObjectTransitionParameters = field(default_factory=ObjectTransitionParameters)
    center_object_frame: bool = False   # ref:config/particle_tracker.yaml:30, ref:...particle_tracker_node.cpp:212-214


# ---- Tracker::track post-processing shared by both trackers (dbot tracker/tracker.cpp, un-vendored) ----------
def _quat(v):
    a = float(np.linalg.norm(v))
    k = np.sin(0.5 * a) / a if a > 1e-12 else 0.5
    return np.concatenate([[np.cos(0.5 * a)], k * np.asarray(v, dtype=np.float64)])


def _rotvec_of_quat(q):
    q = -q if q[0] < 0.0 else q
    sn = float(np.linalg.norm(q[1:]))
    a = 2.0 * np.arctan2(sn, q[0])
    return (a / sn if sn > 1e-12 else 2.0) * q[1:]


def moving_average(smoothed, current, rate):
    """states [n_obj][12]: positions and velocities blend linearly, orientations as quaternions (sign-aligned,
    renormalised) -- dbot's Tracker keeps a moving average of the state it returns
    (moving_average_update_rate, ref:config/particle_tracker.yaml:25)"""
    if smoothed is None or rate >= 1.0:
        return current.copy()
    out = (1.0 - rate) * smoothed + rate * current
    for o in range(current.shape[0]):
        qa, qb = _quat(smoothed[o, 3:6]), _quat(current[o, 3:6])
        if np.dot(qa, qb) < 0.0:
            qb = -qb
        q = (1.0 - rate) * qa + rate * qb
        out[o, 3:6] = _rotvec_of_quat(q / np.linalg.norm(q))
    return out


def to_center_frame(poses, centers):
    """model frame -> frame of the centred mesh (dbot Tracker::to_center_coordinate_system): t += R c"""
    from .rb_sensor import rotation_matrices

    out = np.array(poses, dtype=np.float64, copy=True)
    for o, c in enumerate(centers):
        out[o, 0:3] += rotation_matrices(out[o, 3:6]) @ np.asarray(c, dtype=np.float64)
    return out


def to_model_frame(states, centers):
    """inverse of to_center_frame (dbot Tracker::to_model_coordinate_system): t -= R c"""
    from .rb_sensor import rotation_matrices

    out = np.array(states, dtype=np.float64, copy=True)
    for o, c in enumerate(centers):
        out[o, 0:3] -= rotation_matrices(out[o, 3:6]) @ np.asarray(c, dtype=np.float64)
    return out


def _centers_for(params, n_obj, centers):
    if centers is None:
        if params.center_object_frame:
            raise ValueError("center_object_frame is set: pass the ObjectModel's centers (object_model.centers) "
                             "so that poses can be converted between the model frame and the centred mesh")
        return np.zeros((n_obj, 3))
    c = np.asarray(centers, dtype=np.float64).reshape(n_obj, 3)
    return c if params.center_object_frame else np.zeros((n_obj, 3))


# ---- counter-based noise shared with the device-resident filter (csrc/dbot_filter.cuh) ----------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    with np.errstate(over="ignore"):
        x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def noise_key(seed, frame, block, particle, pair):
    with np.errstate(over="ignore"):
        base = splitmix64(np.uint64(seed) ^ ((np.uint64(frame) * np.uint64(0x100000001B3)) & _M64))
        return (base ^ (np.uint64(block) << np.uint64(56)) ^ (np.asarray(particle, dtype=np.uint64) << np.uint64(8))
                ^ np.uint64(pair))


def counter_normals(seed, frame, block, n):
    """[n][6] standard normals exactly as dbot_filter_propagate_kernel draws them (Box-Muller on two
    splitmix64 outputs per pair; equal up to the last bits of log/sincos)"""
    out = np.empty((n, 6))
    particles = np.arange(n, dtype=np.uint64)
    for pair in range(3):
        h1 = splitmix64(noise_key(seed, frame, block, particles, pair))
        h2 = splitmix64(h1)
        u1 = ((h1 >> np.uint64(11)).astype(np.float64) + 1.0) / 9007199254740992.0
        u2 = (h2 >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
        r = np.sqrt(-2.0 * np.log(u1))
        out[:, 2 * pair] = r * np.cos(2.0 * np.pi * u2)
        out[:, 2 * pair + 1] = r * np.sin(2.0 * np.pi * u2)
    return out


def counter_uniform(seed, frame, block):
    h = splitmix64(noise_key(seed, frame, block, 0x7FFFFF, 7))
    return float((h >> np.uint64(11)).astype(np.float64) / 9007199254740992.0)


def systematic_resample(weights, n, u0):
    """n parents from normalised weights with one uniform draw (sorted output)"""
    cdf = np.cumsum(weights)
    cdf[-1] = 1.0
    u = (u0 + np.arange(n)) / n
    return np.minimum(np.searchsorted(cdf, u, side="right"), len(weights) - 1).astype(np.int32)


class ParticleTracker:
    """tracker->initialize(poses); tracker->track(image) (ref:...particle_tracker_node.cpp:252,
    ref:source/dbot_ros/object_tracker_ros.hpp:49).  States are deltas relative to the sensor's
    integrated poses: [n][n_obj][12] = position, rotation vector, linear velocity, angular velocity."""

    def __init__(self, sensor, n_objects, params: TrackerParameters, seed=0, counter_noise=False, centers=None):
        """counter_noise: draw the noise from the counter-based generator of the device-resident
        filter (same seed -> same noise on both) instead of numpy's generator.
        centers: ObjectModel.centers when the meshes were centred (center_object_frame): initialize() and
        track() speak the MODEL frame, as the reference's tracker does."""
        self.sensor = sensor
        self.n_obj = int(n_objects)
        self.centers = _centers_for(params, self.n_obj, centers)
        self.p = params
        self.seed = int(seed)
        self.counter_noise = counter_noise
        self.frame = 0
        self.rng = np.random.default_rng(seed)
        self.n = max(1, params.evaluation_count // self.n_obj)
        self.states = np.zeros((self.n, self.n_obj, 12))
        self.log_w = np.full(self.n, -np.log(self.n))
        self.indices = np.zeros(self.n, dtype=np.int32)
        self.loglikes = np.zeros(self.n)
        self.resample_count = 0
        self.smoothed = None

    # -- Tracker::initialize ------------------------------------------------------------------
    def initialize(self, initial_poses):
        """initial_poses: [n_obj][6] position + rotation vector of the model frame in the camera frame"""
        poses = np.asarray(initial_poses, dtype=np.float64).reshape(self.n_obj, 6)
        self.sensor.integrated_poses = to_center_frame(poses, self.centers)
        self.smoothed = None
        self.states[:] = 0.0
        self.log_w[:] = -np.log(self.n)
        self.indices[:] = 0
        self.loglikes[:] = 0.0
        self.frame = 0
        self.sensor.reset()

    # -- RBC-PF::filter -----------------------------------------------------------------------
    def kl_given_uniform(self):
        w = np.exp(self.log_w)
        return float(np.log(self.n) + np.sum(w * self.log_w))

    def resample(self):
        w = np.exp(self.log_w)
        u0 = counter_uniform(self.seed, self.frame, self._block) if self.counter_noise else self.rng.random()
        parents = systematic_resample(w / w.sum(), self.n, u0)
        self.states = self.states[parents]
        self.next_states = self.next_states[parents]
        self.noise = self.noise[parents]
        self.indices = self.indices[parents].copy()
        self.loglikes = self.loglikes[parents]
        self.log_w[:] = -np.log(self.n)
        self.resample_count += 1

    def propagate(self, states, noise):
        t = self.p.transition
        dt = self.p.delta_time
        sig = np.concatenate([t.linear_sigma, t.angular_sigma])
        nxt = states.copy()
        nxt[:, :, 6:12] = t.velocity_factor * states[:, :, 6:12] + noise * (sig / dt)
        nxt[:, :, 0:6] = states[:, :, 0:6] + nxt[:, :, 6:12] * dt
        return nxt

    def filter(self, image):
        self.sensor.set_observation(image)
        self.loglikes[:] = 0.0
        self.noise = np.zeros((self.n, self.n_obj, 6))
        self.next_states = self.states
        for block in range(self.n_obj):                 # one sampling block per object
            self._block = block
            self.noise[:, block, :] = (counter_normals(self.seed, self.frame, block, self.n) if self.counter_noise
                                       else self.rng.standard_normal((self.n, 6)))
            self.next_states = self.propagate(self.states, self.noise)
            update = block == self.n_obj - 1
            new = self.sensor.loglikes(self.next_states.reshape(self.n, -1), self.indices, update)
            self.log_w += new - self.loglikes
            self.log_w -= np.logaddexp.reduce(self.log_w)
            self.loglikes = new
            if self.kl_given_uniform() > self.p.max_kl_divergence:
                self.resample()
        self.states = self.next_states
        self.frame += 1

    # -- ParticleTracker::on_track + Tracker::track -------------------------------------------
    def track(self, image):
        from .rb_sensor import rotation_matrices

        """-> [n_obj][12]: position, rotation vector, linear and angular velocity (the State the reference
        publishes, ref:source/dbot_ros/object_tracker_ros.hpp:49-61), moving average applied, model frame"""
        self.filter(image)
        w = np.exp(self.log_w)
        mean = np.tensordot(w, self.states[:, :, 0:6], axes=(0, 0))            # [n_obj][6]
        mean_vel = np.tensordot(w, self.states[:, :, 6:12], axes=(0, 0))
        # fold the mean delta into the integrated pose and re-base the particles on it
        poses = self.sensor.integrated_poses
        for o in range(self.n_obj):
            Rm = rotation_matrices(mean[o, 3:6])
            R0 = rotation_matrices(poses[o, 3:6])
            poses[o, 0:3] += mean[o, 0:3]
            poses[o, 3:6] = rotation_vector(Rm @ R0)
            self.states[:, o, 0:3] -= mean[o, 0:3]
            Rd = rotation_matrices(self.states[:, o, 3:6])
            self.states[:, o, 3:6] = rotation_vectors(Rd @ Rm.T)
        self.sensor.integrated_poses = poses
        self.smoothed = moving_average(self.smoothed, np.hstack([poses, mean_vel]), self.p.moving_average_update_rate)
        return to_model_frame(self.smoothed, self.centers)


def rotation_vector(R):
    return rotation_vectors(np.asarray(R)[None])[0]


def rotation_vectors(R):
    """log map of rotation matrices [..., 3, 3] -> rotation vectors [..., 3] (angles below pi)"""
    R = np.asarray(R, dtype=np.float64)
    c = np.clip((np.trace(R, axis1=-2, axis2=-1) - 1.0) * 0.5, -1.0, 1.0)
    angle = np.arccos(c)
    v = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], axis=-1)
    s = np.sin(angle)
    scale = np.where(s > 1e-9, angle / np.where(s > 1e-9, 2.0 * s, 1.0), 0.5)
    return v * scale[..., None]


class DeviceParticleTracker:
    """The same tracker with the whole filter step on the GPU (libdbot_b200.so, dbot_b200_filter_*):
    per frame the depth image goes in and the integrated poses come out; particles, poses, parent
    indices and weights never leave HBM."""

    def __init__(self, sensor, params: TrackerParameters, seed=0, centers=None):
        import ctypes as C

        from . import abi

        self._C, self._abi = C, abi
        self.sensor = sensor
        self.n_obj = sensor.n_objects
        self.centers = _centers_for(params, self.n_obj, centers)
        self.n = max(1, params.evaluation_count // self.n_obj)
        cfg = abi.FilterConfig()
        for k in range(3):
            cfg.linear_sigma[k] = params.transition.linear_sigma[k]
            cfg.angular_sigma[k] = params.transition.angular_sigma[k]
        cfg.velocity_factor = params.transition.velocity_factor
        cfg.max_kl_divergence = params.max_kl_divergence
        cfg.seed = int(seed)
        cfg.n_particles = self.n
        cfg.moving_average_update_rate = params.moving_average_update_rate
        for o in range(self.n_obj):
            for k in range(3):
                cfg.centers[o][k] = float(self.centers[o, k])
        self._lib = sensor._lib
        self._h = C.c_void_p()
        abi.check(self._lib, self._lib.dbot_b200_filter_create(sensor._h, C.byref(cfg), C.byref(self._h)))
        from .rb_sensor import _LIVE_FILTERS

        _LIVE_FILTERS.add(self)
        self.resample_count = 0

    def initialize(self, initial_poses):
        poses = np.ascontiguousarray(initial_poses, dtype=np.float64).reshape(self.n_obj, 6)
        self._abi.check(self._lib, self._lib.dbot_b200_filter_initialize(self._h, self._abi.ptr(poses)))
        self.resample_count = 0

    def track(self, image):
        """-> [n_obj][12], as ParticleTracker.track"""
        img = np.ascontiguousarray(image, dtype=np.float32).reshape(-1)
        out = np.empty((self.n_obj, 12), dtype=np.float64)
        cnt = self._C.c_int32(0)
        self._abi.check(self._lib, self._lib.dbot_b200_filter_track(self._h, self._abi.ptr(img), self._abi.ptr(out),
                                                                   self._C.byref(cnt)))
        self.resample_count = cnt.value
        return out

    def close(self):
        if self._h:
            self._lib.dbot_b200_filter_destroy(self._h)
            self._h = self._C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
