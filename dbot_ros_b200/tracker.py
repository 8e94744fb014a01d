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

    def __init__(self, sensor, n_objects, params: TrackerParameters, seed=0):
        self.sensor = sensor
        self.n_obj = int(n_objects)
        self.p = params
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
        """initial_poses: [n_obj][6] position + rotation vector in the camera frame"""
        self.sensor.integrated_poses = np.asarray(initial_poses, dtype=np.float64).reshape(self.n_obj, 6)
        self.states[:] = 0.0
        self.log_w[:] = -np.log(self.n)
        self.indices[:] = 0
        self.loglikes[:] = 0.0
        self.sensor.reset()

    # -- RBC-PF::filter -----------------------------------------------------------------------
    def kl_given_uniform(self):
        w = np.exp(self.log_w)
        return float(np.log(self.n) + np.sum(w * self.log_w))

    def resample(self):
        w = np.exp(self.log_w)
        parents = systematic_resample(w / w.sum(), self.n, self.rng.random())
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
            self.noise[:, block, :] = self.rng.standard_normal((self.n, 6))
            self.next_states = self.propagate(self.states, self.noise)
            update = block == self.n_obj - 1
            new = self.sensor.loglikes(self.next_states.reshape(self.n, -1), self.indices, update)
            self.log_w += new - self.loglikes
            self.log_w -= np.logaddexp.reduce(self.log_w)
            self.loglikes = new
            if self.kl_given_uniform() > self.p.max_kl_divergence:
                self.resample()
        self.states = self.next_states

    # -- ParticleTracker::on_track + Tracker::track -------------------------------------------
    def track(self, image):
        from .rb_sensor import rotation_matrices

        self.filter(image)
        w = np.exp(self.log_w)
        mean = np.tensordot(w, self.states[:, :, 0:6], axes=(0, 0))            # [n_obj][6]
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
        a = self.p.moving_average_update_rate
        self.smoothed = poses.copy() if self.smoothed is None or a >= 1.0 else (1 - a) * self.smoothed + a * poses
        return self.smoothed.copy()


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
