"""TEST INFRASTRUCTURE: one rank of the world_size-2 gloo test of dbot_ros_b200.sharded.  The local
evaluator is the CPU oracle (tests may use it; the package never does) so that the host logic of
the sharded sensor -- shard bounds, global->local parent slots, ragged all-gather, state hand-over
-- runs on a machine without GPUs."""
import os
import sys

import numpy as np
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import oracle_lib  # noqa: E402
from dbot_ros_b200 import synth  # noqa: E402
from dbot_ros_b200.sharded import CrossShardParent, ShardedRbSensor, shard_bounds  # noqa: E402


class OracleLocal(oracle_lib.Oracle):
    """the oracle with the two members of the RbSensor interface a tracker uses beyond loglikes_poses:
    integrated_poses and compose (what B200KinectImageModel offers on the product side)"""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.integrated_poses = np.zeros((self.n_objects, 6))

    def compose(self, deltas):
        from dbot_ros_b200.rb_sensor import compose_poses

        return compose_poses(deltas, self.integrated_poses)

    def loglikes(self, deltas, indices, update=False):
        return self.loglikes_poses(self.compose(deltas), indices, update)


def track(sensor, render, n_particles, frames, seed):
    """a ParticleTracker over `sensor` on a short moving-object sequence -> [frames, 12] estimates"""
    from dbot_ros_b200.tracker import ParticleTracker, TrackerParameters

    rows, cols = synth.resolution(8)
    rng = np.random.default_rng(seed)
    start, step = np.array([0.0, 0.0, 0.8]), np.array([0.002, -0.001, 0.0015])
    tr = ParticleTracker(sensor, 1, TrackerParameters(evaluation_count=n_particles), seed=seed)
    tr.initialize([[start[0] + 0.004, start[1] - 0.003, start[2] + 0.004, 0.0, 0.0, 0.0]])
    est = []
    for f in range(frames):
        img = synth.make_frame(render, [(np.eye(3), start + f * step)], rows, cols, rng, occluder=False)
        est.append(tr.track(img)[0])
    return np.array(est), tr.resample_count


def scenario(n, frames, seed):
    """inputs every rank (and the single-process check) rebuilds identically"""
    rows, cols = synth.resolution(8)
    K = synth.camera_matrix(8)
    V, T = synth.mesh_set("ico80")
    rng = np.random.default_rng(seed)
    gt = synth.ground_truth_poses(1, rng)
    return rows, cols, K, V, T, rng, gt


def shard_local_parents(n, world, rng):
    """resampling that keeps every parent inside its child's shard (BASELINE config 4)"""
    out = np.empty(n, dtype=np.int32)
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        out[lo:hi] = rng.integers(lo, hi, size=hi - lo)
    return out


def run(sensor, n, frames, world, seed, render, global_parents=False):
    rows, cols, K, V, T, rng, gt = scenario(n, frames, seed)
    idx = np.zeros(n, dtype=np.int32)
    outs = []
    for f in range(frames):
        frame = synth.make_frame(render, gt, rows, cols, rng)
        poses = synth.particle_poses(gt, n, rng)
        if f > 0:
            # global_parents: sorted draws from all slots, as systematic resampling over the whole
            # particle set produces them -- children get parents on other ranks
            idx = (np.sort(rng.integers(0, n, size=n)).astype(np.int32) if global_parents
                   else shard_local_parents(n, world, rng))
        sensor.set_observation(frame)
        outs.append(sensor.loglikes_poses(poses, idx, update=False))
        outs.append(sensor.loglikes_poses(poses, idx, update=True))
        assert np.array_equal(idx, np.arange(n))
    return np.stack(outs)


def main():
    out_path, n, frames, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rows, cols, K, V, T, _, _ = scenario(n, frames, seed)
    local = OracleLocal(rows, cols, K, V, T, synth.default_parameters())
    sensor = ShardedRbSensor(local)
    got = run(sensor, n, frames, world, seed, local.render)
    # with migration disabled a parent on the other rank must be refused, not silently evaluated
    refused = False
    bad = np.arange(n, dtype=np.int32)[::-1].copy()
    strict = ShardedRbSensor(local, allow_migration=False)
    strict.n_slots = sensor.n_slots
    try:
        strict.loglikes_poses(np.zeros((n, 1, 12), np.float32), bad, update=False)
    except CrossShardParent:
        refused = True
    # cross-rank resampling: parents migrate between the ranks
    sensor.reset()
    got_global = run(sensor, n, frames, world, seed + 1, local.render, global_parents=True)
    migrated = sensor.migrated_slots
    flags = [None] * world
    dist.all_gather_object(flags, refused)
    # the tracker drives the sharded sensor like any RbSensor: it assigns integrated_poses on it and calls
    # loglikes(deltas): every rank runs the same filter and must see the track of the unsharded sensor
    sensor.reset()
    tracked, resamplings = track(sensor, local.render, 60, 6, seed)
    sensor.reset()
    assert sensor.n_slots == 1
    np.savez(out_path + ".rank%d.npz" % rank, got=got, refused=np.array(flags), lo_hi=np.array(shard_bounds(n, world, rank)),
             local_slots=local.n_slots, got_global=got_global, migrated=migrated, tracked=tracked,
             resamplings=resamplings)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
