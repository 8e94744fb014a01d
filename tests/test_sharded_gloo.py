"""world_size-2 gloo test of the particle-sharded evaluator: sharding must not
change a single bit of any log-weight, the ragged all-gather must put every weight at its particle's
index on every rank, and cross-shard parents must be refused."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib
import sharded_worker
from dbot_ros_b200 import synth
from dbot_ros_b200.sharded import shard_bounds, shard_counts, shard_owner

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_partition():
    for n in (0, 1, 2, 7, 64, 1000, 80000):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            c = shard_counts(n, world)
            assert sum(c) == n and max(c) - min(c) <= 1
            for slot in set([0, n // 3, n // 2, max(n - 1, 0)]) if n else ():
                r = shard_owner(slot, n, world)
                assert b[r][0] <= slot < b[r][1]


@pytest.mark.parametrize("n", [37, 64])  # ragged and even shards
def test_two_ranks_match_one_process(tmp_path, n):
    frames, seed, world = 3, 11, 2
    out = str(tmp_path / "res")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(HERE, "sharded_worker.py"), out, str(n), str(frames), str(seed)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    # single-process run of the same scenario on the plain oracle
    rows, cols, K, V, T, _, _ = sharded_worker.scenario(n, frames, seed)
    orc = oracle_lib.Oracle(rows, cols, K, V, T, synth.default_parameters())
    want = sharded_worker.run(orc, n, frames, world, seed, orc.render)
    orc.reset()
    want_global = sharded_worker.run(orc, n, frames, world, seed + 1, orc.render, global_parents=True)
    single = sharded_worker.OracleLocal(rows, cols, K, V, T, synth.default_parameters())
    want_track, want_resamplings = sharded_worker.track(single, single.render, 60, 6, seed)
    assert np.linalg.norm(want_track[-1, :3] - [0.01, -0.005, 0.8075]) < 0.004     # it does follow the object
    migrated = 0
    for rank in range(world):
        z = np.load(out + ".rank%d.npz" % rank)
        assert np.array_equal(z["got"], want), "rank %d differs from the unsharded evaluation" % rank
        # parents drawn from all slots: occlusion images migrated between the ranks, results unchanged
        assert np.array_equal(z["got_global"], want_global), "rank %d: cross-rank parents" % rank
        migrated += int(z["migrated"])
        # a ParticleTracker over the sharded sensor (it sets integrated_poses on it and calls loglikes(deltas))
        assert np.array_equal(z["tracked"], want_track) and int(z["resamplings"]) == want_resamplings
        assert z["refused"].all()
        lo, hi = z["lo_hi"]
        assert int(z["local_slots"]) == 1  # reset collapsed the local state again
        assert (lo, hi) == shard_bounds(n, world, rank)
    assert migrated > 0  # the scenario did move occlusion images between the ranks
