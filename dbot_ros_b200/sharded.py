"""Particle-sharded evaluator: one process per GPU, launched by torchrun.

Particles are independent, so rank g of G evaluates the contiguous slice [lo_g, hi_g) of every
`loglikes` call on its own sensor and owns the occlusion slots of exactly those particles.  The
depth frame, the mesh and the camera are replicated (every rank calls set_observation with the
same image).  The one exchange step of the path is an all-gather of the log-weights, so that every
rank -- each runs the same filter -- sees all N weights and takes the same resampling decision.
There is no other collective on the data path.

`indices` keeps the reference's meaning (occlusion slot of each particle's parent, global
numbering; ref semantics in SURVEY.md A.5).  A parent slot that lives on another rank would need
its 4*P-byte occlusion image moved over NVLink first; this class refuses such a call with
CrossShardParent instead of computing something else.

The local sensor is anything with reset / set_observation / loglikes_poses / compose -- in the
product the B200KinectImageModel; the CPU test-suite injects the oracle to check the host logic
under gloo.
"""
from __future__ import annotations

import numpy as np

from .rb_sensor import RbSensor


class CrossShardParent(RuntimeError):
    pass


def shard_bounds(n, world, rank):
    """contiguous split of n items: the first n % world ranks hold one more"""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_counts(n, world):
    return [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]


class ShardedRbSensor(RbSensor):
    def __init__(self, local_sensor, group=None, device=None):
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("ShardedRbSensor needs an initialised torch.distributed process group")
        self.local = local_sensor
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        backend = dist.get_backend(group)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        self.device = device
        self.n_slots = 1          # global slot count; after reset every rank holds slot 0
        self._buf = None

    # -- RbSensor ---------------------------------------------------------------------------
    def reset(self):
        self.local.reset()
        self.n_slots = 1

    def set_observation(self, image):
        self.local.set_observation(image)

    def loglikes(self, deltas, indices, update=False):
        return self.loglikes_poses(self.local.compose(deltas), indices, update)

    # -- pieces ------------------------------------------------------------------------------
    def local_parents(self, indices, n):
        """global parent slots of this rank's particles -> local slot numbers"""
        lo, hi = shard_bounds(n, self.world, self.rank)
        parents = np.asarray(indices[lo:hi], dtype=np.int64)
        if np.any(parents < 0) or np.any(parents >= self.n_slots):
            raise IndexError("loglikes: parent slot out of range (have %d)" % self.n_slots)
        if self.n_slots == 1:
            return lo, hi, np.zeros(hi - lo, dtype=np.int32)
        plo, phi = shard_bounds(self.n_slots, self.world, self.rank)
        if np.any(parents < plo) or np.any(parents >= phi):
            bad = parents[(parents < plo) | (parents >= phi)][0]
            raise CrossShardParent(
                "rank %d owns occlusion slots [%d,%d) but a particle of its shard has parent %d; "
                "cross-rank occlusion-row migration is not implemented" % (self.rank, plo, phi, bad))
        return lo, hi, (parents - plo).astype(np.int32)

    def all_gather(self, local_values, n):
        """all-gather of ragged shards: pad to the largest shard, one collective, unpad"""
        import torch
        import torch.distributed as dist

        counts = shard_counts(n, self.world)
        width = max(counts) if counts else 0
        if width == 0:
            return np.zeros(0, dtype=np.float64)
        send = torch.zeros(width, dtype=torch.float64)
        send[: len(local_values)] = torch.from_numpy(np.ascontiguousarray(local_values, dtype=np.float64))
        send = send.to(self.device)
        if self._buf is None or self._buf.numel() != width * self.world:
            self._buf = torch.empty(width * self.world, dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(self._buf, send, group=self.group)
        got = self._buf.cpu().numpy().reshape(self.world, width)
        return np.concatenate([got[r, : counts[r]] for r in range(self.world)])

    def loglikes_poses(self, poses, indices, update=False):
        n_obj = self.local.n_objects
        poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, n_obj, 12)
        n = poses.shape[0]
        if not (isinstance(indices, np.ndarray) and indices.dtype == np.int32 and indices.size == n):
            raise TypeError("loglikes: indices must be an int32 numpy array with one entry per state")
        lo, hi, local_idx = self.local_parents(indices, n)
        mine = self.local.loglikes_poses(poses[lo:hi], local_idx, update) if hi > lo else np.zeros(0)
        out = self.all_gather(mine, n)
        if update:
            indices[:] = np.arange(n, dtype=np.int32)
            self.n_slots = n
        return out
