"""Particle-sharded evaluator: one process per GPU, launched by torchrun.

Particles are independent, so rank g of G evaluates the contiguous slice [lo_g, hi_g) of every
`loglikes` call on its own sensor and owns the occlusion slots of exactly those particles.  The
depth frame, the mesh and the camera are replicated (every rank calls set_observation with the
same image).  The one exchange step of the evaluation is an all-gather of the log-weights, so that
every rank -- each runs the same filter -- sees all N weights and takes the same resampling
decision.

`indices` keeps the reference's meaning (occlusion slot of each particle's parent, global
numbering).  After a resampling step a child may have its parent on another rank
(SURVEY.md section 8f row N2): the parent's occlusion image then migrates first -- the owner
exports the slot, it travels point to point (NCCL send/recv over NVLink on GPUs), the receiver
imports it as an extra local slot that the child names as its parent; like every parent slot it is
dropped (or adopted) by the next update.  Every rank sees the same global `indices`, so every rank
derives the same transfer schedule without negotiation.  `allow_migration=False` refuses such calls
instead (CrossShardParent), which is what BASELINE config 4 (shard-local parents) needs.

The local sensor is anything with reset / set_observation / loglikes_poses / compose and the slot
export / import pair -- in the product the B200KinectImageModel (device-pointer variants); the CPU
test-suite injects the oracle to check the host logic under gloo.
"""
from __future__ import annotations

import ctypes
import os

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


def shard_owner(slot, n, world):
    """rank that owns item `slot` of n (inverse of shard_bounds)"""
    base, extra = divmod(int(n), int(world))
    cut = extra * (base + 1)
    if slot < cut:
        return int(slot // (base + 1))
    return int(extra + (slot - cut) // max(base, 1))


class ShardedRbSensor(RbSensor):
    def __init__(self, local_sensor, group=None, device=None, allow_migration=True):
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
        self.allow_migration = allow_migration
        self.n_slots = 1          # global slot count; after reset every rank holds slot 0
        self._buf = None
        self._out_all = None
        # the library's one-pass check of a call's parent vector (host code, needs no GPU)
        lib = getattr(local_sensor, "_lib", None)
        if lib is None:
            try:
                from . import abi

                lib = abi.load_library()
            except (ImportError, OSError):
                lib = None      # a test double as the local sensor on a machine without the built library
        self._scan = getattr(lib, "dbot_b200_shard_scan", None)
        # parents imported since the last update: for EVERY rank the set of global slots it holds a
        # copy of (all ranks track all sets: the schedule must come out the same everywhere), and for
        # this rank the local slot index of each
        self._imported_keys = [set() for _ in range(self.world)]
        self._imported_local = {}
        self.migrated_slots = 0   # slots this rank has received so far (statistics)
        from .rb_sensor import _LIVE_WRAPPERS

        _LIVE_WRAPPERS.add(self)
        if self.world > 2 and hasattr(self.local, "set_staging_threads"):
            # ranks run in lock step (every call ends in a gather): a helper thread that gets a core late on ONE
            # rank holds up all of them.  Measured end to end: helpers +10 % with 2 ranks, +-0 with 4, -20 % with
            # 8 -- beyond two ranks every rank stages its depth frames on the calling thread alone
            self.local.set_staging_threads(0)
        self.gather = "collective"   # or "fused": peer stores from the particle kernel (csrc/peer_gather.cuh)
        self.gather_reason = None
        self._connect_peers()

    def _connect_peers(self):
        """GPUs: share every rank's gathered vector with every other rank as CUDA IPC memory, so that the particle
        kernel's epilogue stores the log-weights straight into all of them over NVLink and no collective is
        launched per call.  Collective itself (every rank constructs its ShardedRbSensor); all ranks end up in the
        same mode: if one of them cannot map a peer, all use the NCCL all-gather."""
        import torch.distributed as dist

        if self.device.type != "cuda" or not hasattr(self.local, "peer_gather_export") or self.world < 2:
            return
        if os.environ.get("DBOT_B200_SHARDED_GATHER", "fused") != "fused":
            self.gather_reason = "DBOT_B200_SHARDED_GATHER asks for the collective"
            return
        width_max = int(getattr(self.local, "max_sample_count"))
        handle, why = b"", None
        try:
            handle = self.local.peer_gather_export(self.world, self.rank, width_max)
        except Exception as e:  # noqa: BLE001 -- any failure means "use the collective", on every rank
            why = str(e)
        got = [None] * self.world
        dist.all_gather_object(got, (handle, why), group=self.group)
        if all(h for h, _ in got):
            try:
                self.local.peer_gather_connect([h for h, _ in got])
            except Exception as e:  # noqa: BLE001
                why = str(e)
        else:
            why = why or next(w for h, w in got if not h)
        agreed = [None] * self.world
        dist.all_gather_object(agreed, why, group=self.group)
        if all(w is None for w in agreed):
            self.gather = "fused"
        else:
            self.gather_reason = next(w for w in agreed if w is not None)

    def close(self):
        """Release the torch buffers that were used on the local sensor's stream while that stream still exists:
        page-locked host tensors remember the streams they were copied on, and releasing them after the library
        has destroyed its stream aborts the process.  Called from the package's atexit hook as well."""
        if self._buf is not None and self.device.type == "cuda":
            import torch

            try:
                self.local.synchronize()
            except Exception:
                pass
            torch.cuda.synchronize(self.device)
        for name in ("_h_poses", "_h_idx", "_d_poses", "_d_idx", "_d_out", "_buf", "_h_all", "_ext", "_send_host",
                     "_recv_host", "_send"):
            if hasattr(self, name):
                setattr(self, name, None)

    # -- RbSensor ---------------------------------------------------------------------------
    def reset(self):
        self.local.reset()
        self.n_slots = 1
        self._forget_imports()

    def set_observation(self, image):
        self.local.set_observation(image)

    def loglikes(self, deltas, indices, update=False):
        return self.loglikes_poses(self.local.compose(deltas), indices, update)

    # RbSensor::integrated_poses / default poses live in the local sensor (it composes the deltas):
    # a tracker that assigns them on the sharded sensor must reach that one
    @property
    def integrated_poses(self):
        return self.local.integrated_poses

    @integrated_poses.setter
    def integrated_poses(self, poses):
        self.local.integrated_poses = poses

    @property
    def n_objects(self):
        return self.local.n_objects

    def compose(self, deltas):
        return self.local.compose(deltas)

    # -- pieces ------------------------------------------------------------------------------
    def _forget_imports(self):
        self._imported_keys = [set() for _ in range(self.world)]
        self._imported_local = {}

    def transfer_schedule(self, indices, n):
        """{(owner, needer): sorted global parent slots} for the parents that live on another rank
        than their children and have not been imported there yet -- identical on every rank"""
        sched = {}
        if self.n_slots == 1:
            return sched
        # fast path (the common case): every parent lives on its child's rank
        key = (n, self.n_slots)
        if getattr(self, "_sched_key", None) != key:      # the partition tables only change with n / n_slots
            c_hi = np.array([shard_bounds(n, self.world, r)[1] for r in range(self.world)])
            self._p_hi = np.array([shard_bounds(self.n_slots, self.world, r)[1] for r in range(self.world)])
            self._child_rank = np.searchsorted(c_hi, np.arange(n), side="right")
            self._sched_key = key
        parent_rank = np.searchsorted(self._p_hi, indices, side="right")
        if np.array_equal(self._child_rank, parent_rank):
            return sched
        for r in range(self.world):
            lo, hi = shard_bounds(n, self.world, r)
            plo, phi = shard_bounds(self.n_slots, self.world, r)
            par = np.unique(np.asarray(indices[lo:hi], dtype=np.int64))
            for s in par[(par < plo) | (par >= phi)]:
                s = int(s)
                if s in self._imported_keys[r]:
                    continue
                sched.setdefault((shard_owner(s, self.n_slots, self.world), r), []).append(s)
        return sched

    def _export(self, local_slots):
        """stored images of local slots as one [k, slot_floats] tensor on self.device"""
        import torch

        F = self.local.slot_floats
        if self.device.type == "cuda":
            buf = torch.empty((len(local_slots), F), dtype=torch.float32, device=self.device)
            for j, s in enumerate(local_slots):
                self.local.export_slot_device(s, buf[j].data_ptr())
            self.local.synchronize()       # exports run on the library's stream, the send on torch's
            return buf
        return torch.from_numpy(np.stack([np.asarray(self.local.export_slot(s), dtype=np.float32)
                                          for s in local_slots]))

    def _import(self, buf):
        import torch

        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()   # the receive has landed
            out = [self.local.import_slot_device(buf[j].data_ptr()) for j in range(buf.shape[0])]
            self.local.synchronize()       # buf may be released after this
            return out
        return [self.local.import_slot(buf[j].numpy()) for j in range(buf.shape[0])]

    def migrate(self, indices, n):
        """move every remote parent of this rank's children here (point to point, one message per
        owner/needer pair); returns the number of slots received"""
        import torch
        import torch.distributed as dist

        sched = self.transfer_schedule(indices, n)
        if not sched:
            return 0
        if not self.allow_migration:
            (o, r), slots = sorted(sched.items())[0]
            raise CrossShardParent(
                "rank %d needs occlusion slot %d, which rank %d owns; migration is disabled" % (r, slots[0], o))
        F = self.local.slot_floats
        ops, recv = [], []
        keep = []
        plo, _ = shard_bounds(self.n_slots, self.world, self.rank)
        for (o, r), slots in sorted(sched.items()):
            if o == self.rank:
                buf = self._export([s - plo for s in slots])
                keep.append(buf)
                ops.append(dist.P2POp(dist.isend, buf, self._peer(r), group=self.group))
            elif r == self.rank:
                buf = torch.empty((len(slots), F), dtype=torch.float32, device=self.device)
                recv.append((slots, buf))
                ops.append(dist.P2POp(dist.irecv, buf, self._peer(o), group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        got = 0
        for slots, buf in recv:
            for s, local_idx in zip(slots, self._import(buf)):
                self._imported_local[s] = local_idx
            got += len(slots)
        for (o, r), slots in sched.items():
            self._imported_keys[r].update(slots)
        self.migrated_slots += got
        return got

    def _peer(self, r):
        import torch.distributed as dist

        return r if self.group is None else dist.get_global_rank(self.group, r)

    def local_parents(self, indices, n):
        """global parent slots of this rank's particles -> local slot numbers (after migrate())"""
        lo, hi = shard_bounds(n, self.world, self.rank)
        parents = np.asarray(indices[lo:hi], dtype=np.int64)
        if self.n_slots == 1:
            return lo, hi, np.zeros(hi - lo, dtype=np.int32)
        plo, phi = shard_bounds(self.n_slots, self.world, self.rank)
        out = (parents - plo).astype(np.int32)
        for k in np.nonzero((parents < plo) | (parents >= phi))[0]:   # migrated parents (rare)
            out[k] = self._imported_local[int(parents[k])]
        return lo, hi, out

    def all_gather(self, local_values, n):
        """all-gather of ragged shards: pad to the largest shard, one collective, unpad.  Buffers are
        kept between calls (page-locked on the host when the collective runs on a GPU)."""
        import torch
        import torch.distributed as dist

        counts = shard_counts(n, self.world)
        width = max(counts) if counts else 0
        if width == 0:
            return np.zeros(0, dtype=np.float64)
        if self._buf is None or self._buf.numel() != width * self.world:
            cuda = self.device.type == "cuda"
            self._send_host = torch.zeros(width, dtype=torch.float64, pin_memory=cuda)
            self._recv_host = torch.zeros(width * self.world, dtype=torch.float64, pin_memory=cuda)
            self._send = self._send_host.to(self.device) if cuda else self._send_host
            self._buf = torch.empty(width * self.world, dtype=torch.float64, device=self.device)
        self._send_host[: len(local_values)] = torch.from_numpy(np.ascontiguousarray(local_values, dtype=np.float64))
        if self.device.type == "cuda":
            self._send.copy_(self._send_host, non_blocking=True)
            dist.all_gather_into_tensor(self._buf, self._send, group=self.group)
            self._recv_host.copy_(self._buf, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            got = self._recv_host.numpy().reshape(self.world, width)
        else:
            dist.all_gather_into_tensor(self._buf, self._send, group=self.group)
            got = self._buf.numpy().reshape(self.world, width)
        return np.concatenate([got[r, : counts[r]] for r in range(self.world)])

    def _evaluate_and_gather_on_device(self, poses, lo, hi, local_idx, n, update):
        """CUDA path: this rank's slice goes to the device once, the kernel's output is all-gathered where it
        lies (NCCL over NVLink), the gathered vector comes back with one copy -- one stream, one wait per call"""
        import torch
        import torch.distributed as dist

        counts = shard_counts(n, self.world)
        width = max(counts)
        cnt = hi - lo
        n_obj = self.local.n_objects
        if self._buf is None or self._buf.numel() < width * self.world:
            cap = max(width, getattr(self.local, "max_sample_count", width))
            self._h_poses = torch.zeros((cap, n_obj, 12), dtype=torch.float32).pin_memory()
            self._h_idx = torch.zeros(cap, dtype=torch.int32).pin_memory()
            self._d_poses = torch.zeros((cap, n_obj, 12), dtype=torch.float32, device=self.device)
            self._d_idx = torch.zeros(cap, dtype=torch.int32, device=self.device)
            self._d_out = torch.zeros(cap, dtype=torch.float64, device=self.device)
            self._buf = torch.zeros(cap * self.world, dtype=torch.float64, device=self.device)
            self._h_all = torch.zeros(cap * self.world, dtype=torch.float64).pin_memory()
            self._ext = torch.cuda.ExternalStream(self.local.stream, device=self.device)
        if cnt:
            self._h_poses[:cnt] = torch.from_numpy(poses[lo:hi])
            self._h_idx[:cnt] = torch.from_numpy(local_idx)
        with torch.cuda.stream(self._ext):
            if cnt:
                self._d_poses[:cnt].copy_(self._h_poses[:cnt], non_blocking=True)
                self._d_idx[:cnt].copy_(self._h_idx[:cnt], non_blocking=True)
                self.local.loglikes_device(self._d_poses.data_ptr(), self._d_idx.data_ptr(), cnt, update,
                                           self._d_out.data_ptr())
            elif update:
                self.local.update_empty()
            dist.all_gather_into_tensor(self._buf[: width * self.world], self._d_out[:width], group=self.group)
            self._h_all[: width * self.world].copy_(self._buf[: width * self.world], non_blocking=True)
        self.local.synchronize()   # the one wait of the call; also surfaces device-side errors
        got = self._h_all[: width * self.world].numpy().reshape(self.world, width)
        return np.concatenate([got[r, : counts[r]] for r in range(self.world)])

    def _evaluate_and_gather_fused(self, poses, lo, hi, local_idx, n, update):
        """CUDA path with mapped peers: one library call stages this rank's slice, runs the particle kernel --
        whose epilogue stores every log-weight into every rank's gathered vector over NVLink --, exchanges the
        arrival counters and copies the gathered vector back.  No torch op, no collective launch."""
        width = -(-n // self.world)
        if self._out_all is None or self._out_all.size != width * self.world:
            self._out_all = np.empty(width * self.world, dtype=np.float64)
        self.local.loglikes_gather(poses[lo:hi], np.ascontiguousarray(local_idx, dtype=np.int32), update, width,
                                   self._out_all)
        if n == width * self.world:
            return self._out_all.copy()
        counts = shard_counts(n, self.world)
        got = self._out_all.reshape(self.world, width)
        return np.concatenate([got[r, : counts[r]] for r in range(self.world)])

    def loglikes_poses(self, poses, indices, update=False):
        n_obj = self.local.n_objects
        poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, n_obj, 12)
        n = poses.shape[0]
        if not (isinstance(indices, np.ndarray) and indices.dtype == np.int32 and indices.size == n):
            raise TypeError("loglikes: indices must be an int32 numpy array with one entry per state")
        if self._scan is not None and indices.flags.c_contiguous:
            # the library's one pass over the parents: range check + "does any child need a remote parent?"
            rc = self._scan(indices.ctypes.data_as(ctypes.c_void_p), n, self.n_slots, self.world, None)
            if rc < 0:
                raise IndexError("loglikes: parent slot out of range (have %d)" % self.n_slots)
            if rc == 1:
                self.migrate(indices, n)
            all_local = rc == 0 and not self._imported_local
        else:
            if np.any(indices < 0) or np.any(indices >= self.n_slots):
                raise IndexError("loglikes: parent slot out of range (have %d)" % self.n_slots)
            self.migrate(indices, n)
            all_local = False
        if all_local:
            # the common case, without the general path's bookkeeping: this rank's parents are its own slots
            lo, hi = shard_bounds(n, self.world, self.rank)
            if self.n_slots == 1:
                local_idx = np.zeros(hi - lo, dtype=np.int32)
            else:
                local_idx = indices[lo:hi] - np.int32(shard_bounds(self.n_slots, self.world, self.rank)[0])
        else:
            lo, hi, local_idx = self.local_parents(indices, n)
        if self.gather == "fused":
            out = self._evaluate_and_gather_fused(poses, lo, hi, local_idx, n, update)
        elif self.device.type == "cuda" and hasattr(self.local, "loglikes_device"):
            out = self._evaluate_and_gather_on_device(poses, lo, hi, local_idx, n, update)
        else:
            if hi > lo:
                mine = self.local.loglikes_poses(poses[lo:hi], local_idx, update)
            else:
                mine = np.zeros(0)
                if update and hasattr(self.local, "update_empty"):
                    self.local.update_empty()   # stay in step with the shards that did get particles
            out = self.all_gather(mine, n)
        if update:
            indices[:] = np.arange(n, dtype=np.int32)
            self.n_slots = n
            self._forget_imports()
        return out
