"""-m gpu: the particle-sharded evaluator over NCCL on 2 GPUs (skipped on a 1-GPU box): sharding
must not change a single bit of any log-weight, with shard-local parents and
with parents that migrate between the GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

WORKER = r'''
import faulthandler, os, sys
faulthandler.enable()
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(sys.argv[0]))) if False else os.environ["DBOT_ROOT"])
from dbot_ros_b200 import CameraData, ObjectModel, Parameters, RbSensorBuilder, synth
from dbot_ros_b200.sharded import ShardedRbSensor, shard_bounds
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n, ds = 501, 2
rows, cols = synth.resolution(ds); K = synth.camera_matrix(ds); V, T = synth.mesh_set("icosphere5k")
build = lambda cnt, dev: RbSensorBuilder(ObjectModel(V, T), CameraData(K, rows, cols), Parameters(sample_count=cnt), device_id=dev).build()
local = build(n, lr)
sharded = ShardedRbSensor(local)
print("rank %d gather: %s (%s)" % (rank, sharded.gather, sharded.gather_reason), flush=True)
single = build(n, lr) if rank == 0 else None
rng = np.random.default_rng(5)
gt = synth.ground_truth_poses(1, rng)
ok = True
idx = np.zeros(n, dtype=np.int32)
for f in range(5):
    frame = synth.make_frame(local.render, gt, rows, cols, rng)
    poses = synth.particle_poses(gt, n, rng)
    if f in (1, 2):      # shard-local parents (BASELINE config 4)
        idx = np.empty(n, dtype=np.int32)
        for r in range(world):
            lo, hi = shard_bounds(n, world, r)
            idx[lo:hi] = rng.integers(lo, hi, size=hi - lo)
    elif f > 2:          # global resampling: occlusion images migrate over NVLink
        idx = np.sort(rng.integers(0, n, size=n)).astype(np.int32)
    i1, i2 = idx.copy(), idx.copy()
    print("rank", rank, "frame", f, "migrated so far", sharded.migrated_slots, file=sys.stderr, flush=True)
    sharded.set_observation(frame)
    got = sharded.loglikes_poses(poses, i1, True)
    if rank == 0:
        single.set_observation(frame)
        want = single.loglikes_poses(poses, i2, True)
        ok = ok and np.array_equal(got, want) and np.array_equal(i1, i2)
moved = torch.tensor([sharded.migrated_slots], device="cuda")
dist.all_reduce(moved)
ok = ok and int(moved.item()) > 0
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
print("rank %d bit-identical to one GPU: %s slots migrated (all ranks): %d\n" % (rank, bool(ok), int(moved.item())), end="", flush=True)
sharded.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 3)
'''


@pytest.mark.parametrize("gather", ["fused", "collective"])
def test_two_gpu_sharding_is_bit_identical(tmp_path, gather):
    """gather = fused: the particle kernel stores the log-weights into both ranks' gathered vectors (CUDA IPC
    memory over NVLink, csrc/peer_gather.cuh); collective: ncclAllGather through torch.distributed"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DBOT_ROOT=os.path.dirname(HERE), DBOT_B200_SHARDED_GATHER=gather)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "bit-identical to one GPU: True" in r.stdout, r.stdout[-2000:]
    assert "rank 0 gather: %s" % gather in r.stdout and "rank 1 gather: %s" % gather in r.stdout, r.stdout[-2000:]
    print(r.stdout[-600:])


def test_slot_export_import_between_two_handles_on_one_gpu():
    """the kernels behind cross-GPU resampling (dbot_b200_export_slot_device / import_slot_device) between two
    handles on ONE GPU, so that a one-GPU box exercises them: a slot exported from A and imported into B is
    an extra parent slot of B whose children see exactly A's occlusion image; the next update drops it."""
    import torch

    from dbot_ros_b200 import CameraData, DbotB200Error, ObjectModel, Parameters, RbSensorBuilder, synth

    rows, cols = synth.resolution(4)
    K = synth.camera_matrix(4)
    V, T = synth.mesh_set("uvsphere480")
    n = 24
    mk = lambda: RbSensorBuilder(ObjectModel(V, T), CameraData(K, rows, cols), Parameters(sample_count=n)).build()  # noqa: E731
    a, b = mk(), mk()
    rng = np.random.default_rng(9)
    gt = synth.ground_truth_poses(1, rng)
    frames = [synth.make_frame(a.render, gt, rows, cols, rng) for _ in range(3)]
    pa, pb = synth.particle_poses(gt, n, rng), synth.particle_poses(gt, n, rng, sigma_t=0.02)
    # same observation / update history on both handles (that is what makes stored images portable),
    # different particles: their occlusion images differ
    for s, p in ((a, pa), (b, pb)):
        s.set_observation(frames[0])
        s.loglikes_poses(p, np.zeros(n, np.int32), True)
        s.set_observation(frames[1])
    assert not np.array_equal(a.occlusion(5), b.occlusion(5))
    buf = torch.empty(a.slot_floats, dtype=torch.float32, device="cuda")
    free_before = b.free_slots
    a.export_slot_device(5, buf.data_ptr())
    a.synchronize()
    slot = b.import_slot_device(buf.data_ptr())
    b.synchronize()
    assert slot == n and b.n_slots == n + 1 and b.free_slots == free_before - 1
    assert np.array_equal(b.occlusion(slot), a.occlusion(5))
    probe = pa[7:8]
    want = a.loglikes_poses(probe, np.array([5], np.int32), False)
    got = b.loglikes_poses(probe, np.array([slot], np.int32), False)
    assert np.array_equal(got, want)
    # children of the imported slot in an update; afterwards it is gone like any parent
    idx = np.full(n, slot, np.int32)
    idx[::2] = np.arange(n, dtype=np.int32)[::2]
    ia = idx.copy()
    ia[ia == slot] = 5
    ra = a.loglikes_poses(pa, ia, True)
    rb = b.loglikes_poses(pa, idx, True)
    assert np.array_equal(ra[1::2], rb[1::2])           # the odd particles' parent images were identical
    assert b.n_slots == n and b.free_slots == free_before
    assert np.array_equal(a.occlusion(1), b.occlusion(1))
    with pytest.raises(DbotB200Error):
        b.export_slot_device(n, buf.data_ptr())
