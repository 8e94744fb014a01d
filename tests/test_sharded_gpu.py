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
import os, sys
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
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 3)
'''


def test_two_gpu_sharding_is_bit_identical(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DBOT_ROOT=os.path.dirname(HERE))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
