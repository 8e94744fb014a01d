#!/usr/bin/env python
"""bench.py -- particle-likelihoods/s of the depth observation model (BASELINE.json metric).

A step is one depth frame through the sensor exactly as the particle filter drives it for a
single-object tracker: set_observation(frame) then loglikes(poses, indices, update=true) for all
particles (ref:source/dbot_ros/object_tracker_ros.hpp:49 -> dbot filter).

  value   device-resident: frames, poses and parent indices already in HBM, the two kernels
          launched through the *_device C-ABI entry points, timed with CUDA events on the
          library's stream.  The K timed steps are repeated (at least REPEATS times, until a quarter of a second of
          device time has passed); value is the median repeat.
  e2e     the same step through the reference-facing host call with ordinary (pageable) host numpy
          buffers in and host log-weights out -- what the C++ adapter hands over; H2D/D2H inside the
          timed region.  e2e_pinned: the same from page-locked buffers.
  roofline  the fused particle kernel (update).  achieved = the bytes the launch actually has to move
          (copying children x 8 B x pixels of the parent's touched rectangle + covered pixels x 8 B per particle) / launch duration from CUDA
          events; frac_dram = DRAM bytes ncu counted for the same launch / duration / peak; the
          dense-formulation figure of SURVEY.md 8(d) (8P per particle whether copied or not) is kept
          beside it as dense_equivalent.  `bound` says what ncu shows limits the kernel.
  cpu_baseline / --impl reference   the CPU oracle (oracle/, lazily propagated occlusion state as in
          dbot's CPU model, OpenMP over particles, all host cores) on a bounded particle sample of
          the same frames.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from dbot_ros_b200 import CameraData, ObjectModel, Parameters, RbSensorBuilder, synth  # noqa: E402

METRIC = "particle-likelihoods/sec"  # BASELINE.json: "particle-likelihoods/sec & FPS @640x480, 5k-tri mesh, ..." (FPS is the line's "fps" key)
UNIT = "particle-likelihoods/s"
# BASELINE.json configs[1] (1000 particles, 640x480, single 5k-tri mesh) is the workload the metric is quoted on and
# what the driver runs.  DBOT_BENCH_CONFIG=3 selects configs[2] (10 000 particles; under torchrun that is configs[3]:
# 10 000 particles per GPU, 80 000 on eight) for the same line.
CONFIG_ID = int(os.environ.get("DBOT_BENCH_CONFIG", "2"))
if CONFIG_ID == 4:
    CONFIG_ID = 3        # the same per-GPU workload
if CONFIG_ID not in (2, 3):
    raise SystemExit("bench.py: DBOT_BENCH_CONFIG must be 2, 3 or 4")
N_FRAMES = 4             # distinct synthetic frames / pose sets cycled through the steps
REPEATS = 5              # the K timed steps are run this many times; the median repeat is reported


def workload(config_id=CONFIG_ID):
    mesh, n, ds = synth.CONFIGS[config_id]
    rows, cols = synth.resolution(ds)
    V, T = synth.mesh_set(mesh)
    return dict(mesh=mesh, n=n, ds=ds, rows=rows, cols=cols, V=V, T=T, K=synth.camera_matrix(ds),
                tris=int(sum(len(t) for t in T)))


def describe(w, n_gpus):
    return {"workload": "config%d: %d particles/GPU, %dx%d synthetic depth, %s (%d triangles), "
                        "set_observation + loglikes(update=true) per step"
                        % (CONFIG_ID, w["n"], w["cols"], w["rows"], w["mesh"], w["tris"]),
            "particles_per_gpu": w["n"], "particles_total": w["n"] * n_gpus,
            "resolution": [w["cols"], w["rows"]], "triangles": w["tris"],
            "occlusion_mode": "dense", "parallelism": "particles sharded x%d" % n_gpus,
            "parents": "systematic resampling of Gamma(0.5) weights on every step (about 50 % distinct parents)",
            "l2_policy": "inputs larger than L2 (%.2f GB occlusion state read + written per step)"
                         % (2 * w["n"] * w["rows"] * w["cols"] * 4 / 1e9)}


def resampled_parents(n, rng):
    """systematic resampling of random weights: the parent pattern the filter produces"""
    wgt = rng.gamma(0.5, size=n)
    cdf = np.cumsum(wgt / wgt.sum())
    u = (rng.random() + np.arange(n)) / n
    return np.minimum(np.searchsorted(cdf, u), n - 1).astype(np.int32)


def make_inputs(w, render, rank):
    """One scene for the whole job -- the sharded filter tracks ONE object and every rank sees the same depth
    frames (frame, mesh and camera are replicated) -- and every rank's own particles
    around it.  (Up to round 2's first 8-GPU run every rank drew its own scene: an object 0.7 m away covers
    twice the pixels of one 1.0 m away, and the step time was that of the rank with the nearest object.)"""
    rng = synth.rng_for(CONFIG_ID)
    gt = synth.ground_truth_poses(len(w["V"]), rng)
    frames = [synth.make_frame(render, gt, w["rows"], w["cols"], rng) for _ in range(N_FRAMES)]
    if rank:
        rng = synth.rng_for(CONFIG_ID, stream=rank)
    poses = [synth.particle_poses(gt, w["n"], rng) for _ in range(N_FRAMES)]
    parents = [resampled_parents(w["n"], rng) for _ in range(N_FRAMES)]
    return gt, frames, poses, parents


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region, read through NVML every 5 ms (an
    nvidia-smi process takes longer to start than a timed region lasts); falls back to nvidia-smi."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], threading.Event()
        self.thread = threading.Thread(target=self.run, daemon=True)
        self.nvml = None
        self.sm_max = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            try:
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            return mhz, mask
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True,
                             timeout=5).stdout.strip().split(",")
        self.sm_max = float(out[1])
        mask = sum(bit for (nm, bit), v in zip(self.REASONS, out[2:6]) if v.strip() == "Active")
        return float(out[0]), mask

    def run(self):
        while not self.stop.is_set():
            try:
                self.samples.append(self.sample())
            except Exception:
                pass
            self.stop.wait(0.005 if self.nvml is not None else 0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["clock query unavailable"]}
        sm = sorted(m for m, _ in self.samples)
        mask = 0
        for _, k in self.samples:
            mask |= k
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max,
                "reasons": [nm for nm, bit in self.REASONS if mask & bit], "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


_HOST_THREADS = None


def cpu_oracle_rate(w, frames, poses, sample, reps, threads=None, min_seconds=0.0, max_reps=None):
    """oracle on the host cores: set_observation + loglikes(update) on `sample` particles per step;
    runs `reps` steps, then keeps going until `min_seconds` of CPU work have been timed"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    # eager=0: per-pixel time stamps and propagation on demand, the way dbot's CPU KinectImageModel keeps
    # its occlusion state (2.7x faster on the CPU than propagating whole images as the GPU evaluators do)
    orc = oracle_lib.Oracle(w["rows"], w["cols"], w["K"], w["V"], w["T"], synth.default_parameters(), eager=0)
    global _HOST_THREADS
    if _HOST_THREADS is None:
        # every core this process may run on -- not OpenMP's default, which torchrun pins to 1 by
        # exporting OMP_NUM_THREADS=1 to its workers
        _HOST_THREADS = max(len(os.sched_getaffinity(0)), oracle_lib.oracle_lib().oracle_get_max_threads())
    cores = _HOST_THREADS if threads is None else threads
    oracle_lib.oracle_lib().oracle_set_num_threads(cores)
    idx = np.zeros(sample, dtype=np.int32)
    orc.set_observation(frames[0])
    orc.loglikes_poses(poses[0][:sample], idx, True)  # warm-up, leaves `sample` slots
    times = []
    r = 0
    while r < reps or (sum(times) < min_seconds and (max_reps is None or r < max_reps)):
        f = r % N_FRAMES
        t0 = time.perf_counter()
        orc.set_observation(frames[f])
        orc.loglikes_poses(poses[f][:sample], idx, True)
        times.append(time.perf_counter() - t0)
        r += 1
    orc.close()
    return sample / float(np.mean(times)), cores, float(np.mean(times)), len(times)


def run_reference(args):
    """--impl reference: the reference's CPU evaluator for this path.  Its own implementation is in
    the un-vendored dbot library and cannot be built here, so this arm times the
    oracle port (OpenMP over particles, all host cores) on a bounded particle sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    orc = oracle_lib.Oracle(w["rows"], w["cols"], w["K"], w["V"], w["T"], synth.default_parameters())
    _, frames, poses, _ = make_inputs(w, orc.render, 0)
    orc.close()
    sample = min(w["n"], 250)           # bounded sample of the step's particles: ~0.1 s of CPU work per step
    cpu_oracle_rate(w, frames, poses, sample, max(args.warmup, 1))
    rate, cores, sec, _ = cpu_oracle_rate(w, frames, poses, sample, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": describe(w, args.gpus),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d of the %d particles per step, same frames, OpenMP over particles"
                                       % (sample, w["n"]),
                             "why_port": "the reference's evaluator lives in the un-vendored dbot library "
                                         "(needs Eigen 3.2, fl, catkin): not buildable here, see DESIGN.md"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def single_process_multi_gpu(w, world, n_per_gpu, gather_mode, frames, steps, warmup):
    """BASELINE config 4's shape through ONE handle in ONE process (dbot_b200_config.n_devices, what a C++
    RbSensorBuilder::build() returns): host buffers in, host log-weights out, particles sharded over `world`
    GPUs inside the library, shard-local parents.  -> particle-likelihoods/s (wall clock, median of REPEATS)"""
    n = n_per_gpu * world
    sensor = RbSensorBuilder(ObjectModel(w["V"], w["T"]), CameraData(w["K"], w["rows"], w["cols"]),
                             Parameters(sample_count=n), device_ids=list(range(world)), gather_mode=gather_mode).build()
    rng = synth.rng_for(4)
    gt = synth.ground_truth_poses(len(w["V"]), rng)
    poses = [synth.particle_poses(gt, n, rng) for _ in range(2)]
    parents = [np.concatenate([resampled_parents(n_per_gpu, rng) + g * n_per_gpu for g in range(world)]).astype(np.int32)
               for _ in range(2)]
    idx = np.zeros(n, dtype=np.int32)
    sensor.set_observation(frames[0])
    sensor.loglikes_poses(poses[0], idx, True)

    def step(k):
        idx[:] = parents[k % 2]
        sensor.set_observation(frames[k % len(frames)])
        return sensor.loglikes_poses(poses[k % 2], idx, True)

    for k in range(warmup):
        step(k)
    secs = []
    for _ in range(REPEATS):
        t0 = time.perf_counter()
        for k in range(steps):
            step(k)
        secs.append(time.perf_counter() - t0)
    out = {"value": n * steps / float(np.median(secs)), "unit": UNIT, "particles_total": n, "n_gpus": world,
           "gather": "ncclAllGather" if gather_mode == 0 else "fused peer stores from the kernel epilogue",
           "fps": steps / float(np.median(secs)), "device_bytes": sensor.device_bytes,
           "value_min": n * steps / max(secs), "value_max": n * steps / min(secs)}
    sensor.close()
    return out


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly one JSON line: everything else that writes to fd 1 (NCCL prints its
    version banner there) is sent to stderr, and emit() writes the line to the real stdout"""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (dbot_ros_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w = workload()
    n, rows, cols = w["n"], w["rows"], w["cols"]
    P = rows * cols
    sensor = RbSensorBuilder(ObjectModel(w["V"], w["T"]), CameraData(w["K"], rows, cols),
                             Parameters(sample_count=n), device_id=local_rank).build()
    gt, frames, poses, parents = make_inputs(w, sensor.render, rank)
    dev = torch.device("cuda", local_rank)
    ext = torch.cuda.ExternalStream(sensor.stream, device=dev)
    d_frames = [torch.from_numpy(f).to(dev) for f in frames]
    d_poses = [torch.from_numpy(p).to(dev) for p in poses]
    d_parents = [torch.from_numpy(p).to(dev) for p in parents]
    d_idx = torch.zeros(n, dtype=torch.int32, device=dev)
    d_out = torch.zeros(n, dtype=torch.float64, device=dev)
    d_all = torch.zeros(n * world, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    sharded = None
    if world > 1:
        # the public multi-GPU API under torchrun: particle-sharded sensor, gather of log-weights inside the
        # call -- fused into the particle kernel over NVLink peer memory when the ranks can map each other's
        # memory (csrc/peer_gather.cuh), else one NCCL all-gather
        from dbot_ros_b200.sharded import ShardedRbSensor, shard_bounds

        sharded = ShardedRbSensor(sensor)
    fused = sharded is not None and sharded.gather == "fused"

    def device_step(k):
        f = k % N_FRAMES
        with torch.cuda.stream(ext):
            d_idx.copy_(d_parents[f], non_blocking=True)   # resampled parents of this frame
            sensor.set_observation_device(d_frames[f].data_ptr())
            if fused:      # the kernel's epilogue stores the log-weights into every rank's gathered vector
                sensor.loglikes_gather_device(d_poses[f].data_ptr(), d_idx.data_ptr(), n, True, d_out.data_ptr(), n)
                return
            sensor.loglikes_device(d_poses[f].data_ptr(), d_idx.data_ptr(), n, True, d_out.data_ptr())
            if world > 1:  # the one exchange of the sharded evaluator: all ranks see all weights
                dist.all_gather_into_tensor(d_all, d_out)

    def sync_all():
        sensor.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        sensor.timer_begin()
        for k in range(steps):
            fn(k)
        ms = sensor.timer_end()
        sync_all()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # establish n slots, then warm up
    sensor.set_observation(frames[0])
    sensor.loglikes_poses(poses[0], np.zeros(n, dtype=np.int32), True)
    timed(device_step, args.warmup)
    launches0 = sensor.launch_count
    # clocks are sampled from here to the end of the three timed loops (device-resident, end-to-end,
    # per-kernel); a few tens of milliseconds of GPU work are too short for one loop alone
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    # the K timed steps are repeated until the device has been under load for a quarter of a second (at least
    # REPEATS times, at most 100): K = 20 steps of 0.3 ms alone are 6 ms, too short for clocks and caches to
    # settle and for the driver's sampler to see the GPU busy.  Every repeat times exactly K steps; the median
    # repeat is reported.  (timed() returns the maximum over ranks, so every rank takes the same decision.)
    rep_ms = [timed(device_step, args.steps)]
    n_rep = int(min(100, max(REPEATS, np.ceil(250.0 / max(rep_ms[0], 1e-3)))))
    while len(rep_ms) < n_rep:
        rep_ms.append(timed(device_step, args.steps))
    gpu_launches = int(sensor.launch_count - launches0) // n_rep
    ms = float(np.median(rep_ms))
    value = n * world * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers through the reference-facing call ----
    # headline: ordinary pageable numpy arrays, as the C++ adapter's std::vector (the library stages them);
    # e2e_pinned: the same from page-locked memory (the library copies straight from it)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    host_inputs = {"pageable": ([f.copy() for f in frames], [q.copy() for q in poses]),
                   "pinned": ([pin(f) for f in frames], [pin(q) for q in poses])}
    idx_host = [p.copy() for p in parents]
    if world > 1:
        # Every rank passes the global arrays; it reads only its own contiguous shard of them.
        sharded.n_slots = n * world                   # the warm-up above left n slots on every rank
        lo, hi = shard_bounds(n * world, world, rank)
        g_poses, g_parents = [], []
        for f in range(N_FRAMES):
            gp = np.zeros((n * world,) + poses[f].shape[1:], dtype=np.float32)
            gp[lo:hi] = poses[f]
            gi = np.arange(n * world, dtype=np.int32)
            gi[lo:hi] = parents[f] + lo                # shard-local parents (BASELINE config 4)
            g_poses.append(gp)
            g_parents.append(gi)
        g_idx = [g.copy() for g in g_parents]

    def run_e2e(kind):
        fr, po = host_inputs[kind]

        def host_step(k):
            f = k % N_FRAMES
            if world > 1:
                g_idx[f][:] = g_parents[f]
                sharded.set_observation(fr[f])
                return sharded.loglikes_poses(g_poses[f], g_idx[f], True)
            idx_host[f][:] = parents[f]
            sensor.set_observation(fr[f])
            return sensor.loglikes_poses(po[f], idx_host[f], True)

        for k in range(args.warmup):
            host_step(k)
        secs = []
        for _ in range(REPEATS):
            sync_all()
            t0 = time.perf_counter()
            for k in range(args.steps):
                host_step(k)
            sync_all()
            dt_s = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt_s], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt_s = float(t.item())
            secs.append(dt_s)
        return float(np.median(secs)), secs

    e2e_s, e2e_all = run_e2e("pageable")
    e2e_pin_s, _ = run_e2e("pinned")
    e2e = n * world * args.steps / e2e_s
    h2d = P * 4 + n * len(w["V"]) * 12 * 4 + n * 4
    d2h = n * 8

    # ---- roofline of the dominant kernel: per-launch CUDA events on the library's stream ----
    sensor.profile_enable(True)
    for k in range(args.steps):
        device_step(k)
    sync_all()
    prof = sensor.profile_read()
    sensor.profile_enable(False)
    clocks.__exit__()
    k_ms, k_cnt = prof["update"]
    o_ms, o_cnt = prof["observation"]
    p_ms, p_cnt = prof["plan"]
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ncu = {}
    try:  # one ncu --set full capture of the same kernel and workload, committed under profiles/
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            ncu = json.load(f)
    except Exception:
        pass
    traffic = ncu.get("dram_bytes_per_launch")
    avg_ms = k_ms / max(k_cnt, 1)
    # what this launch has to move: a child that does not adopt its parent's slot reads and writes a whole
    # image (8P); every particle reads the parent's value and writes the posterior at its covered pixels
    copies = float(np.mean([len(p) - len(np.unique(p)) for p in parents]))
    sample = poses[0][:: max(1, n // 16)][:16]
    n_cov = float(np.mean([np.isfinite(sensor.render(q)).sum() for q in sample]))
    # a copy moves the rectangle in which the parent's image can differ from the background (dbot_b200_slot_rect)
    rects = [sensor.slot_rect(sl) for sl in range(0, n, max(1, n // 32))]
    rect_px = float(np.mean([max(0, r1 - r0 + 1) * max(0, c1 - c0 + 1) for r0, r1, c0, c1 in rects]))
    alg_bytes = copies * 8.0 * rect_px + n * 8.0 * n_cov
    dense_bytes = 8.0 * P * n                     # SURVEY.md 8(d), dense formulation: 4P read + 4P written per particle
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    frac_dram = (traffic / (avg_ms * 1e-3) / 1e9 / peak) if traffic else None
    issue_pct = ncu.get("issue_active_pct")
    # the kernel is HBM-bound only if DRAM is what ncu shows busy; otherwise it is bound by instruction issue
    bound = "hbm" if (frac_dram is not None and frac_dram >= 0.5) else "issue"
    roofline = {"bound": bound, "kernel": "dbot_particle_kernel<UPDATE>", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "traffic": traffic, "frac_dram": frac_dram,
                "traffic_source": "profiles/traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, one "
                                  "launch of the same workload)" if traffic else None,
                "issue_active_pct_ncu": issue_pct,
                "algorithmic_bytes_per_launch": alg_bytes,
                "algorithmic_bytes": "copying children (%.0f of %d) x 8 B x %.0f px of the parent's touched rectangle + "
                                     "%d covered pixels x 8 B x %d particles" % (copies, n, rect_px, round(n_cov), n),
                "dense_equivalent": {"bytes_per_launch": dense_bytes, "GBps": dense_bytes / (avg_ms * 1e-3) / 1e9,
                                     "frac": dense_bytes / (avg_ms * 1e-3) / 1e9 / peak,
                                     "note": "8P per particle whether copied or not (SURVEY.md 8d); not a DRAM rate"},
                "avg_launch_ms": avg_ms, "launches_timed": k_cnt,
                "kernel_share_of_step": k_ms / max(k_ms + o_ms + p_ms, 1e-9),
                "observation_kernel_avg_ms": o_ms / max(o_cnt, 1),
                "plan_kernel_avg_ms": p_ms / max(p_cnt, 1)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": describe(w, world), "fps": args.steps / (ms * 1e-3),
            "repeats": {"count": n_rep, "steps_each": args.steps,
                        "ms_per_step": [round(m / args.steps, 6) for m in rep_ms],
                        "value_min": n * world * args.steps / (max(rep_ms) * 1e-3),
                        "value_max": n * world * args.steps / (min(rep_ms) * 1e-3), "reported": "median"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "fps": args.steps / e2e_s, "host_buffers": "pageable",
                    "value_min": n * world * args.steps / max(e2e_all), "value_max": n * world * args.steps / min(e2e_all)},
            "e2e_pinned": {"value": n * world * args.steps / e2e_pin_s, "unit": UNIT, "host_buffers": "page-locked"},
            "gpu_launches": gpu_launches, "clocks": clocks.summary(), "roofline": roofline}
    if world > 1:
        line["gather"] = {"mode": sharded.gather,
                          "what": "log-weights stored by the particle kernel's epilogue into every rank's gathered "
                                  "vector (CUDA IPC peer memory over NVLink) + one warp exchanging arrival counters"
                          if fused else "ncclAllGather through torch.distributed",
                          "fallback_reason": sharded.gather_reason}

    if rank == 0 and world == 1:
        # the same frames through the device-resident filter step (dbot_b200_filter_*): frame in, pose out
        try:
            from dbot_ros_b200.tracker import DeviceParticleTracker, TrackerParameters

            R0, t0 = gt[0]
            from dbot_ros_b200.tracker import rotation_vector

            trk = DeviceParticleTracker(sensor, TrackerParameters(evaluation_count=n), seed=1)
            trk.initialize([np.concatenate([t0, rotation_vector(R0)])])
            for k in range(3):
                trk.track(frames[k % N_FRAMES])
            t1 = time.perf_counter()
            for k in range(args.steps):
                est = trk.track(frames[k % N_FRAMES])
            dt_f = (time.perf_counter() - t1) / args.steps
            line["filter"] = {"frames_per_s": 1.0 / dt_f, "particle_likelihoods_per_s": n / dt_f,
                              "position_error_mm": 1e3 * float(np.linalg.norm(est[0, :3] - t0)),
                              "what": "dbot_b200_filter_track: host depth frame in, integrated pose out, "
                                      "filter state resident in HBM"}
            trk.close()
        except Exception as e:  # the bench line does not depend on it
            line["filter"] = {"error": str(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = min(n, 250)
        rate, cores, sec, reps = cpu_oracle_rate(w, frames, poses, sample, reps=4, min_seconds=12.0, max_reps=400)
        # the reference's own loglikes is a serial loop over the particles: one thread, small sample
        rate1, _, sec1, reps1 = cpu_oracle_rate(w, frames, poses, 32, reps=3, threads=1)
        line["cpu_baseline_single_thread"] = {"value": rate1, "unit": UNIT, "cores": 1, "kind": "port",
                                              "sample": "32 particles per step, %d steps" % reps1}
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d of the %d particles per step, %d steps (%.1f s of CPU work), same "
                                          "frames, OpenMP over particles" % (sample, n, reps, sec * reps)}
    if world > 1 and os.environ.get("DBOT_BENCH_SKIP_SINGLE") != "1":
        # multi-GPU behind the C-ABI (one process, one handle): measured by rank 0 alone while the other ranks
        # wait with idle GPUs.  Weak-scaled config 2 (comparable with this line's value) and, on 8 GPUs or with
        # DBOT_BENCH_CONFIG4=1, BASELINE config 4 at its stated size (8 x 10 000 = 80 000 particles).
        # (the others wait on the rendezvous store, i.e. on the CPU: a NCCL barrier would keep a kernel spinning
        # on their GPUs and the two processes would time-slice them)
        sync_all()
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            sp = {}
            try:
                for name, mode in (("nccl", 0), ("fused", 1)):
                    sp["config2_per_gpu_%s" % name] = single_process_multi_gpu(w, world, n, mode, frames, args.steps,
                                                                               args.warmup)
                if world == 8 or os.environ.get("DBOT_BENCH_CONFIG4") == "1":
                    n4 = synth.CONFIGS[4][1]
                    for name, mode in (("nccl", 0), ("fused", 1)):
                        sp["config4_%s" % name] = single_process_multi_gpu(w, world, n4, mode, frames,
                                                                           max(4, args.steps // 4), 3)
            except Exception as e:  # the bench line does not depend on it
                sp["error"] = str(e)
            line["single_process"] = sp
            store.set("dbot_single_process_done", "1")
        else:
            store.wait(["dbot_single_process_done"])
        dist.barrier()
    if rank == 0:
        emit(line)
    if world > 1:
        sharded.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
