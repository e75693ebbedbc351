#!/usr/bin/env python
"""bench.py -- particle-observation updates/s and ms per filter step (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one FastSLAM2.iterate() equivalent (motion -> 32 sequential observation updates ->
normalise -> Neff -> conditional resample -> estimate) over the synthetic stream of SURVEY.md 8(d):
    N = 1 : config[2] of BASELINE.json, 2^20 particles x 256 landmarks x 32 observations (16 GB of map);
    N > 1 : the same shard per GPU (weak scaling), particles sharded, no data-path collective in the
            update; weight totals all-gathered, global systematic resample.
Prints ONE JSON line (see the keys at the bottom).  `--impl reference` times the CPU restatement of the
reference (oracle/, all host threads) on a bounded sample of the same workload instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 1234
P_PER_GPU = 1 << 20
L, LCAP, M = 256, 320, 32
# BASELINE.json configs[1..3]: particles per GPU, landmarks, capacity, observations per step
WORKLOADS = {"cfg2": (10000, 64, 128, 16), "cfg3": (1 << 20, 256, 320, 32), "cfg4": (1 << 20, 1024, 1088, 32)}
if os.environ.get("FS2_BENCH_L"):          # experiments only: a different initial map size (must be a square)
    L = int(os.environ["FS2_BENCH_L"])
SZ, B_LM = 8, 48


def algorithmic_bytes_update(P, Lm, Mo):
    """pose+weight read and written, count read, map read once, <= M landmark writes."""
    return P * (2 * 4 * SZ + 4) + P * Lm * B_LM + P * Mo * B_LM


def make_synthetic_filter(P, Lm, lcap, seed=SEED, device=None, global_particles=0, global_offset=0):
    from fast_slam_b200 import DeviceFilter
    from fast_slam_b200.synthetic import fill_synthetic_device
    f = DeviceFilter(P, lcap, device=device, seed=seed, global_particles=global_particles, global_offset=global_offset)
    world = fill_synthetic_device(f, Lm, seed)
    return f, world


def synthetic_step_inputs(seed, step, world, Mo, novel=0):
    from fast_slam_b200.synthetic import synthetic_obs, synthetic_odometry
    rot, tr = synthetic_odometry(step)
    return rot, tr, synthetic_obs(seed, step, world, Mo, novel=novel, max_range=12.0)


class ClockSampler:
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md), sampled through NVML every 5 ms
    (a 100 ms nvidia-smi loop sees one sample of a 100 ms region); falls back to one nvidia-smi query."""

    def __init__(self, device=0):
        self.device, self.rows, self.t0, self.stop_flag, self.thread = device, [], 0.0, False, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[device])
                except Exception:
                    idx = device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _loop(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((time.time(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nvml is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def mark(self):
        """Only samples taken after this call count (the timed region)."""
        self.t0 = time.time()

    def stop(self):
        if self.nvml is None:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = [float(v) for v in out.strip().split(",")]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": ["sampled after the run (no NVML)"], "samples": 1}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"], "samples": 0}
        self.t1 = time.time()
        self.stop_flag = True
        self.thread.join(timeout=1)
        n = self.nvml
        rows = [r for r in self.rows if self.t0 <= r[0] <= self.t1] or self.rows[-3:]
        names = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, bit in names.items() if any(r[2] & bit for r in rows))
        try:
            mx = float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM))
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(rows), "power_w_max": max((r[3] for r in rows), default=None),
                "sm_mhz_min": float(min(r[1] for r in rows)) if rows else None}


CPU_SAMPLE_PARTICLES = 1 << 17     # bounded sample of the 2^20-particle workload for the CPU arm (the filter is linear in particles)


def _oracle_with_all_threads():
    """The C restatement, its OpenMP team pinned to every host core this process may use.  Launchers export
    OMP_NUM_THREADS=1 (torchrun does), which would silently turn the "all host threads" baseline into one thread."""
    from oracle import fs2_oracle as fo                    # the one place bench.py may touch oracle/
    try:
        want = len(os.sched_getaffinity(0))
    except AttributeError:
        want = os.cpu_count() or 1
    cores = int(fo.lib().fs2o_set_num_threads(int(want)))
    return fo, cores


def _cpu_port_steps(steps, warmup, sample_particles=CPU_SAMPLE_PARTICLES, min_seconds=0.0, steps_cap=None):
    """`steps` timed filter steps (update + weights, no resample: conservative for the GPU arm) of the CPU restatement on a
    bounded sample of the cfg3 workload; keeps going past `steps` until min_seconds of timed work (up to steps_cap)."""
    fo, cores = _oracle_with_all_threads()
    from fast_slam_b200.synthetic import synthetic_state
    init = synthetic_state(SEED, sample_particles, L, LCAP)
    o = fo.OracleFilter(sample_particles, LCAP, track_pytypes=False)
    o.set_state(init["x"], init["y"], init["yaw"], init["w"], init["count"], lm=init["lm"])
    rng = np.random.default_rng(0)
    per, s = [], 0
    while True:
        rot, tr, obs = synthetic_step_inputs(SEED, s, init["world"], M)
        noise = rng.normal(0, 0.0055, sample_particles)
        t0 = time.perf_counter()
        o.step_noresample(rot, tr, obs, noise)
        dt = time.perf_counter() - t0
        if s >= warmup:
            per.append(dt)
        s += 1
        if len(per) >= steps and (sum(per) >= min_seconds or (steps_cap is not None and len(per) >= steps_cap)):
            break
    t = float(np.sum(per))
    rate = sample_particles * M * len(per) / t
    sample = ("%d particles x %d landmarks x %d observations per step, %d steps (%.1f s) -- a bounded sample of the "
              "2^20-particle workload (the filter is linear in particles); C restatement of the reference, OpenMP over "
              "particles on %d threads" % (sample_particles, L, M, len(per), t, cores))
    return rate, t, len(per), cores, sample


def cpu_port_rate(target_seconds=10.0):
    """`cpu_baseline` of the GPU arm's line: the CPU restatement (oracle/, `kind: port`), all host threads, ~10 s."""
    rate, t, n, cores, sample = _cpu_port_steps(steps=3, warmup=1, min_seconds=target_seconds, steps_cap=400)
    return dict(value=rate, unit="particle-observation updates/s", cores=cores, kind="port", sample=sample,
                reference_python="not on this box: the reference is pure Python and /root/reference does not travel; its own "
                                 "speed at config 1, timed in the build container, is in profiles/r01_reference_python_cfg1.json "
                                 "(214 updates/s)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" is one step of the bounded sample; exactly K steps after W warm-up
    rate, t, n, cores, sample = _cpu_port_steps(steps=args.steps, warmup=args.warmup, steps_cap=args.steps)
    print(json.dumps({
        "impl": "reference", "metric": "particle-observation updates/sec", "value": rate,
        "unit": "particle-observation updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / n, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg3: 2^20 particles x 256 landmarks x 32 observations (CPU arm: %d-particle sample per step)" % CPU_SAMPLE_PARTICLES},
        "cpu_baseline": {"value": rate, "unit": "particle-observation updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "particle-observation updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from fast_slam_b200 import _lib
    global L, LCAP, M
    if args.workload != "cfg3" or args.particles is None:
        wp, wl, wcap, wm = WORKLOADS[args.workload]
        if args.workload != "cfg3":
            L, LCAP, M = wl, wcap, wm
        if args.particles is None:
            args.particles = wp

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    P = args.particles
    Pglobal = P * world_size

    if world_size == 1:
        flt, world = make_synthetic_filter(P, L, LCAP, device=local_rank)
        stepper = None
    else:
        from fast_slam_b200.dist import ShardedFilter
        stepper = ShardedFilter(P, LCAP, seed=SEED)
        flt = stepper.store
        from fast_slam_b200.synthetic import fill_synthetic_device
        world = fill_synthetic_device(flt, L, SEED)

    from fast_slam_b200.filter import _hash_uniform

    copies, deferred = [], []

    def one_step(s, ev=None):
        rot, tr, obs = synthetic_step_inputs(SEED, s, world, M, novel=args.novel)
        u0 = _hash_uniform(SEED, s) / Pglobal
        sigma = 0.001 if rot != 0 else 0.0055
        if stepper is not None:
            return stepper.step(rot, tr, obs, u0, s, events=ev)
        # device-resident inputs, the launches fs2_step_host makes, with events around the update launch and around the
        # weight half (normalise, Neff, resample decided on the device, estimate); ONE read-back + sync per step
        flt.draw_noise(sigma, s)
        if ev is not None:
            ev[0].record()
        flt.motion_update(rot, tr, obs)
        if ev is not None:
            ev[1].record()
        flt.finish_step(u0)
        if ev is not None:
            ev[2].record()
        stats = flt.stats.cpu()                             # 128 B D2H: estimate, Neff, "resampled", maps copied
        copies.append(float(stats[_lib.STAT_COPIES]))
        deferred.append(float(stats[_lib.STAT_DEFERRED]))
        return bool(stats[_lib.STAT_RESAMPLED] != 0)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                 # nvidia-smi needs ~100 ms to start: begin before the warm-up
    for s in range(args.warmup):
        one_step(s)
    barrier()
    launches0 = flt.launches
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    resampled = []
    barrier()
    sampler.mark()
    t_start.record()
    step_wall = []
    for k in range(args.steps):
        t_w = time.perf_counter()
        resampled.append(one_step(args.warmup + k, evs[k]))
        step_wall.append(time.perf_counter() - t_w)
    t_end.record()
    barrier()
    total_ms = t_start.elapsed_time(t_end)
    clocks = sampler.stop() if rank == 0 else None
    launches = flt.launches - launches0
    landmarks_mean_end = float(flt.count.double().mean().item())
    upd_ms = [e[0].elapsed_time(e[1]) for e in evs]
    fin_ms = [e[1].elapsed_time(e[2]) for e in evs] if stepper is None else []
    res_ms = [t for t, r in zip(fin_ms, resampled) if r]
    nores_ms = [t for t, r in zip(fin_ms, resampled) if not r]
    copies_timed = copies[-args.steps:] if stepper is None else []
    deferred_timed = deferred[-args.steps:] if stepper is None else []
    # update launches that follow a resample write the deferred map copies (leaders stream, followers share the screening)
    after_res = [bool(d > 0) for d in deferred[-args.steps - 1:-1]] if stepper is None and len(deferred) > args.steps else []
    if world_size > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        u = torch.tensor([float(np.mean(upd_ms))], device=dev, dtype=torch.float64)
        dist.all_reduce(u, op=dist.ReduceOp.MAX)
        upd_mean = float(u.item())
    else:
        upd_mean = float(np.mean(upd_ms))
    if os.environ.get("FS2_BENCH_VERBOSE") and rank == 0:
        print("update ms per step:", " ".join("%.2f" % v for v in upd_ms), file=sys.stderr)
        print("host wall ms per step:", " ".join("%.2f%s" % (1e3 * v, "*" if r else "") for v, r in zip(step_wall, resampled)), file=sys.stderr)
    ms_per_step = total_ms / args.steps
    value = Pglobal * M / (ms_per_step * 1e-3)

    # ---- end to end through the public API (host arguments, host result), 1 GPU ----
    e2e = None
    known = None
    if world_size == 1:
        from fast_slam_b200 import config as cfg
        from fast_slam_b200.filter import FastSLAM2
        from fast_slam_b200.models import Measurement
        flt.close()
        del flt
        torch.cuda.empty_cache()
        cfg.NUM_PARTICLES, cfg.LANDMARK_CAPACITY, cfg.SEED, cfg.DEVICE, cfg.RNG = P, LCAP, SEED, local_rank, "device"
        api = FastSLAM2()
        from fast_slam_b200.synthetic import fill_synthetic_device
        fill_synthetic_device(api.store, L, SEED)
        import contextlib, io
        steps_in = []
        for s in range(args.warmup + args.steps):
            rot, tr, obs = synthetic_step_inputs(SEED, s, world, M, novel=args.novel)
            steps_in.append((rot, tr, [Measurement(float(d), float(a)) for d, a in obs]))
        with contextlib.redirect_stdout(io.StringIO()):
            for s in range(args.warmup):
                api.iterate(*steps_in[s])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for s in range(args.warmup, args.warmup + args.steps):
                api.iterate(*steps_in[s])              # returns host floats: the step is complete
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        e2e = {"value": P * M * args.steps / dt, "unit": "particle-observation updates/s",
               "ms_per_step": 1e3 * dt / args.steps,
               "h2d_bytes_per_step": M * 16 + 32, "d2h_bytes_per_step": 8 * _lib.FS2_STATS_LEN,
               "note": "FastSLAM2.iterate(rotation, translation, list[Measurement]) -> (x, y, yaw); observations "
                       "travel in the kernel parameter block, motion noise is drawn on the device, the 64-byte "
                       "stats block is read back every step (twice on a resampling step)"}
        # ---- row N1: LandmarkUtils.update_known_landmarks on the maps the run just produced ----
        if not args.no_known:
            try:
                api.store.known_landmarks()                      # first call allocates the workspace
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                reps = 3
                for _ in range(reps):
                    kl = api.store.known_landmarks()
                dtk = (time.perf_counter() - t0) / reps
                info = kl[2]
                known = {"ms_per_call": 1e3 * dtk, "points": int(info["n_points"]), "points_per_s": info["n_points"] / dtk,
                         "min_samples": int(info["min_samples"]), "clusters": int(info["clusters"]),
                         "noise_points": int(info["noise_points"]), "point_level_points": int(info["involved_points"]),
                         "map_bytes_read_per_pass": int(info["n_points"]) * 32,
                         "note": "DBSCAN(eps 0.5, min_samples 0.7 x mean map length) over every landmark of every particle, "
                                 "host call to host centroids; exact (see csrc/fs2_known.cuh)"}
            except Exception as e:                               # capacity limits are reported, not fatal for the bench
                known = {"error": str(e)[:300]}
        api.store.close()

    elif stepper is not None:
        # same metric through the sharded public API: host observations in, host estimate out, every step
        s0 = args.warmup + args.steps
        barrier()
        t0 = time.perf_counter()
        for s in range(s0, s0 + args.steps):
            one_step(s)
            est = stepper.last["estimate"]                 # host floats (combine_stats read the stats blocks back)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        e2e = {"value": Pglobal * M * args.steps / dt, "unit": "particle-observation updates/s",
               "ms_per_step": 1e3 * dt / args.steps,
               "h2d_bytes_per_step": M * 16 + 32, "d2h_bytes_per_step": 8 * _lib.FS2_STATS_LEN,
               "note": "ShardedFilter.step(rotation, translation, observations, u0, step) on every rank with host "
                       "arguments; the global estimate is read back on the host every step; max over ranks, "
                       "steps %d..%d of the same stream" % (s0, s0 + args.steps - 1)}

    # ---- sharded parity, inside the driver's own run: the sharded filter against one GPU holding all particles on a small
    # configuration (same code path: global resample, placement plan, NVLink pulls) -- every decision, estimate, particle
    sharded_parity = None
    if stepper is not None and not args.no_parity:
        from fast_slam_b200.selfcheck import sharded_equals_single
        migrated_bench = list(stepper.migrated_total) if hasattr(stepper, "migrated_total") else None
        mode_bench = stepper.mode
        flt.close()
        del stepper
        torch.cuda.empty_cache()
        try:
            chk = sharded_equals_single(1 << 14, 64, 96, 16, 99, 20)      # P = the synthetic generator's chunk: shards see the single filter's particles
            assert chk["resamples"] >= 2, "the check stream did not resample"
            sharded_parity = {"result": "ok", "particles_per_gpu": 1 << 14, "steps": 20, "resamples": chk["resamples"],
                              "migrated": chk["migrated"], "mode": chk["mode"]}
            chk["sharded"].store.close()
        except AssertionError as e:
            sharded_parity = {"result": "FAILED: %s" % (str(e)[:300],)}
        sharded_parity["bench_mode"] = mode_bench
        if migrated_bench is not None:
            sharded_parity["bench_offspring_moved_total"] = migrated_bench[0]
            sharded_parity["bench_maps_pulled_rank0"] = migrated_bench[1]

    # ---- scan front-end (BASELINE.json config 5): batched 1081-beam scans -> measurements ----
    frontend = None
    if world_size == 1 and not args.no_frontend:
        from fast_slam_b200.frontend import frontend_batch
        from fast_slam_b200.synthetic import room_scans
        pinned = torch.empty((args.frontend_scans, 1081, 2), dtype=torch.float64, pin_memory=True)
        scans = pinned.numpy()                             # the batch lives in pinned host memory (the H2D copy is timed)
        scans[:] = room_scans(args.frontend_scans, 1081, 1.5 * np.pi, seed=99)
        frontend_batch(scans)                              # warm-up: trig tables, device scratch sized for the batch
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            _, kk, stt = frontend_batch(scans)
        dt = (time.perf_counter() - t0) / reps
        frontend_batch(scans, sigma=1.0)
        t0 = time.perf_counter()
        for _ in range(reps):
            _, kk1, _ = frontend_batch(scans, sigma=1.0)
        dt1 = (time.perf_counter() - t0) / reps
        one = scans[:1]
        frontend_batch(one)
        t0 = time.perf_counter()
        for _ in range(50):
            frontend_batch(one)
        dt_one = (time.perf_counter() - t0) / 50
        os.environ["FS2_FE_LEGACY"] = "1"                  # the global-accumulator Hough stage (round 1), for comparison
        frontend_batch(scans)
        t0 = time.perf_counter()
        for _ in range(2):
            _, kk0, _ = frontend_batch(scans)
        dt0 = (time.perf_counter() - t0) / 2
        del os.environ["FS2_FE_LEGACY"]
        frontend = {"scans": int(len(scans)), "beams": 1081, "ms_per_batch": 1e3 * dt, "scans_per_s": len(scans) / dt,
                    "measurements_per_scan": float(np.mean(kk)), "overflow": int((stt != 0).sum()),
                    "single_scan_ms": 1e3 * dt_one,
                    "sigma_1.0": {"ms_per_batch": 1e3 * dt1, "scans_per_s": len(scans) / dt1, "measurements_per_scan": float(np.mean(kk1))},
                    "global_accumulator_path": {"ms_per_batch": 1e3 * dt0, "scans_per_s": len(scans) / dt0,
                                                "same_result": bool(np.array_equal(kk0, kk))},
                    "h2d_bytes_per_batch": int(scans.nbytes), "d2h_bytes_per_batch": int(len(scans) * (64 * 16 + 8)),
                    "note": "pinned host arrays in, host arrays out (H2D + kernels + D2H inside the timed call; device scratch "
                            "is kept between calls), mean of %d calls; sigma 0.1 is the reference's default (identity filter); "
                            "Hough votes and peak search in shared memory (csrc/fs2_frontend.cuh: fe_raster_list, "
                            "fe_vote_peaks) -- bound by shared-memory atomics (180 votes per set pixel), not by HBM" % reps}

    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
        peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    alg = algorithmic_bytes_update(P, L, M)
    achieved = alg / (upd_mean * 1e-3) / 1e9
    # the two forms of the launch: plain, and the one after a resample (leaders stream, followers' copies written): a launch
    # moves P maps either way, so the algorithmic bytes are the same
    by_form = None
    if world_size == 1 and after_res:
        pl = [t for t, a in zip(upd_ms, after_res) if not a]
        af = [t for t, a in zip(upd_ms, after_res) if a]
        by_form = {k: {"launches": len(v), "ms_per_launch": float(np.mean(v)), "ms_min": float(np.min(v)),
                       "frac": alg / (float(np.mean(v)) * 1e-3) / 1e9 / peak, "frac_best": alg / (float(np.min(v)) * 1e-3) / 1e9 / peak}
                   for k, v in (("plain", pl), ("after_resample", af)) if v}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "update_kernel_traffic.json")       # ncu --set full capture of this kernel
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if int(tj.get("particles", 0)) == P and args.workload == "cfg3" and L == 256:
            traffic = float(tj["dram_bytes_read"]) + float(tj["dram_bytes_write"])
    # the resample chain (exact scan + search + copy-on-resample gather): algorithmic bytes of SURVEY.md 8(d)'s B_res
    # with the maps actually copied (first offspring inherit their ancestor's map, the others are copies)
    resample = None
    if res_ms and nores_ms is not None:
        cm = float(np.mean([c for c, r in zip(copies_timed, resampled) if r]))
        dm = float(np.mean([d for d, r in zip(deferred_timed, resampled) if r]))
        lm_mean = landmarks_mean_end
        b_res = 2 * (cm - dm) * lm_mean * B_LM + 2 * P * (4 * SZ + 4) + 3 * P * SZ + 2 * P * 4
        t_res = float(np.mean(res_ms)) - (float(np.mean(nores_ms)) if nores_ms else 0.0)
        upd_after = [t for t, a in zip(upd_ms, after_res) if a]
        upd_plain = [t for t, a in zip(upd_ms, after_res) if not a]
        resample = {"ms": t_res, "extra_offspring_mean": cm, "maps_copied_by_the_resample_mean": cm - dm,
                    "maps_copied_by_the_next_update_mean": dm, "algorithmic_bytes": b_res,
                    "achieved": b_res / (t_res * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": b_res / (t_res * 1e-3) / 1e9 / peak, "launches": 6,
                    "ms_update_after_resample": float(np.mean(upd_after)) if upd_after else None,
                    "ms_update_plain": float(np.mean(upd_plain)) if upd_plain else None,
                    "note": "scan + search + marks + one-pass slot scan + gather (poses; maps of every 8th sibling) + "
                            "commit/estimate; time = weight half of a resampling step minus that of a step without one "
                            "(CUDA events); bytes = 2 x 48 B x landmarks x maps copied here + poses and weights read and "
                            "written + running sums + indices.  The other copies are written by the next update kernel out "
                            "of the shared-memory stage it screens (no second read; a launch moves P maps either way: "
                            "leaders read, followers written -- roofline.algorithmic_bytes_per_launch is unchanged)"}
    # the whole step against the same roofline: update + (on resampling steps) the deep copies of fast_slam_2.py:192-196 as
    # SURVEY.md 8(d) counts them (read + write of every copied map), whoever makes them
    step_roof = None
    if world_size == 1 and copies_timed:
        b_copy = float(np.mean([2 * c * landmarks_mean_end * B_LM if r else 0.0 for c, r in zip(copies_timed, resampled)]))
        b_all = alg + b_copy + 2 * P * SZ
        step_roof = {"algorithmic_bytes_per_step": b_all, "ms_per_step": ms_per_step, "achieved": b_all / (ms_per_step * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": b_all / (ms_per_step * 1e-3) / 1e9 / peak,
                     "note": "update bytes + 2 x 48 B x landmarks x extra offspring on resampling steps (mean over the stream) + the "
                             "weight passes; the copies that ride in the next update kernel are not read a second time, which is why "
                             "this can exceed what separate kernels could reach"}
    cpu = None
    if not args.no_cpu_baseline:                   # rank 0 only (the other ranks have returned), at every N
        cpu = cpu_port_rate(10.0 if world_size == 1 else 4.0)
    out = {
        "metric": "particle-observation updates/sec", "value": value, "unit": "particle-observation updates/s",
        "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "%s per GPU: %d particles x %d landmarks (capacity %d) x %d observations per step; "
                        "state %.3g GB per GPU (%s)" % (args.workload, P, L, LCAP, M, P * LCAP * 48 / 1e9,
                                                         ">> 126 MB L2, no flush needed" if P * LCAP * 48 > 1e9 else
                                                         "L2-resident: launch-latency bound, not a roofline case"),
            "particles_total": Pglobal, "novel_per_step": args.novel,
            "resampled_steps": int(np.sum(resampled)), "ms_update_kernel": upd_mean,
            # device time of the weight half of a step (normalise, Neff, decision, [scan, search, gather], estimate)
            "ms_weights_with_resample": float(np.mean(res_ms)) if res_ms else None,
            "ms_weights_without_resample": float(np.mean(nores_ms)) if nores_ms else None,
            "launches_per_step": float(launches) / args.steps, "host_syncs_per_step": 1,
            # host wall clock per step (each step ends with a 128-byte read-back, so this is the step's latency)
            "ms_step_with_resample": float(1e3 * np.mean([t for t, r in zip(step_wall, resampled) if r])) if any(resampled) else None,
            "ms_step_without_resample": float(1e3 * np.mean([t for t, r in zip(step_wall, resampled) if not r])) if not all(resampled) else None,
            "landmarks_per_particle_at_end": landmarks_mean_end,
        },
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "fs2_update_ws_kernel (fused motion + association + EKF + weights)",
                     "algorithmic_bytes_per_launch": alg, "ms_per_launch": upd_mean, "peak_source": peak_src,
                     "by_form": by_form,
                     "limiter": "instruction supply: ncu gcc__cache_requests_type_instruction at 91-97 % of peak, launch time = "
                                "instruction-cache misses / 7.8e6 per ms in every capture (profiles/r02_update_kernel_ncu_summary.json); "
                                "DRAM traffic is 1.05 x the algorithmic bytes"},
        "step_roofline": step_roof,
        "resample": resample,
        "sharded_parity": (sharded_parity or {}).get("result") if sharded_parity else None,
        "sharded_parity_detail": sharded_parity,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "frontend": frontend,
        "known_landmarks": known,
    }
    print(json.dumps(out))
    if world_size > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=None, help="particles per GPU (default: the workload's)")
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS),
                    help="BASELINE.json config: cfg3 (default, the one the metric is quoted on), cfg2, or one GPU's share of cfg4")
    ap.add_argument("--novel", type=int, default=0, help="observations per step that start a new landmark")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-frontend", action="store_true")
    ap.add_argument("--no-known", action="store_true", help="skip the map-clustering timing (row N1)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the sharded-vs-single equivalence check after the timed region")
    ap.add_argument("--frontend-scans", type=int, default=1024, help="batch of the scan front-end (BASELINE.json config 5: 1024)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
