#!/usr/bin/env python
"""Update-kernel micro-benchmark for build variants (FS2_LIB=<variant .so> python scripts/bench_update.py).

Config 3 state (2^20 particles x L landmarks, capacity 320, 32 observations), the bench stream, CUDA events around
each fused update launch.  Prints one JSON line: mean / min ms per launch over the timed steps, split into the
steps before and after the maps have grown past a chunk boundary, and the fraction of the measured HBM peak.
A cross-check of the weights against a second library (FS2_CHECK_LIB) can be asked for with --check.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=1 << 20)
    ap.add_argument("--landmarks", type=int, default=256)
    ap.add_argument("--lcap", type=int, default=320)
    ap.add_argument("--obs", type=int, default=32)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--novel", type=int, default=0)
    ap.add_argument("--tag", default=os.environ.get("FS2_LIB", "default"))
    a = ap.parse_args()
    import torch
    from fast_slam_b200 import DeviceFilter, _lib
    from fast_slam_b200.synthetic import fill_synthetic_device, synthetic_obs, synthetic_odometry
    P, L, M = a.particles, a.landmarks, a.obs
    f = DeviceFilter(P, a.lcap, device=0, seed=1234)
    world = fill_synthetic_device(f, L, 1234)
    ms, cnts = [], []
    for s in range(a.warmup + a.steps):
        rot, tr = synthetic_odometry(s)
        obs = synthetic_obs(1234, s, world, M, novel=a.novel, max_range=12.0)
        f.draw_noise(0.001 if rot != 0 else 0.0055, s)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f.motion_update(rot, tr, obs)
        e1.record()
        # keep the weights in range without resampling (the update alone is timed here)
        f.weight_total()
        f.normalize()
        torch.cuda.synchronize()
        if s >= a.warmup:
            ms.append(e0.elapsed_time(e1))
            cnts.append(float(f.count.double().mean().item()))
    ms = np.array(ms)
    alg = P * (2 * 4 * 8 + 4) + P * L * 48 + P * M * 48
    peak = 6543.7
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    out = {"tag": a.tag, "particles": P, "landmarks": L, "obs": M, "novel": a.novel,
           "ms_mean": float(ms.mean()), "ms_min": float(ms.min()), "ms_median": float(np.median(ms)), "ms_max": float(ms.max()),
           "ms_first4": float(ms[:4].mean()), "ms_last4": float(ms[-4:].mean()),
           "landmarks_end": cnts[-1], "frac_mean": alg / (ms.mean() * 1e-3) / 1e9 / peak,
           "frac_first4": alg / (ms[:4].mean() * 1e-3) / 1e9 / peak,
           "status_or": int(torch.bitwise_or(f.status, 0).max().item()),
           "w_checksum": float(f.w.sum().item()), "count_sum": int(f.count.sum().item()),
           "x_checksum": float(f.x.abs().sum().item()),
           "ms_per_step": [round(float(v), 3) for v in ms], "landmarks_per_step": [round(v, 2) for v in cnts]}
    print(json.dumps(out))
    f.close()


if __name__ == "__main__":
    main()
