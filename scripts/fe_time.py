"""Front-end at BASELINE.json config 5 (1024 scans x 1081 beams): host-to-host time per batch and the latency of a
single scan, for the default Hough stage (votes and peaks in shared memory) and the global-accumulator one
(FS2_FE_LEGACY=1), with a bit-for-bit comparison of their results.  `quick` as second argument: three default-path
calls only (the target of an ncu launch list)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_slam_b200.frontend import frontend_batch
from fast_slam_b200.synthetic import room_scans

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
scans = torch.empty((B, 1081, 2), dtype=torch.float64, pin_memory=True).numpy()
scans[:] = room_scans(B, 1081, 1.5 * np.pi, seed=99)


def timed(x, reps, **kw):
    frontend_batch(x, **kw)
    t0 = time.perf_counter()
    for _ in range(reps):
        out = frontend_batch(x, **kw)
    return (time.perf_counter() - t0) / reps, out


if quick:
    dt, (m, k, st) = timed(scans, 3)
    print("B=%d  %.3f ms per batch  %.0f scans/s  %.2f measurements per scan" % (B, 1e3 * dt, B / dt, float(np.mean(k))))
    sys.exit(0)

res = {"scans": B, "beams": 1081}
dt, (m1, k1, s1) = timed(scans, 5)
res["fused"] = {"ms_per_batch": 1e3 * dt, "scans_per_s": B / dt}
dt, _ = timed(scans, 5, sigma=1.0)
res["fused_sigma1"] = {"ms_per_batch": 1e3 * dt, "scans_per_s": B / dt}
dt, _ = timed(scans[:1], 50)
res["fused"]["single_scan_ms"] = 1e3 * dt
os.environ["FS2_FE_LEGACY"] = "1"
dt, (m0, k0, s0) = timed(scans, 3)
res["global_accumulator"] = {"ms_per_batch": 1e3 * dt, "scans_per_s": B / dt}
dt, _ = timed(scans[:1], 50)
res["global_accumulator"]["single_scan_ms"] = 1e3 * dt
del os.environ["FS2_FE_LEGACY"]
res["same_result"] = bool(np.array_equal(m1, m0) and np.array_equal(k1, k0) and np.array_equal(s1, s0))
res["measurements_per_scan"] = float(np.mean(k1))
res["status_nonzero"] = int((s1 != 0).sum())
print(json.dumps(res))
