"""Front-end at BASELINE.json config 5 (1024 scans x 1081 beams): host-to-host time per batch (also the target of an ncu launch list)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_slam_b200.frontend import frontend_batch
from fast_slam_b200.synthetic import room_scans
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
scans = room_scans(B, 1081, 1.5 * np.pi, seed=99)
frontend_batch(scans)
t0 = time.perf_counter()
for _ in range(3):
    _, k, st = frontend_batch(scans)
dt = (time.perf_counter() - t0) / 3
print("B=%d  %.3f ms per batch  %.0f scans/s  %.2f measurements per scan" % (B, 1e3 * dt, B / dt, float(np.mean(k))))
