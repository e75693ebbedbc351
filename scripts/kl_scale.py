"""Times fs2_known_landmarks on 2^20 particles x 256 landmarks (2.7e8 points); FS2_KL_PROFILE=1 prints the stage times."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from fast_slam_b200 import DeviceFilter
from fast_slam_b200.synthetic import fill_synthetic_device
P, L = 1 << 20, 256
f = DeviceFilter(P, 320)
fill_synthetic_device(f, L, 1234)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cent, mem, info = f.known_landmarks()
    print("ms", 1e3 * (time.perf_counter() - t0), info["clusters"], info["involved_points"], flush=True)
