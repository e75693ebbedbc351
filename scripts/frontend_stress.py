"""Randomised comparison of the batched CUDA front-end with the numpy restatement (oracle/frontend_oracle.py): rooms of
different size, beam counts, fields of view, range noise, filter widths, range limits that drop beams.
    python scripts/frontend_stress.py [scans] [seed]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_slam_b200.frontend import frontend_batch_polar      # noqa: E402
from fast_slam_b200.synthetic import room_ranges               # noqa: E402
from oracle import frontend_oracle as fe                       # noqa: E402


def main(n=None, seed=None):
    if n is None:
        n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    if seed is None:
        seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    bad = 0
    t0 = time.time()
    done = 0
    while done < n:
        beams = int(rng.choice([180, 270, 360, 541, 720, 1081]))
        fov = float(rng.choice([np.pi, 1.5 * np.pi, 2 * np.pi]))
        angles = np.linspace(-fov / 2, fov / 2, beams, endpoint=False)
        width, height = float(rng.uniform(3, 12)), float(rng.uniform(3, 10))
        sigma = float(rng.choice([0.1, 0.1, 0.5, 1.0, 2.0]))
        lo, hi = 0.2, float(rng.choice([4.0, 6.0, 30.0]))
        B = int(rng.integers(1, 5))
        vals = np.stack([room_ranges(angles, (rng.uniform(-width / 2 + 0.5, width / 2 - 0.5), rng.uniform(-height / 2 + 0.5, height / 2 - 0.5),
                                              rng.uniform(-np.pi, np.pi)), width=width, height=height, noise=float(rng.choice([0.0, 0.01, 0.03])),
                                     seed=int(rng.integers(1 << 30))) for _ in range(B)])
        meas, cnt, status = frontend_batch_polar(vals, angles, lo, hi, sigma=sigma)
        for b in range(B):
            pts = fe.scan_environment(vals[b], angles, lo, hi)
            if len(pts) == 0:
                ok = status[b] == 8 and cnt[b] == 0
            else:
                ref = fe.get_measurements(pts, sigma=sigma)
                ok = (status[b] & ~7) == 0 and ((status[b] & 7) != 0 or (cnt[b] == len(ref) and np.allclose(meas[b, :cnt[b]], ref, rtol=2e-5, atol=2e-5)))
            if not ok:
                bad += 1
                print("MISMATCH beams %d fov %.2f room %.1fx%.1f sigma %.1f hi %.0f: k %d vs %d status %d" % (beams, fov, width, height, sigma, hi, cnt[b], len(ref) if len(pts) else -1, status[b]), flush=True)
            done += 1
    print("frontend_stress: %d scans, %d mismatches, %.1f s" % (done, bad, time.time() - t0))
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
