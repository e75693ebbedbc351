"""Randomised comparison of fs2_cluster_points with the numpy restatement (oracle/known_landmarks_oracle.py): mixtures of
blobs, lattices (exact-eps distances, points on cell borders), duplicates and scatter, several eps / min_samples.
    python scripts/kl_stress.py [cases] [seed]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_slam_b200.frontend import GeometryUtils      # noqa: E402
from oracle import known_landmarks_oracle as ko        # noqa: E402


def make(rng):
    parts = []
    kind = rng.integers(0, 5)
    eps = float(rng.choice([0.5, 0.5, 0.5, 0.25, 1.0, 0.37]))
    n = int(rng.integers(50, 1800))
    box = float(rng.uniform(2, 25)) * eps * 2
    off = rng.choice([0.0, -100.0, 1000.25, 3.0e4]) * np.array([1.0, -1.0])
    if kind == 0:                                   # scatter
        parts.append(rng.uniform(-box / 2, box / 2, (n, 2)))
    elif kind == 1:                                 # blobs of different density + scatter
        for _ in range(int(rng.integers(2, 12))):
            c = rng.uniform(-box / 2, box / 2, 2)
            parts.append(c + rng.normal(0, rng.uniform(0.01, 0.6) * eps, (int(rng.integers(3, 200)), 2)))
        parts.append(rng.uniform(-box / 2, box / 2, (n // 4, 2)))
    elif kind == 2:                                 # lattice with multiplicities: exact distances, cell borders
        h = eps / float(rng.choice([1, 2, 4, 16, 32]))
        parts.append(rng.integers(-12, 13, (n, 2)) * h)
    elif kind == 3:                                 # chains: points spaced just under / over eps
        for _ in range(int(rng.integers(1, 6))):
            m = int(rng.integers(5, 120))
            step = eps * float(rng.choice([0.9, 0.999, 1.0, 1.001, 1.1]))
            t = np.arange(m) * step
            a = rng.uniform(0, np.pi)
            parts.append(np.stack([t * np.cos(a), t * np.sin(a)], 1) + rng.uniform(-box / 2, box / 2, 2))
        parts.append(rng.uniform(-box / 2, box / 2, (n // 8, 2)))
    else:                                           # duplicates of a few points + scatter
        base = rng.uniform(-box / 2, box / 2, (int(rng.integers(2, 30)), 2))
        parts.append(base[rng.integers(0, len(base), n)])
        parts.append(rng.uniform(-box / 2, box / 2, (n // 5, 2)))
    pts = np.concatenate(parts) + off
    pts = pts[rng.permutation(len(pts))]
    ms = int(rng.choice([1, 2, 3, 5, 8, 13, 30, 80]))
    return pts, eps, ms, kind


def main(cases=None, seed=None):
    if cases is None:
        cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    if seed is None:
        seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    bad = 0
    t0 = time.time()
    for k in range(cases):
        pts, eps, ms, kind = make(rng)
        cent, mem = GeometryUtils.cluster_points(pts, eps, ms, with_members=True)
        ref_c, ref_n = ko.cluster_points(pts, eps, ms)
        ok = len(cent) == len(ref_c) and np.array_equal(mem, ref_n) and (len(ref_c) == 0 or np.abs(np.array(cent) - ref_c).max() < 1e-9 * max(1.0, np.abs(pts).max()))
        if not ok:
            bad += 1
            print("MISMATCH case %d kind %d n %d eps %g ms %d: clusters %d vs %d" % (k, kind, len(pts), eps, ms, len(cent), len(ref_c)), flush=True)
            if os.environ.get("FS2_STRESS_DUMP"):
                os.makedirs("gpurun_out", exist_ok=True)
                np.save("gpurun_out/kl_bad_%d.npy" % k, pts)
    print("kl_stress: %d cases, %d mismatches, %.1f s" % (cases, bad, time.time() - t0))
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
