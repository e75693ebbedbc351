"""TEST INFRASTRUCTURE ONLY -- seeded input streams shared by gen_golden.py, the tests and bench.py.

Every stream is a list of (rotation, translation, measurements[(distance, yaw), ...]) tuples, i.e. the
arguments of ``FastSLAM2.iterate`` (reference fast_slam_2.py:33).  Because of quirk Q1 (observations
are gated in the robot frame against world-frame landmarks) the robot stays near the origin.
"""
from __future__ import annotations

import numpy as np

ROOM_WORLD = np.array([[3.0, 1.0], [3.2, 1.1], [-2.0, 2.5], [0.5, -3.0], [4.0, -2.0], [-3.0, -1.0]])


def drive_stream(seed: int, steps: int, world=ROOM_WORLD, mmax: int = 4, sigma: float = 0.03,
                 rot_every: int = 10):
    """9 translation steps of 0.018 m then one 0.05 rad rotation (robot.py:141-149 semantics), with
    0..mmax noisy observations of random world landmarks per step (two of them 0.22 m apart: Q2)."""
    rng = np.random.default_rng(seed)
    out = []
    for s in range(steps):
        rot, tr = (0.05, 0.0) if s % rot_every == rot_every - 1 else (0.0, 0.018)
        k = int(rng.integers(0, mmax + 1))
        sel = rng.choice(len(world), size=min(k, len(world)), replace=False)
        meas = []
        for j in sel:
            wx, wy = world[j]
            meas.append((float(np.hypot(wx, wy) + rng.normal(0, sigma)),
                         float(np.arctan2(wy, wx) + rng.normal(0, sigma))))
        out.append((rot, tr, meas))
    return out


def repeat_stream(seed: int, steps: int, world=ROOM_WORLD, reps: int = 2, sigma: float = 0.02):
    """Every step observes every world landmark ``reps`` times (so several observations of one step
    hit the SAME landmark, and a landmark appended by observation k is matched by k' > k): the
    sequential-dependency case Q7."""
    rng = np.random.default_rng(seed)
    out = []
    for s in range(steps):
        rot, tr = (-0.03, 0.0) if s % 4 == 3 else (0.0, 0.0)
        meas = []
        for _ in range(reps):
            for wx, wy in world:
                meas.append((float(np.hypot(wx, wy) + rng.normal(0, sigma)),
                             float(np.arctan2(wy, wx) + rng.normal(0, sigma))))
        order = rng.permutation(len(meas))
        out.append((rot, tr, [meas[i] for i in order]))
    return out


# the benchmark-shaped generators live with the package (bench.py's GPU arm may not import oracle/)
from fast_slam_b200.synthetic import grid_world, synthetic_state, synthetic_obs, synthetic_odometry  # noqa: E402,F401
