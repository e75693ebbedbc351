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


def grid_world(L: int, pitch: float = 1.5):
    """sqrt(L) x sqrt(L) grid centred on the origin, row-major."""
    n = int(round(np.sqrt(L)))
    assert n * n == L, "L must be a square"
    c = (np.arange(n) - (n - 1) / 2.0) * pitch
    gx, gy = np.meshgrid(c, c, indexing="xy")
    return np.stack([gx.ravel(), gy.ravel()], axis=1)


def synthetic_state(seed: int, P: int, L: int, lcap: int, pitch: float = 1.5, shuffle: bool = False):
    """Pre-populated particle set of SURVEY.md 8(d): pose ~ N(0, .05^2), yaw ~ N(0, .01^2), w = 1/P,
    map mean = grid + N(0, .02^2), cov = [[a,b],[b,c]], a,c ~ U(.002,.006), b ~ U(-.001,.001).
    Returns dict(x,y,yaw,w,count,lm[P][6][lcap], world[L][2])."""
    rng = np.random.default_rng(seed)
    world = grid_world(L, pitch)
    if shuffle:
        world = world[rng.permutation(L)]
    lm = np.zeros((P, 6, lcap))
    lm[:, 0, :L] = world[None, :, 0] + rng.normal(0, 0.02, (P, L))
    lm[:, 1, :L] = world[None, :, 1] + rng.normal(0, 0.02, (P, L))
    lm[:, 2, :L] = rng.uniform(0.002, 0.006, (P, L))
    b = rng.uniform(-0.001, 0.001, (P, L))
    lm[:, 3, :L] = b
    lm[:, 4, :L] = b
    lm[:, 5, :L] = rng.uniform(0.002, 0.006, (P, L))
    return dict(x=rng.normal(0, 0.05, P), y=rng.normal(0, 0.05, P), yaw=rng.normal(0, 0.01, P),
                w=np.full(P, 1.0 / P), count=np.full(P, L, dtype=np.int32), lm=lm, world=world)


def synthetic_obs(seed: int, step: int, world, M: int, novel: int = 0, max_range: float = 12.0,
                  sigma: float = 0.0316):
    """M observations of distinct grid landmarks within max_range of the origin (true robot pose is the
    origin), range/bearing noise sigma; the last ``novel`` of them are replaced by points >= 1 m from
    any landmark (append path).  Returns float64 [M][2] (distance, yaw)."""
    rng = np.random.default_rng(seed + step)
    r = np.hypot(world[:, 0], world[:, 1])
    cand = np.flatnonzero((r <= max_range) & (r > 0.2))
    sel = rng.choice(cand, size=min(M, len(cand)), replace=False)
    obs = np.empty((M, 2))
    for k in range(M):
        wx, wy = world[sel[k % len(sel)]]
        if k >= M - novel:
            # cell centre of the grid: >= pitch/sqrt(2) ~ 1.06 m from every landmark for pitch 1.5
            wx, wy = wx + 0.75, wy + 0.75
        obs[k] = (np.hypot(wx, wy) + rng.normal(0, sigma), np.arctan2(wy, wx) + rng.normal(0, sigma))
    return obs


def synthetic_odometry(step: int):
    """zero odometry on 9 of 10 steps (Q12 still draws translation noise), a +-0.001 rad
    rotation on every 10th."""
    if step % 10 == 9:
        return (0.001 if (step // 10) % 2 == 0 else -0.001), 0.0
    return 0.0, 0.0
