"""Synthetic workloads of SURVEY.md section 8(d) / BASELINE.json configs 2-4 (bench.py and the tests use them).

Host-side generators for small sizes (numpy) and a device-side generator for the full-size state
(torch on the GPU, chunk-keyed so that every GPU count sees identical particles).  Because of quirk Q1
(observations are gated in the robot frame against world-frame landmarks) the true robot pose is the
origin and the odometry is (almost) zero.
"""
from __future__ import annotations

import numpy as np


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
    Returns dict(x,y,yaw,w,count,lm[P][lcap][6], world[L][2])."""
    rng = np.random.default_rng(seed)
    world = grid_world(L, pitch)
    if shuffle:
        world = world[rng.permutation(L)]
    lm = np.zeros((P, lcap, 6))
    lm[:, :L, 0] = world[None, :, 0] + rng.normal(0, 0.02, (P, L))
    lm[:, :L, 1] = world[None, :, 1] + rng.normal(0, 0.02, (P, L))
    lm[:, :L, 2] = rng.uniform(0.002, 0.006, (P, L))
    b = rng.uniform(-0.001, 0.001, (P, L))
    lm[:, :L, 3] = b
    lm[:, :L, 4] = b
    lm[:, :L, 5] = rng.uniform(0.002, 0.006, (P, L))
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
        r_k = np.hypot(wx, wy)
        # bearing noise sigma, but at most 0.1 m lateral: the reference gates with the landmark's own
        # covariance only (quirk Q3, radius 0.36-0.62 m here), so a larger lateral error at long range
        # would turn "observations of known landmarks" into new landmarks
        sb = min(sigma, 0.1 / max(r_k, 1e-9))
        obs[k] = (r_k + rng.normal(0, sigma), np.arctan2(wy, wx) + rng.normal(0, sb))
    return obs


def synthetic_odometry(step: int):
    """zero odometry on 9 of 10 steps (Q12 still draws translation noise), a +-0.001 rad
    rotation on every 10th."""
    if step % 10 == 9:
        return (0.001 if (step // 10) % 2 == 0 else -0.001), 0.0
    return 0.0, 0.0


CHUNK = 1 << 14   # particles per generator key


def fill_synthetic_device(flt, L: int, seed: int, pitch: float = 1.5):
    """Device-side version of synthetic_state for a DeviceFilter shard of any size: particle i of the
    GLOBAL set is generated from the key (seed, i // CHUNK), whatever the number of GPUs.
    Returns world[L][2] (numpy)."""
    import torch
    assert flt.cfg.global_offset % CHUNK == 0 or flt.P < CHUNK
    dev = flt.x.device
    world = grid_world(L, pitch)
    wt = torch.as_tensor(world, device=dev)
    lm = flt.lm_raw
    lm.zero_()
    g = torch.Generator(device=dev)
    goff = int(flt.cfg.global_offset)
    for lo in range(0, flt.P, CHUNK):
        n = min(CHUNK, flt.P - lo)
        g.manual_seed(int(seed) * 1000003 + (goff + lo) // CHUNK)
        r = lambda *shape: torch.rand(*shape, generator=g, device=dev, dtype=torch.float64)    # noqa: E731
        rn = lambda *shape: torch.randn(*shape, generator=g, device=dev, dtype=torch.float64)  # noqa: E731
        blk = lm[lo:lo + n]
        blk[:, :L, 0] = wt[None, :, 0] + 0.02 * rn(n, L)
        blk[:, :L, 1] = wt[None, :, 1] + 0.02 * rn(n, L)
        blk[:, :L, 2] = 0.002 + 0.004 * r(n, L)
        b = -0.001 + 0.002 * r(n, L)
        blk[:, :L, 3] = b
        blk[:, :L, 4] = b
        blk[:, :L, 5] = 0.002 + 0.004 * r(n, L)
        flt.x[lo:lo + n] = 0.05 * rn(n)
        flt.y[lo:lo + n] = 0.05 * rn(n)
        flt.yaw[lo:lo + n] = 0.01 * rn(n)
    flt.w.fill_(1.0 / flt.Pglobal)
    flt.count.fill_(L)
    torch.cuda.synchronize(dev)
    return world


def room_scan(beams: int = 360, fov: float = 2 * np.pi, pose=(0.0, 0.0, 0.0), width: float = 8.0, height: float = 6.0,
              noise: float = 0.01, seed: int = 0):
    """Ray-cast scan of an axis-aligned width x height room centred on the origin from `pose` (x, y, yaw):
    beams x 2 points in the ROBOT frame, range noise N(0, noise^2)."""
    rng = np.random.default_rng(seed)
    px, py, pyaw = pose
    angs = np.linspace(-fov / 2, fov / 2, beams, endpoint=False)
    pts = np.empty((beams, 2))
    for i, a in enumerate(angs):
        dx, dy = np.cos(a + pyaw), np.sin(a + pyaw)
        ts = []
        if dx > 1e-12: ts.append((width / 2 - px) / dx)
        if dx < -1e-12: ts.append((-width / 2 - px) / dx)
        if dy > 1e-12: ts.append((height / 2 - py) / dy)
        if dy < -1e-12: ts.append((-height / 2 - py) / dy)
        t = min(ts) + rng.normal(0, noise)
        pts[i] = (t * np.cos(a), t * np.sin(a))
    return pts


def room_ranges(angles, pose=(0.0, 0.0, 0.0), width: float = 8.0, height: float = 6.0, noise: float = 0.01, seed: int = 0):
    """Laser ranges of the same room for the given beam angles (robot frame), as HAL.getLaserData().values would
    hold them (models/robot.py:38-45): one float per beam."""
    rng = np.random.default_rng(seed)
    px, py, pyaw = pose
    out = np.empty(len(angles))
    for i, a in enumerate(angles):
        dx, dy = np.cos(a + pyaw), np.sin(a + pyaw)
        ts = []
        if dx > 1e-12: ts.append((width / 2 - px) / dx)
        if dx < -1e-12: ts.append((-width / 2 - px) / dx)
        if dy > 1e-12: ts.append((height / 2 - py) / dy)
        if dy < -1e-12: ts.append((-height / 2 - py) / dy)
        out[i] = min(ts) + rng.normal(0, noise)
    return out


def room_scans(batch: int, beams: int, fov: float, seed: int = 99):
    """`batch` scans from poses ~ U over the room interior."""
    rng = np.random.default_rng(seed)
    out = np.empty((batch, beams, 2))
    for b in range(batch):
        pose = (rng.uniform(-3.5, 3.5), rng.uniform(-2.5, 2.5), rng.uniform(-np.pi, np.pi))
        out[b] = room_scan(beams, fov, pose, seed=seed * 1000 + b)
    return out
