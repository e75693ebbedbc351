"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's ICP (row N4 of SURVEY.md 8f):

    ICP.get_transformation     fast_slam_2/algorithms/icp.py:13-58
    ICP.best_fit_transform     fast_slam_2/algorithms/icp.py:60-90

The nearest-neighbour query lives in scipy.spatial.KDTree (scipy 1.18.1, not under /root/reference): for every source
point the closest target point in the Euclidean metric; restated as an exhaustive search (lowest index on an exact
tie).  Pinned by executing the reference in the build container: tests/golden/icp_kats.npz (oracle/gen_golden.py).
"""
from __future__ import annotations

import numpy as np


def nearest(source, target):
    d2 = ((source[:, None, :] - target[None, :, :]) ** 2).sum(-1)
    idx = np.argmin(d2, axis=1)
    return np.sqrt(d2[np.arange(len(source)), idx]), idx


def best_fit_transform(source, target):
    """icp.py:60-90"""
    cs, ct = source.mean(axis=0), target.mean(axis=0)
    cov = (source - cs).T @ (target - ct)
    u, _, vt = np.linalg.svd(cov)
    r = vt.T @ u.T
    if np.linalg.det(r) < 0:
        vt[-1, :] *= -1
        r = vt.T @ u.T
    return r, ct - r @ cs


def get_transformation(source, target, max_iterations: int = 100, threshold: float = 1e-5):
    """icp.py:13-58 -> (rotation [2][2], translation [2], iterations run)"""
    source = np.asarray(source, dtype=np.float64).reshape(-1, 2).copy()
    target = np.asarray(target, dtype=np.float64).reshape(-1, 2)
    prev = float("inf")
    rot, tr = np.eye(2), np.zeros(2)
    it = 0
    for _ in range(max_iterations):
        dist, idx = nearest(source, target)
        r, t = best_fit_transform(source, target[idx])
        source = source @ r.T + t
        rot, tr = r @ rot, r @ tr + t
        it += 1
        m = dist.mean()
        if abs(prev - m) < threshold:
            break
        prev = m
    return rot, tr, it
