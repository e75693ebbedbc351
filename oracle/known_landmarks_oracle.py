"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's map clustering (row N1 of SURVEY.md 8f):

    LandmarkUtils.update_known_landmarks      fast_slam_2/utils/landmark_utils.py:120-144
      GeometryUtils.cluster_points            fast_slam_2/utils/geometry_utils.py:26-62

The arithmetic lives in sklearn.cluster.DBSCAN (scikit-learn 1.9.0; not under /root/reference).  Its published
algorithm, restated here without the traversal:

  * neighbourhood of a point = every point (itself included) with  dx*dx + dy*dy <= eps*eps  in float64 (the
    KD-tree's reduced distance; sklearn/neighbors/_binary_tree.pxi.tp);
  * core point  = neighbourhood size >= min_samples;
  * clusters    = connected components of the core points under that neighbourhood relation, numbered by their
    lowest core point index (dbscan_inner walks the points in index order and finishes one cluster before it
    starts the next);
  * a non-core point next to core points takes the LOWEST-numbered cluster among them (the first to reach it);
    with no core neighbour it is noise (-1);
  * the reference then averages every cluster's points (numpy mean over rows) in ascending label order.

Pinned by executing the reference in the build container: tests/golden/known_landmarks_kats.npz
(oracle/gen_golden.py).  Labels and member counts must be bit exact; centroids agree to summation rounding.
Quadratic in the number of points -- meant for the few thousand points of a test, not for a benchmark.
"""
from __future__ import annotations

import numpy as np

EPS = 0.5                 # landmark_utils.py:137
MIN_SAMPLES_FRAC = 0.7    # landmark_utils.py:130


def neighbour_matrix(pts: np.ndarray, eps: float) -> np.ndarray:
    p = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    dx = p[:, None, 0] - p[None, :, 0]
    dy = p[:, None, 1] - p[None, :, 1]
    return (dx * dx + dy * dy) <= eps * eps


def dbscan_labels(pts, eps: float, min_samples: int) -> np.ndarray:
    """DBSCAN(eps, min_samples).fit(pts).labels_"""
    p = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    n = len(p)
    lab = -np.ones(n, np.int64)
    if n == 0:
        return lab
    nb = neighbour_matrix(p, eps)
    core = nb.sum(1) >= min_samples
    comp = -np.ones(n, np.int64)              # component id of core points, in order of their lowest index
    k = 0
    for i in range(n):
        if not core[i] or comp[i] >= 0:
            continue
        comp[i] = k
        stack = [i]
        while stack:
            a = stack.pop()
            for b in np.flatnonzero(nb[a] & core & (comp < 0)):
                comp[b] = k
                stack.append(int(b))
        k += 1
    lab[core] = comp[core]
    for i in np.flatnonzero(~core):
        c = comp[nb[i] & core]
        if len(c):
            lab[i] = c.min()
    return lab


def cluster_points(pts, eps: float, min_samples: int):
    """GeometryUtils.cluster_points -> (centroids [K][2], member counts [K]) in ascending label order."""
    p = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    lab = dbscan_labels(p, eps, min_samples)
    k = int(lab.max()) + 1 if len(lab) else 0
    cent = np.zeros((k, 2))
    cnt = np.zeros(k, np.int64)
    for l in range(k):
        sel = p[lab == l]
        cent[l] = sel.mean(axis=0)
        cnt[l] = len(sel)
    return cent, cnt


def min_samples_rule(n_points: int, n_particles: int) -> int:
    """landmark_utils.py:129-130: int(0.7 * average number of landmarks per particle)."""
    return int((n_points / n_particles) * MIN_SAMPLES_FRAC)


def update_known_landmarks(maps):
    """maps: one [count_p][2] array per particle.  Returns (centroids, counts), or None when the reference
    returns without touching known_landmarks (min_samples < 1, landmark_utils.py:133-134)."""
    pts = np.concatenate([np.asarray(m, dtype=np.float64).reshape(-1, 2) for m in maps]) if len(maps) else np.zeros((0, 2))
    ms = min_samples_rule(len(pts), len(maps))
    if ms < 1:
        return None
    return cluster_points(pts, EPS, ms)
