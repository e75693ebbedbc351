"""GPU: batched ICP (fs2_icp, csrc/fs2_icp.cuh) against the reference's frozen outputs and the numpy restatement."""
import numpy as np
import pytest

from oracle import icp_oracle as io
from tests.util import load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_icp_matches_reference_outputs():
    from fast_slam_2 import ICP
    g = load_golden("icp_kats.npz")
    for k in range(int(g["n"])):
        r, t = ICP.get_transformation(g["c%d_source" % k], g["c%d_target" % k])
        assert np.abs(r - g["c%d_rotation" % k]).max() < TOL and np.abs(t - g["c%d_translation" % k]).max() < TOL
        r3, t3 = ICP.get_transformation(g["c%d_source" % k], g["c%d_target" % k], max_iterations=3)
        assert np.abs(r3 - g["c%d_rotation_3it" % k]).max() < TOL and np.abs(t3 - g["c%d_translation_3it" % k]).max() < TOL


def test_icp_batch_against_oracle():
    from fast_slam_b200.frontend import ICP
    from fast_slam_b200.synthetic import room_scan
    rng = np.random.default_rng(9)
    src, tgt = [], []
    for b in range(12):
        a = (rng.uniform(-2, 2), rng.uniform(-1.5, 1.5), rng.uniform(-3, 3))
        c = (a[0] + rng.normal(0, 0.04), a[1] + rng.normal(0, 0.04), a[2] + rng.normal(0, 0.03))
        src.append(room_scan(360, 2 * np.pi, a, seed=b))
        tgt.append(room_scan(360, 2 * np.pi, c, seed=100 + b))
    rot, tr, it = ICP.get_transformation_batch(np.stack(src), np.stack(tgt))
    for b in range(12):
        r, t, n = io.get_transformation(src[b], tgt[b])
        assert it[b] == n, (b, it[b], n)
        assert np.abs(rot[b] - r).max() < TOL and np.abs(tr[b] - t).max() < TOL
        assert abs(np.linalg.det(rot[b]) - 1.0) < 1e-12
