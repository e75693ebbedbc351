"""CPU: the numpy restatement of the reference's ICP (oracle/icp_oracle.py) against outputs frozen from the reference's
own ICP.get_transformation (scipy KDTree + numpy SVD; oracle/gen_golden.py: icp_kats)."""
import numpy as np

from oracle import icp_oracle as io
from tests.util import load_golden


def test_icp_matches_reference():
    g = load_golden("icp_kats.npz")
    assert int(g["n"]) == 5
    for k in range(int(g["n"])):
        r, t, it = io.get_transformation(g["c%d_source" % k], g["c%d_target" % k])
        np.testing.assert_array_equal(r, g["c%d_rotation" % k])
        np.testing.assert_array_equal(t, g["c%d_translation" % k])
        assert 3 < it < 100                                                  # converged by the threshold, not the cap
        r3, t3, it3 = io.get_transformation(g["c%d_source" % k], g["c%d_target" % k], max_iterations=3)
        assert it3 == 3
        np.testing.assert_array_equal(r3, g["c%d_rotation_3it" % k])
        np.testing.assert_array_equal(t3, g["c%d_translation_3it" % k])


def test_best_fit_transform_closed_form_equals_the_svd():
    """what the kernel computes instead of the 2 x 2 SVD (icp.py:76-85), reflections included"""
    from fast_slam_b200.frontend import ICP
    rng = np.random.default_rng(4)
    for k in range(50):
        s = rng.normal(0, 1, (20, 2))
        th = rng.uniform(-3, 3)
        r = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        t = s @ r.T + rng.normal(0, 0.05, (20, 2)) + rng.normal(0, 1, 2)
        if k % 5 == 0:
            t[:, 0] = -t[:, 0]                                                # mirrored target: the det < 0 branch
        r0, t0 = io.best_fit_transform(s, t)
        r1, t1 = ICP.best_fit_transform(s, t)
        assert np.abs(r0 - r1).max() < 1e-12 and np.abs(t0 - t1).max() < 1e-12
