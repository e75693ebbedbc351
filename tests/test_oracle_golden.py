"""CPU: the C oracle (oracle/fs2_oracle.c) against the golden vectors frozen from the unmodified
reference (oracle/gen_golden.py).  This is what pins the oracle (prompt section 3)."""
import numpy as np
import pytest

from oracle import fs2_oracle as fo
from tests.util import load_golden, max_rel, replay_trajectory


@pytest.fixture(scope="module")
def kats():
    return load_golden("stage_kats.npz")


def test_mahalanobis_kat(kats):
    # geometry_utils.py:14-23; closed-form inverse vs LAPACK: a few ulp
    n = len(kats["maha_d"])
    d = np.array([fo.mahalanobis(kats["maha_a"][i], kats["maha_b"][i], kats["maha_cov"][i])[0] for i in range(n)])
    assert np.isnan(kats["maha_d"][5]) and np.isnan(d[5])        # indefinite covariance -> NaN -> no match
    assert max_rel(kats["maha_d"], d) < 1e-12
    # gate decisions identical
    assert ((kats["maha_d"] < 8) == (d < 8)).all()


def test_associate_kat(kats):
    # landmark_utils.py:92-117: FIRST match in list order (Q2)
    lm, cnt, obs, idx = kats["assoc_lm"], kats["assoc_count"], kats["assoc_obs"], kats["assoc_idx"]
    L = lm.shape[1]
    got = []
    for c in range(len(idx)):
        blk = np.ascontiguousarray(lm[c])                        # [L][6]
        got.append(fo.lib().fs2o_associate(obs[c, 0], obs[c, 1], fo._dp(blk), int(cnt[c]), L, 8.0))
    np.testing.assert_array_equal(np.array(got), idx)
    assert (idx >= 0).sum() > 20 and (idx < 0).sum() > 5


def test_mvn_pdf_kat(kats):
    # scipy.stats.multivariate_normal.pdf at fast_slam_2.py:156 (lower triangle of Q, Q6)
    got = np.array([fo.mvn_pdf2(kats["pdf_nu"][i], kats["pdf_Q"][i]) for i in range(len(kats["pdf_val"]))])
    assert max_rel(kats["pdf_val"], got) < 1e-11


def test_mvn_pdf_rejects_like_scipy():
    assert fo.mvn_pdf2([0.0, 0.0], [[1.0, 0.0], [1.0, 1.0]]) is None       # singular (lower triangle)
    assert fo.mvn_pdf2([0.0, 0.0], [[1.0, 0.0], [2.0, 1.0]]) is None       # indefinite
    assert fo.mvn_pdf2([0.0, 0.0], [[np.nan, 0.0], [0.0, 1.0]]) is None    # non-finite
    assert fo.mvn_pdf2([0.0, 0.0], [[1.0, 0.0], [0.0, 1.0]]) == pytest.approx(1 / (2 * np.pi), rel=1e-15)


def test_motion_kat(kats):
    # fast_slam_2.py:69-87 (Q12).  sin/cos: numpy's vs glibc's, <= 1 ulp each.
    P = len(kats["motion_x0"])
    for c in range(len(kats["motion_rot"])):
        f = fo.OracleFilter(P, 1)
        f.x[:] = kats["motion_x0"]; f.y[:] = kats["motion_y0"]; f.yaw[:] = kats["motion_yaw0"]
        f.motion(kats["motion_rot"][c], kats["motion_tr"][c], kats["motion_noise"][c])
        np.testing.assert_array_equal(f.yaw, kats["motion_yaw"][c])          # + - fmod only: bit exact
        assert max_rel(kats["motion_x"][c], f.x) < 1e-14
        assert max_rel(kats["motion_y"][c], f.y) < 1e-14
        if kats["motion_rot"][c] != 0:                                        # translation dropped
            np.testing.assert_array_equal(f.x, kats["motion_x0"])


def test_normalize_and_neff_kat(kats):
    # fast_slam_2.py:161-175 and :212-223, including CPython's builtin-sum behaviour (Q18): bit exact
    for c in range(len(kats["norm_in"])):
        P = kats["norm_in"].shape[1]
        f = fo.OracleFilter(P, 1)
        f.w[:] = kats["norm_in"][c]
        f.wkind[:] = kats["norm_kind"][c]
        f.normalize()
        np.testing.assert_array_equal(f.w, kats["norm_out"][c], err_msg="case %d" % c)
        assert f.neff() == kats["neff_out"][c]


@pytest.mark.parametrize("tag", ["n1", "n2", "n7", "n64", "n1000", "n4096"])
def test_resample_kat(kats, tag):
    # fast_slam_2.py:177-199 (Q10): indices bit exact
    W, U, I = kats["res_%s_w" % tag], kats["res_%s_u0" % tag], kats["res_%s_idx" % tag]
    for c in range(len(U)):
        idx, stuck = fo.resample_indices(W[c], U[c])
        assert not stuck
        np.testing.assert_array_equal(idx, I[c], err_msg="%s case %d" % (tag, c))
        # Q11 after the resample: first arg-max over the copied weights
        assert fo.lib().fs2o_argmax(len(idx), fo._dp(np.ascontiguousarray(W[c][idx]))) == kats["res_%s_argmax_after" % tag][c]


def test_argmax_first_on_ties(kats):
    w = np.ascontiguousarray(kats["argmax_w"])
    assert fo.lib().fs2o_argmax(len(w), fo._dp(w)) == int(kats["argmax_first"]) == 1


@pytest.mark.parametrize("name,P,lcap", [("traj_drive.npz", 24, 64), ("traj_repeat.npz", 16, 48)])
def test_trajectory_from_origin(name, P, lcap):
    g = load_golden(name)
    f = fo.OracleFilter(P, lcap)
    worst, nres = replay_trajectory(g, f, lcap, rtol=1e-9)
    assert nres >= 1
    assert (g["assoc"] >= 0).sum() > 50 and (g["assoc"] == -1).sum() > 10


def test_trajectory_synthetic_state():
    g = load_golden("traj_synth.npz")
    f = fo.OracleFilter(12, 48)
    f.set_state(g["init_x"], g["init_y"], g["init_yaw"], g["init_w"], g["init_counts"], lm_p_l_6=g["init_lm"])
    f.wkind[:] = 1          # rh.set_state stores Python floats
    replay_trajectory(g, f, 48, rtol=1e-9)
