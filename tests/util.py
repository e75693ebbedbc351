"""Helpers shared by the CPU (oracle) and GPU parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance for fp64 state: 1e-5 relative.  The tests use a much tighter bound where the
# arithmetic allows it, and say so at the call site.
RTOL_FP64 = 1e-5


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def max_rel(a, b, floor=1e-12):
    """max |a-b| / max(|a|, floor) over finite entries; the NaN patterns must be identical."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    fa, fb = np.isfinite(a), np.isfinite(b)
    assert (fa == fb).all(), "finite/NaN pattern differs"
    if not fa.any():
        return 0.0
    return float(np.max(np.abs(a[fa] - b[fa]) / np.maximum(np.abs(a[fa]), floor)))


def replay_trajectory(g, flt, lcap, check_state_every=1, rtol=1e-9):
    """Drive ``flt`` (OracleFilter-like: .step(rot,tr,obs,noise,u0) -> dict, .x/.y/.yaw/.w/.count,
    .lm_p_l_6(L)) with the recorded inputs of golden trajectory ``g`` and compare every step.
    Indices (association, resampling) must be identical; state within rtol (relative)."""
    S = len(g["rotation"])
    worst = 0.0
    nres = 0
    for s in range(S):
        M = int(g["nmeas"][s])
        obs = g["meas"][s, :M]
        out = flt.step(float(g["rotation"][s]), float(g["translation"][s]), obs, g["noise"][s], float(g["u0"][s]))
        assert bool(out["resampled"]) == bool(g["resampled"][s]), "resample decision differs at step %d" % s
        nres += int(out["resampled"])
        if M:
            np.testing.assert_array_equal(np.asarray(out["assoc"]), g["assoc"][s, :M], err_msg="association, step %d" % s)
        np.testing.assert_array_equal(np.asarray(out["resample_idx"]), g["resample_idx"][s], err_msg="resample idx, step %d" % s)
        if s % check_state_every == 0 or s == S - 1:
            st = flt.host_state() if hasattr(flt, "host_state") else dict(
                x=flt.x, y=flt.y, yaw=flt.yaw, w=flt.w, counts=flt.count, lm=flt.lm_p_l_6(lcap))
            np.testing.assert_array_equal(st["counts"], g["counts"][s], err_msg="landmark counts, step %d" % s)
            for k in ("x", "y", "yaw", "w"):
                worst = max(worst, max_rel(g[k][s], st[k]))
            worst = max(worst, max_rel(g["lm"][s][:, :lcap], st["lm"], floor=1e-9))
            worst = max(worst, max_rel(g["estimate"][s], out["estimate"]))
            assert worst <= rtol, "state differs at step %d: %g" % (s, worst)
    return worst, nres
