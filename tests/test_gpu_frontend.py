"""GPU: the batched CUDA front-end (fs2_frontend) against the numpy restatement and the reference's frozen outputs.
Integer results (which corners are found, their number) exact; (distance, yaw) within float32 rounding: the
reference computes the intersections with numpy's float32 cos/sin, the device with cosf/sinf."""
import numpy as np
import pytest

from oracle import frontend_oracle as fe
from tests.util import load_golden

pytestmark = pytest.mark.gpu
RTOL = 2e-5      # float32 pipeline (north_star: 1e-3 in fp32)


@pytest.mark.parametrize("tag", ["b180", "b360", "b1081"])
def test_frontend_matches_reference_outputs(tag):
    from fast_slam_b200.frontend import frontend_batch
    k = load_golden("frontend_kats.npz")
    scans = k["%s_scans" % tag]
    meas, cnt, status = frontend_batch(scans)
    assert (status == 0).all()
    np.testing.assert_array_equal(cnt, k["%s_k" % tag])
    for b in range(len(scans)):
        ref = k["%s_meas" % tag][b, :cnt[b]]
        np.testing.assert_allclose(meas[b, :cnt[b]], ref, rtol=RTOL, atol=2e-5)


def test_frontend_batch_against_oracle_and_sigma():
    from fast_slam_b200.frontend import frontend_batch
    from fast_slam_b200.synthetic import room_scans
    scans = room_scans(24, 360, 2 * np.pi, seed=5)
    for sigma in (0.1, 1.0):
        meas, cnt, status = frontend_batch(scans, sigma=sigma)
        assert (status == 0).all()
        for b in range(len(scans)):
            ref = fe.get_measurements(scans[b], sigma=sigma)
            assert cnt[b] == len(ref), (sigma, b)
            np.testing.assert_allclose(meas[b, :cnt[b]], ref, rtol=RTOL, atol=2e-5)
    # a scan's result does not depend on its neighbours in the batch
    m1, c1, _ = frontend_batch(scans[5:6])
    m0, c0, _ = frontend_batch(scans)
    assert c1[0] == c0[5] and np.array_equal(m1[0], m0[5])


def test_landmark_utils_api_and_filter_chain():
    """scan -> LandmarkUtils.get_measurements_to_landmarks -> FastSLAM2.iterate, the loop of jde_robots_main.py:28-38."""
    import contextlib, io
    from fast_slam_2 import FastSLAM2, LandmarkUtils, Measurement, config
    from fast_slam_b200.synthetic import room_scan
    config.NUM_PARTICLES, config.LANDMARK_CAPACITY, config.RNG = 256, 32, "device"
    try:
        f = FastSLAM2()
        with contextlib.redirect_stdout(io.StringIO()):
            for s in range(8):
                ms = LandmarkUtils.get_measurements_to_landmarks(room_scan(360, 2 * np.pi, (0.0, 0.0, 0.0), seed=s))
                assert all(isinstance(m, Measurement) for m in ms) and len(ms) == 4      # the four room corners
                x, y, yaw = f.iterate(0.0, 0.0, ms)
        cnt = np.array([len(p.landmarks) for p in f.particles])
        assert (cnt == 4).all() and abs(x) < 0.1 and abs(y) < 0.1
        f.store.close()
    finally:
        config.NUM_PARTICLES, config.LANDMARK_CAPACITY, config.RNG = 20, 256, "device"


def test_frontend_polar_matches_reference_outputs():
    """laser ranges in: the range filter drops beams on the device, scans of one batch keep different lengths"""
    from fast_slam_b200.frontend import LandmarkUtils, frontend_batch, frontend_batch_polar
    g = load_golden("frontend_polar_kats.npz")
    lo, hi = float(g["min_range"]), float(g["max_range"])
    meas, cnt, status = frontend_batch_polar(g["values"], g["angles"], lo, hi)
    assert (status == 0).all()
    np.testing.assert_array_equal(cnt, g["k"])
    for b in range(len(cnt)):
        np.testing.assert_allclose(meas[b, :cnt[b]], g["meas"][b, :cnt[b]], rtol=RTOL, atol=2e-5)
        # the same scan as points, alone: identical result (a scan does not see its neighbours' lengths)
        m1, c1, _ = frontend_batch(g["pts"][b, :g["npts"][b]][None])
        assert c1[0] == cnt[b] and np.array_equal(m1[0], meas[b])
    one = LandmarkUtils.get_measurements_from_laser(g["values"][4], lo, hi)
    assert len(one) == g["k"][4] and np.allclose([[m.distance, m.yaw] for m in one], g["meas"][4, :g["k"][4]], rtol=RTOL, atol=2e-5)


def test_frontend_polar_against_oracle_with_sigma_and_empty_scan():
    from fast_slam_b200.frontend import frontend_batch_polar
    from fast_slam_b200.synthetic import room_ranges
    angles = np.linspace(-0.75 * np.pi, 0.75 * np.pi, 541, endpoint=False)
    rng = np.random.default_rng(12)
    vals = np.stack([room_ranges(angles, (rng.uniform(-3, 3), rng.uniform(-2, 2), rng.uniform(-3, 3)), seed=b) for b in range(10)])
    vals[3] = 50.0                                                   # nothing in range
    for sigma in (0.1, 1.0):
        meas, cnt, status = frontend_batch_polar(vals, angles, 0.5, 7.0, sigma=sigma)
        assert status[3] == 8 and cnt[3] == 0 and (np.delete(status, 3) == 0).all()
        for b in range(len(vals)):
            if b == 3:
                continue
            ref = fe.get_measurements(fe.scan_environment(vals[b], angles, 0.5, 7.0), sigma=sigma)
            assert cnt[b] == len(ref), (sigma, b)
            np.testing.assert_allclose(meas[b, :cnt[b]], ref, rtol=RTOL, atol=2e-5)


def test_line_filter_alone_matches_scipy_outputs():
    from fast_slam_b200.frontend import LineFilter
    k = load_golden("frontend_kats.npz")
    pts = k["lf_points"]
    np.testing.assert_array_equal(LineFilter.filter(pts), pts)                       # sigma 0.1: identity (Q17)
    assert np.abs(LineFilter.filter(pts, sigma=1.0) - k["lf_sigma1"]).max() < 1e-13
    assert np.abs(LineFilter.filter(pts, sigma=2.5) - k["lf_sigma2"]).max() < 1e-13


def test_hough_intersections_match_the_restatement():
    """HoughTransformation.detect_line_intersections: same number of intersections in the same order"""
    from fast_slam_2 import HoughTransformation
    from fast_slam_b200.synthetic import room_scans
    for pts in room_scans(4, 360, 2 * np.pi, seed=21):
        got = np.array(HoughTransformation.detect_line_intersections(pts), dtype=np.float32).reshape(-1, 2)
        _, det = fe.get_measurements(pts, detail=True)
        ref = det.get("inter", np.zeros((0, 2), np.float32))
        assert got.shape == ref.shape and len(ref) > 0
        np.testing.assert_allclose(got, ref, rtol=0, atol=2e-4)       # cosf/sinf vs numpy's float32 cos/sin, / det


def test_fused_hough_equals_the_global_accumulator_path(monkeypatch):
    """The default Hough stage keeps the accumulator in shared memory (pixel list + bands of angle rows, 16-bit cells,
    column ranges with a halo); FS2_FE_LEGACY=1 runs the global (180 + 2) x (numrho + 2) accumulator.  Votes commute, so
    everything downstream must be equal bit for bit -- also for scans whose rho axis needs several column ranges
    (a room four times the size: numrho > 2 x 4606) and for scans too long for the fused path (> 2016 points)."""
    from fast_slam_b200.frontend import frontend_batch
    from fast_slam_b200.synthetic import room_scans
    cases = [room_scans(40, 1081, 1.5 * np.pi, seed=7), 4.0 * room_scans(6, 1081, 1.5 * np.pi, seed=8),
             room_scans(3, 2500, 2 * np.pi, seed=9)]
    for scans in cases:
        monkeypatch.delenv("FS2_FE_LEGACY", raising=False)
        m1, k1, s1 = frontend_batch(scans)
        monkeypatch.setenv("FS2_FE_LEGACY", "1")
        m0, k0, s0 = frontend_batch(scans)
        monkeypatch.delenv("FS2_FE_LEGACY", raising=False)
        assert np.array_equal(k1, k0) and np.array_equal(s1, s0) and np.array_equal(m1, m0)
        assert (s1 == 0).all()
    # the big rooms against the restatement as well (three column ranges per band)
    scans = cases[1][:2]
    meas, cnt, _ = frontend_batch(scans)
    for b in range(len(scans)):
        ref = fe.get_measurements(scans[b])
        assert cnt[b] == len(ref)
        np.testing.assert_allclose(meas[b, :cnt[b]], ref, rtol=RTOL, atol=2e-5)
