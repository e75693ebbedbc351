"""CPU: the numpy front-end restatement (oracle/frontend_oracle.py) against outputs frozen from the reference's
own pipeline (cv2.HoughLines, scipy gaussian_filter1d, sklearn DBSCAN; oracle/gen_golden.py: frontend_kats)."""
import numpy as np
import pytest

from oracle import frontend_oracle as fe
from tests.util import load_golden


@pytest.fixture(scope="module")
def kats():
    return load_golden("frontend_kats.npz")


@pytest.mark.parametrize("tag", ["b180", "b360", "b1081"])
def test_frontend_matches_reference(kats, tag):
    scans = kats["%s_scans" % tag]
    found = 0
    for b, pts in enumerate(scans):
        meas, det = fe.get_measurements(pts, detail=True)
        n = int(kats["%s_nlines" % tag][b])
        # cv2.HoughLines: same lines, same order, bit for bit
        assert len(det["lines"]) == n
        np.testing.assert_array_equal(det["lines"], kats["%s_lines" % tag][b, :n])
        assert det["geometry"][2:] == tuple(kats["%s_geo" % tag][b, :2])
        if "npix" in det:
            assert det["npix"] == kats["%s_geo" % tag][b, 2]
        k = int(kats["%s_k" % tag][b])
        assert len(meas) == k
        np.testing.assert_array_equal(meas, kats["%s_meas" % tag][b, :k])   # same float32 arithmetic: identical
        found += k
    assert found >= 6


def test_line_filter_matches_scipy(kats):
    pts = kats["lf_points"]
    np.testing.assert_array_equal(fe.line_filter(pts, 0.1), pts)             # identity at the default sigma (Q17)
    assert np.abs(fe.line_filter(pts, 1.0) - kats["lf_sigma1"]).max() < 1e-14
    assert np.abs(fe.line_filter(pts, 2.5) - kats["lf_sigma2"]).max() < 1e-14


def test_disc_is_cv2_radius_two_circle():
    assert len(fe.DISC) == 13 and (0, 0) in fe.DISC and (2, 0) in fe.DISC and (1, 1) in fe.DISC and (2, 1) not in fe.DISC


def test_cluster_labels_are_first_appearance_components():
    pts = np.array([[0, 0], [5, 5], [0.3, 0], [5.2, 5.1], [9, 9], [0.6, 0.1]], np.float32)
    np.testing.assert_array_equal(fe.cluster_labels(pts), [0, 1, 0, 1, 2, 0])   # same as sklearn DBSCAN(0.5, 1)


def test_scan_environment_and_measurements_match_reference():
    """laser ranges -> Robot.scan_environment -> get_measurements_to_landmarks, frozen from the reference with HAL
    stubbed by a recorded laser message (oracle/gen_golden.py: frontend_polar_kats)"""
    g = load_golden("frontend_polar_kats.npz")
    assert (g["npts"] < 180).any() and (g["npts"] == 180).any()           # ragged: some scans lose beams
    for b in range(len(g["npts"])):
        p = fe.scan_environment(g["values"][b], g["angles"], float(g["min_range"]), float(g["max_range"]))
        np.testing.assert_array_equal(p, g["pts"][b, :g["npts"][b]])
        m = fe.get_measurements(p)
        assert len(m) == g["k"][b]
        np.testing.assert_array_equal(m, g["meas"][b, :g["k"][b]])


def test_banded_hough_with_halo_equals_the_full_accumulator():
    """The device keeps the Hough accumulator in shared-memory tiles (bands of angle rows x column ranges, one halo row /
    cell either side); the restatement of that route finds cv2.HoughLines' lines bit for bit, also with tiles so narrow
    that every row needs dozens of column ranges."""
    from fast_slam_b200.synthetic import room_scans
    for b, pts in enumerate(room_scans(3, 361, 1.5 * np.pi, seed=99)):
        f = fe.line_filter(pts)
        ox, oy, w, h = fe.image_geometry(f)
        img = fe.rasterise(f, ox, oy, w, h)
        lines, votes = fe.hough_lines(img)
        assert len(lines) >= 4
        for row_words in (2304, 300, 77):
            l2, v2 = fe.hough_lines_banded(img, row_words=row_words)
            np.testing.assert_array_equal(l2, lines)
            np.testing.assert_array_equal(v2, votes)
        l3, v3 = fe.hough_lines_banded(img, band_rows=15, row_words=500)
        np.testing.assert_array_equal(l3, lines)


def test_two_phase_components_equal_dbscan_labels():
    """first-neighbour trees + one merging sweep in any pair order = connected components at eps, labelled by first
    appearance (what sklearn's DBSCAN(eps, min_samples=1) returns), on random point sets from sparse to one blob"""
    rng = np.random.default_rng(1)
    for trial in range(120):
        n = int(rng.integers(1, 100))
        pts = rng.uniform(0, rng.uniform(0.6, 6.0), (n, 2)).astype(np.float32)
        want = fe.cluster_labels(pts)
        np.testing.assert_array_equal(fe.cluster_labels_two_phase(pts), want)
        np.testing.assert_array_equal(fe.cluster_labels_two_phase(pts, order=rng), want)
    chain = np.stack([np.arange(40) * 0.45, np.zeros(40)], 1).astype(np.float32)[rng.permutation(40)]
    np.testing.assert_array_equal(fe.cluster_labels_two_phase(chain, order=rng), np.zeros(40, np.int64))
