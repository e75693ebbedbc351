"""CPU: the numpy restatement of the reference's map clustering (oracle/known_landmarks_oracle.py) against outputs
frozen from the reference's own LandmarkUtils.update_known_landmarks and sklearn's DBSCAN labels
(oracle/gen_golden.py: known_landmark_kats)."""
import numpy as np
import pytest

from oracle import known_landmarks_oracle as ko
from tests.util import load_golden

TAGS = ["clouds", "touching", "lattice", "border", "scatter", "skip", "noise", "origin", "ridge", "drive"]


@pytest.fixture(scope="module")
def kats():
    return load_golden("known_landmarks_kats.npz")


def test_golden_covers_every_case(kats):
    assert list(kats["tags"]) == TAGS


@pytest.mark.parametrize("tag", TAGS)
def test_update_known_landmarks_matches_reference(kats, tag):
    pts, counts = kats["%s_pts" % tag], kats["%s_counts" % tag]
    maps = np.split(pts, np.cumsum(counts)[:-1])
    res = ko.update_known_landmarks(maps)
    if bool(kats["%s_skipped" % tag]):
        assert res is None                                       # landmark_utils.py:133-134
        return
    ms = ko.min_samples_rule(len(pts), len(maps))
    assert ms == int(kats["%s_min_samples" % tag])
    np.testing.assert_array_equal(ko.dbscan_labels(pts, ko.EPS, ms), kats["%s_labels" % tag])   # sklearn's labels
    np.testing.assert_array_equal(res[0], kats["%s_cent" % tag])                                  # same numpy mean
    lab = kats["%s_labels" % tag]
    np.testing.assert_array_equal(res[1], np.bincount(lab[lab >= 0], minlength=len(res[0])))


def test_cases_reach_every_branch(kats):
    """border points, noise, exact-eps distances and the early return all occur in the frozen cases"""
    lab, pts = kats["border_labels"], kats["border_pts"]
    lone = int(np.flatnonzero(pts[:, 0] == 2.5)[0])
    assert lab[lone] == 0 and ko.neighbour_matrix(pts, 0.5)[lone].sum() == 3 < int(kats["border_min_samples"])
    assert (kats["scatter_labels"] < 0).sum() > 100 and kats["scatter_labels"].max() > 10
    assert (kats["noise_labels"] == -1).all() and len(kats["noise_cent"]) == 0
    d2 = ((kats["lattice_pts"][:, None] - kats["lattice_pts"][None]) ** 2).sum(-1)
    assert (d2 == 0.25).any()
