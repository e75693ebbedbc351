"""GPU: map clustering (fs2_known_landmarks / fs2_cluster_points, csrc/fs2_known.cuh) against the reference's frozen
outputs and the numpy restatement.  Cluster count, label order and member counts are exact; centroids agree to the
rounding of the reference's own float64 mean (the device sums exactly in fixed point)."""
import numpy as np
import pytest

from oracle import known_landmarks_oracle as ko
from tests.util import load_golden

pytestmark = pytest.mark.gpu
TAGS = ["clouds", "touching", "lattice", "border", "scatter", "noise", "origin", "ridge", "drive"]
ATOL = 1e-11


def _check(points, eps, ms, cent, mem):
    ref_c, ref_n = ko.cluster_points(points, eps, ms)
    assert len(cent) == len(ref_c)
    np.testing.assert_array_equal(mem, ref_n)
    if len(ref_c):
        np.testing.assert_allclose(np.array(cent).reshape(-1, 2), ref_c, rtol=0, atol=ATOL)


@pytest.mark.parametrize("tag", TAGS)
def test_cluster_points_matches_reference(tag):
    from fast_slam_b200.frontend import GeometryUtils
    k = load_golden("known_landmarks_kats.npz")
    pts, ms = k["%s_pts" % tag], int(k["%s_min_samples" % tag])
    cent, mem = GeometryUtils.cluster_points(pts, 0.5, ms, with_members=True)
    ref = k["%s_cent" % tag]
    assert len(cent) == len(ref)
    if len(ref):
        np.testing.assert_allclose(np.array(cent), ref, rtol=0, atol=ATOL)       # the reference's own centroids
    lab = k["%s_labels" % tag]
    np.testing.assert_array_equal(mem, np.bincount(lab[lab >= 0], minlength=len(ref)))


@pytest.mark.parametrize("tag", TAGS + ["skip"])
def test_known_landmarks_on_device_maps(tag):
    """the same cases as filter state: maps uploaded into a DeviceFilter, clustered where they live"""
    from fast_slam_b200 import DeviceFilter
    k = load_golden("known_landmarks_kats.npz")
    pts, counts = k["%s_pts" % tag], k["%s_counts" % tag]
    P, lcap = len(counts), max(int(counts.max()), 1)
    lm = np.zeros((P, lcap, 6))
    off = 0
    for p, c in enumerate(counts):
        lm[p, :c, 0:2] = pts[off:off + c]
        off += c
    f = DeviceFilter(P, lcap)
    f.upload(count=counts.astype(np.int32), lm=lm)
    res = f.known_landmarks()
    f.close()
    if bool(k["%s_skipped" % tag]):
        assert res is None
        return
    cent, mem, info = res
    assert info["min_samples"] == int(k["%s_min_samples" % tag]) and info["n_points"] == len(pts)
    ref = k["%s_cent" % tag]
    assert len(cent) == len(ref)
    if len(ref):
        np.testing.assert_allclose(cent, ref, rtol=0, atol=ATOL)
    lab = k["%s_labels" % tag]
    np.testing.assert_array_equal(mem, np.bincount(lab[lab >= 0], minlength=len(ref)))
    assert info["noise_points"] == int((lab < 0).sum())


@pytest.mark.parametrize("seed,n,box,eps,ms", [
    (0, 3000, 10.0, 0.5, 12), (1, 3000, 6.0, 0.5, 40), (2, 2000, 20.0, 0.5, 3), (3, 4000, 8.0, 0.3, 15),
    (4, 1500, 12.0, 1.0, 30), (5, 5000, 4.0, 0.5, 300), (6, 2500, 9.0, 0.5, 1), (7, 2500, 9.0, 0.5, 2),
    (8, 6000, 15.0, 0.7, 25), (9, 800, 3.0, 0.05, 2),
])
def test_cluster_points_random_against_oracle(seed, n, box, eps, ms):
    """uniform scatter at densities around the core threshold: the point-level path does most of the work"""
    from fast_slam_b200.frontend import GeometryUtils
    rng = np.random.default_rng(100 + seed)
    pts = rng.uniform(-box / 2, box / 2, size=(n, 2)) + rng.choice([-37.0, 0.0, 1234.5])
    cent, mem = GeometryUtils.cluster_points(pts, eps, ms, with_members=True)
    _check(pts, eps, ms, cent, mem)


@pytest.mark.parametrize("gap", [0.40, 0.47, 0.5, 0.53, 0.62])
def test_two_dense_clouds_near_eps(gap):
    """two clouds whose nearest points are about eps apart: cell-level cores, the link is settled on the points"""
    from fast_slam_b200.frontend import GeometryUtils
    rng = np.random.default_rng(7)
    a = rng.uniform(-0.15, 0.15, size=(1500, 2))
    b = rng.uniform(-0.15, 0.15, size=(1500, 2)) + [0.3 + gap, 0.02]
    pts = np.concatenate([a, b, rng.uniform(-3, 3, size=(40, 2))])[rng.permutation(3040)]
    cent, mem = GeometryUtils.cluster_points(pts, 0.5, 50, with_members=True)
    _check(pts, 0.5, 50, cent, mem)


def test_lattice_exact_distances_with_multiplicity():
    from fast_slam_b200.frontend import GeometryUtils
    rng = np.random.default_rng(3)
    pts = rng.integers(-6, 7, size=(400, 2)) * 0.25            # many points exactly eps (and 0) apart
    for ms in (3, 9, 14, 20):
        cent, mem = GeometryUtils.cluster_points(pts, 0.5, ms, with_members=True)
        _check(pts, 0.5, ms, cent, mem)


def test_update_known_landmarks_api():
    """LandmarkUtils.update_known_landmarks on FastSLAM2.particles (device) and on a list of Particle objects"""
    import contextlib, io
    from fast_slam_2 import FastSLAM2, LandmarkUtils, Landmark, Measurement, Particle, config
    config.NUM_PARTICLES, config.LANDMARK_CAPACITY, config.RNG = 64, 16, "device"
    try:
        f = FastSLAM2()
        obs = [Measurement(2.0, 0.3), Measurement(3.0, -1.0), Measurement(1.5, 2.5)]
        with contextlib.redirect_stdout(io.StringIO()):
            for _ in range(3):
                f.iterate(0.0, 0.0, obs)
        LandmarkUtils.known_landmarks = []
        LandmarkUtils.update_known_landmarks(f.particles)
        dev = np.array([[l.x, l.y] for l in LandmarkUtils.known_landmarks])
        assert all(isinstance(l, Landmark) for l in LandmarkUtils.known_landmarks)
        host_particles = list(f.particles)
        maps = [np.array([[l.x, l.y] for l in p.landmarks]).reshape(-1, 2) for p in host_particles]
        ref = ko.update_known_landmarks(maps)
        assert ref is not None and len(dev) == len(ref[0]) == 3
        np.testing.assert_allclose(dev, ref[0], rtol=0, atol=ATOL)
        LandmarkUtils.known_landmarks = []
        LandmarkUtils.update_known_landmarks([Particle(p.x, p.y, p.yaw, p.weight, list(p.landmarks)) for p in host_particles])
        np.testing.assert_allclose(np.array([[l.x, l.y] for l in LandmarkUtils.known_landmarks]), ref[0], rtol=0, atol=ATOL)
        # fewer landmarks than particles: the reference returns without touching the list
        keep = LandmarkUtils.known_landmarks
        g = FastSLAM2()
        LandmarkUtils.update_known_landmarks(g.particles)
        assert LandmarkUtils.known_landmarks is keep
        f.store.close(); g.store.close()
    finally:
        config.NUM_PARTICLES, config.LANDMARK_CAPACITY, config.RNG = 20, 256, "device"


def test_known_landmarks_at_scale_is_cell_level():
    """65536 particles x 64 landmarks (4.2M points): every cell is core, nothing goes through the point path, and
    the centroids are the per-landmark means"""
    from fast_slam_b200 import DeviceFilter
    from fast_slam_b200.synthetic import fill_synthetic_device
    P, L = 1 << 16, 64
    f = DeviceFilter(P, L + 16)
    fill_synthetic_device(f, L, 1234)
    cent, mem, info = f.known_landmarks()
    st = f.download()
    f.close()
    assert info["n_points"] == P * L and info["min_samples"] == int(L * 0.7)
    assert info["involved_points"] == 0 and info["noise_points"] == 0
    assert len(cent) == L and (mem == P).all()
    ref = st["lm"][:, :L, 0:2].mean(axis=0)                      # cluster k = landmark k (label order = index order)
    np.testing.assert_allclose(cent, ref, rtol=0, atol=1e-9)


def test_cluster_points_errors():
    from fast_slam_b200._lib import Fs2Error
    from fast_slam_b200.frontend import GeometryUtils
    with pytest.raises(Fs2Error):
        GeometryUtils.cluster_points(np.array([[0.0, np.nan], [1.0, 1.0]]), 0.5, 1)
    assert GeometryUtils.cluster_points(np.zeros((0, 2)), 0.5, 1) == []
    one = GeometryUtils.cluster_points(np.array([[3.0, -2.0]]), 0.5, 1)
    assert len(one) == 1 and np.allclose(one[0], [3.0, -2.0], atol=1e-12)


def test_known_landmarks_full_size():
    """BASELINE.json config 3 at full size: 2^20 particles x 256 landmarks = 2.7e8 points.  Every landmark's cloud is
    one cluster of 2^20 points in index order, and its centroid is the mean of that landmark over the particles
    (computed independently with torch on the device)."""
    import torch
    from fast_slam_b200 import DeviceFilter
    from fast_slam_b200.synthetic import fill_synthetic_device
    P, L = 1 << 20, 256
    f = DeviceFilter(P, 320)
    fill_synthetic_device(f, L, 1234)
    cent, mem, info = f.known_landmarks()
    assert info["n_points"] == P * L and info["min_samples"] == int(L * 0.7) and info["noise_points"] == 0
    assert len(cent) == L and (mem == P).all()
    lm = f.lm_raw[:P, :L, 0:2]                       # a fresh filter: particle p owns map slot p
    ref = lm.mean(dim=0).cpu().numpy()
    np.testing.assert_allclose(cent, ref, rtol=0, atol=1e-9)
    # the same maps give the same bits again (integer sums: no dependence on the order of the atomics)
    cent2, mem2, _ = f.known_landmarks()
    assert np.array_equal(cent, cent2) and np.array_equal(mem, mem2)
    f.close()


def test_tile_grid_grows_with_the_map():
    """1600 isolated landmark clouds on a 3 m lattice cover more tiles than the grid starts with (2048, kept under a quarter full):
    the workspace is rebuilt larger, the result does not change"""
    from fast_slam_b200 import DeviceFilter
    P, L = 16, 1600
    rng = np.random.default_rng(8)
    gx, gy = np.meshgrid(np.arange(40) * 3.0 - 60.0, np.arange(40) * 3.0 - 60.0)
    centres = np.stack([gx.ravel(), gy.ravel()], 1)
    lm = np.zeros((P, L, 6))
    lm[:, :, 0:2] = centres[None] + rng.normal(0, 0.02, (P, L, 2))
    f = DeviceFilter(P, L)
    f.upload(count=np.full(P, L, np.int32), lm=lm)
    first = f.known_landmarks(min_samples=10, max_clusters=2048)
    second = f.known_landmarks(min_samples=10, max_clusters=2048)
    f.close()
    for cent, mem, info in (first, second):
        assert len(cent) == L and (mem == P).all() and info["noise_points"] == 0
        np.testing.assert_allclose(cent, lm[:, :, 0:2].mean(axis=0), rtol=0, atol=1e-11)
    assert first[2]["tiles_used"] >= 1600 and second[2]["tiles"] > first[2]["tiles"] >= 2048
    assert np.array_equal(first[0], second[0])


def test_point_level_work_budget(monkeypatch):
    """two dense blobs whose cells are about eps apart: settled by the projection test when they are clearly apart or
    clearly linked; when neither, the pair search is charged against FS2_KL_WORK and the call stops loudly"""
    from fast_slam_b200._lib import Fs2Error
    from fast_slam_b200.frontend import GeometryUtils
    rng = np.random.default_rng(2)
    a = rng.uniform(0.0, 0.03, size=(6000, 2))
    for gap, clusters in ((0.48, 1), (0.56, 2)):
        pts = np.concatenate([a, rng.uniform(0.0, 0.03, size=(6000, 2)) + [gap, 0.0]])
        cent, mem = GeometryUtils.cluster_points(pts, 0.5, 50, with_members=True)
        assert len(cent) == clusters and mem.sum() == 12000
    # no single direction separates these: A = two clumps 3 cm apart in one cell, B = a clump 0.4999 to the right, half
    # way up -- every pair is 0.50012 apart, but along the direction between the cells the projections are 0.4992 apart
    a = np.concatenate([np.zeros((3000, 2)), np.tile([0.0, 0.03], (3000, 1))])
    b = np.tile([0.4999, 0.015], (6000, 1))
    pts = np.concatenate([a, b])
    assert not ko.neighbour_matrix(pts[::50], 0.5)[:120, 120:].any()
    monkeypatch.setenv("FS2_KL_WORK", "1e5")
    with pytest.raises(Fs2Error):
        GeometryUtils.cluster_points(pts, 0.5, 50)
    monkeypatch.setenv("FS2_KL_WORK", "1e9")
    cent, mem = GeometryUtils.cluster_points(pts, 0.5, 50, with_members=True)
    assert len(cent) == 2 and (mem == 6000).all()
