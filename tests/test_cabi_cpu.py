"""CPU: the C-ABI library loads without a GPU, exports every symbol include/fs2.h declares, validates its
arguments before touching CUDA, fails loudly (never falls back) when there is no device, and builds
conservative observation cell tables (host logic of the screen)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from fast_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(L):
    hdr = open(os.path.join(ROOT, "include", "fs2.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fs2_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(L, name), "include/fs2.h declares %s but libfs2.so does not export it" % name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)


def test_abi_version_and_error_strings(L):
    assert L.fs2_abi_version() == 1
    assert L.fs2_strerror(0) == b"ok"
    assert b"invalid" in L.fs2_strerror(-1)
    assert b"CUDA" in L.fs2_strerror(-2)


def test_argument_validation_happens_before_cuda(L):
    h = C.c_void_p()
    cfg = _lib.Fs2Config()
    assert L.fs2_create(None, C.byref(h)) == -1
    cfg.num_particles, cfg.landmark_capacity = 0, 16
    assert L.fs2_create(C.byref(cfg), C.byref(h)) == -1
    cfg.num_particles, cfg.landmark_capacity = 8, 0
    assert L.fs2_create(C.byref(cfg), C.byref(h)) == -1
    cfg.num_particles, cfg.landmark_capacity, cfg.global_particles, cfg.global_offset = 8, 4, 8, 4
    assert L.fs2_create(C.byref(cfg), C.byref(h)) == -1            # shard does not fit the global range
    assert L.fs2_destroy(None) == 0
    assert L.fs2_update(None, None, 0, None, None) == -1
    assert L.fs2_launch_count(None) == 0


def test_no_cpu_fallback():
    """Without a CUDA device the product path must refuse, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fast_slam_b200 import DeviceFilter, Fs2Error
    with pytest.raises(Fs2Error):
        DeviceFilter(16, 8)
    from fast_slam_b200.filter import FastSLAM2
    with pytest.raises(Fs2Error):
        FastSLAM2()
    # and the C entry point itself reports the CUDA failure
    L = _lib.load()
    cfg = _lib.Fs2Config()
    cfg.num_particles, cfg.landmark_capacity = 16, 8
    h = C.c_void_p()
    assert L.fs2_create(C.byref(cfg), C.byref(h)) in (-2, -3)
    assert L.fs2_last_cuda_error() != b""


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fast_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if f.endswith(".py"):
                    assert "libfs2_oracle" not in src and "fs2o_" not in src, "%s reaches into the oracle" % f
                else:
                    code = re.sub(r"//.*", "", src)                       # comments may cite the oracle
                    assert "fs2o_" not in code and "dlopen" not in code, "%s reaches into the oracle" % f
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


G1, G2 = 48, 8      # FS2_G1, FS2_G2 (csrc/fs2_update.cuh)


class ObsBatch(C.Structure):
    _fields_ = [("zd", C.c_double * 32), ("za", C.c_double * 32), ("ox", C.c_double * 32), ("oy", C.c_double * 32),
                ("sza", C.c_double * 32), ("cza", C.c_double * 32),
                ("oxf", C.c_float * 32), ("oyf", C.c_float * 32), ("tab1", C.c_uint32 * ((G1 + 2) ** 2)),
                ("tab2", C.c_uint32 * ((G2 + 2) ** 2)),
                ("gx0", C.c_float), ("gy0", C.c_float), ("inv_s1", C.c_float), ("inv_s2", C.c_float), ("e1", C.c_float),
                ("e2", C.c_float), ("amax1", C.c_float), ("amax2", C.c_float), ("xymax", C.c_float),
                ("cx1", C.c_float), ("cy1", C.c_float), ("cx2", C.c_float), ("cy2", C.c_float),
                ("slack", C.c_float), ("M", C.c_int32), ("k0", C.c_int32), ("all_mask", C.c_uint32)]


def _obs_batch(L, M, spread, seed=None):
    rng = np.random.default_rng(M if seed is None else seed)
    pts = rng.uniform(-spread, spread, (M, 2)) + rng.uniform(-3, 3, 2)
    obs = np.stack([np.hypot(pts[:, 0], pts[:, 1]), np.arctan2(pts[:, 1], pts[:, 0])], axis=1).copy()
    ob = ObsBatch()
    assert L.fs2_debug_obs_batch(obs.ctypes.data_as(C.POINTER(C.c_double)), M, C.byref(ob)) == 0
    return rng, pts, ob


def _cell(G, x, y, inv_s, cx, cy):
    """fs2_cell: the device's table index of a point, fp32 fma + clamp + floor."""
    f32 = np.float32
    fx = f32(np.float64(f32(x)) * np.float64(f32(inv_s)) + np.float64(f32(cx)))
    fy = f32(np.float64(f32(y)) * np.float64(f32(inv_s)) + np.float64(f32(cy)))
    fx = min(max(fx, f32(0)), f32(G + 1)); fy = min(max(fy, f32(0)), f32(G + 1))
    return int(np.floor(fy)) * (G + 2) + int(np.floor(fx))


@pytest.mark.parametrize("M,spread", [(32, 12.0), (16, 3.0), (5, 0.5), (1, 1.0), (32, 0.0)])
def test_observation_cell_tables_are_conservative(L, M, spread):
    """For any box no wider than a level's limit, the table entry of the cell that holds the box centre
    contains every observation inside the box (false positives allowed, false negatives never)."""
    assert L.fs2_debug_obs_batch_size() == C.sizeof(ObsBatch)
    rng, pts, ob = _obs_batch(L, M, spread)
    assert ob.M == M and ob.all_mask == (0xFFFFFFFF if M == 32 else (1 << M) - 1)
    ox = np.array(ob.oxf[:M], dtype=np.float32); oy = np.array(ob.oyf[:M], dtype=np.float32)
    np.testing.assert_allclose(ox, pts[:, 0], rtol=1e-6, atol=1e-6)
    assert all(np.isinf(ob.oxf[k]) for k in range(M, 32))
    f32 = np.float32
    # the grids reach past the observations by their own margin: the half-infinite border cells are out of every
    # observation's reach, so a landmark outside the covered square is never a candidate at that level
    for G, tab in ((G1, ob.tab1), (G2, ob.tab2)):
        t = np.array(tab[:]).reshape(G + 2, G + 2)
        assert not t[0].any() and not t[-1].any() and not t[:, 0].any() and not t[:, -1].any()
        assert t.any()
    for level, (G, tab, inv_s, e, cx, cy) in enumerate([(G1, ob.tab1, ob.inv_s1, ob.e1, ob.cx1, ob.cy1),
                                                        (G2, ob.tab2, ob.inv_s2, ob.e2, ob.cx2, ob.cy2)]):
        for _ in range(4000):
            # box centre anywhere around the observations, half widths up to the level's limit
            if rng.uniform() < 0.5:
                k = rng.integers(M)
                c = np.array([ox[k], oy[k]], dtype=np.float64) + rng.uniform(-1.5 * e, 1.5 * e, 2)
            else:
                c = np.array([ob.gx0, ob.gy0]) + rng.uniform(-4 * e, (G + 2) / inv_s, 2)
            rx, ry = rng.uniform(0, e, 2)
            mx, my = f32(c[0]), f32(c[1])
            mask = tab[_cell(G, mx, my, inv_s, cx, cy)]
            inside = (np.abs(ox - mx) < f32(rx)) & (np.abs(oy - my) < f32(ry))
            for k in np.flatnonzero(inside):
                assert mask >> int(k) & 1, "level %d: observation %d inside the box but not in the cell's mask" % (level, k)


@pytest.mark.parametrize("M,spread", [(32, 12.0), (16, 3.0), (3, 40.0)])
def test_level_thresholds_bound_the_box(L, M, spread):
    """The warp-specialised kernel picks a landmark's table level from max(c00, c11) alone (fs2_screen): a safe
    landmark below amaxN and inside xymax must have fs2_box half-widths <= eN, and nothing beyond xymax with a
    level <= 2 covariance can gate an observation."""
    rng, pts, ob = _obs_batch(L, M, spread, seed=100 + M)
    f32 = np.float32
    gate_f = f32(f32(8.0) * f32(1.0000002))
    g1 = f32(gate_f * f32(1.00001))
    omax = max(np.abs(np.array(ob.oxf[:M])).max(), np.abs(np.array(ob.oyf[:M])).max())
    assert ob.xymax > omax + ob.e2
    assert 0 < ob.amax1 < ob.amax2
    assert ob.amax2 > 0.1, "a fresh landmark (0.1 * I, landmark.py:13) must stay in level 2"
    for amax, e in ((ob.amax1, ob.e1), (ob.amax2, ob.e2)):
        for _ in range(2000):
            a = f32(amax) if rng.uniform() < 0.3 else f32(rng.uniform(0, amax))
            x = f32(rng.uniform(-1, 1) * ob.xymax)
            if abs(x) >= ob.xymax:
                continue
            # sqrt.approx: relative error <= 2^-22, taken at its worst
            root = np.float64(np.sqrt(np.float64(a))) * (1 + 2.0 ** -22)
            rx = np.float64(g1) * root * (1 + 2.0 ** -23) + (abs(np.float64(x)) * 2.4e-7 + np.float64(ob.slack)) * (1 + 2.0 ** -22)
            assert rx <= np.float64(e), (a, x, rx, e)
    # beyond xymax: |dx| < gate * sqrt(c00) <= gate * sqrt(amax2) < e2 cannot reach an observation
    assert 8.0 * np.sqrt(np.float64(ob.amax2) * (1 + 1e-7)) * (1 + 2e-6) < ob.e2    # exact test: < 1e-6 relative rounding


@pytest.mark.parametrize("eps", [0.5, 0.1, 0.25, 1.0, 0.3, 2.0 / 3.0, 1e-3, 7.0])
def test_squared_distance_threshold_is_the_exact_image_of_the_sqrt_test(L, eps):
    """The front-end's kernels decide np.sqrt(dx**2 + dy**2) <= eps (geometry_utils.py:26-62 through DBSCAN,
    landmark_utils.py:78-87) as s <= T(eps) on the squared distance s.  T must be the largest double whose correctly
    rounded square root is <= eps: then the two tests agree for every s (sqrt is monotone), checked here on the doubles
    around T and on random s."""
    import math
    L.fs2_frontend_sq_threshold.restype = C.c_double
    L.fs2_frontend_sq_threshold.argtypes = [C.c_double]
    T = L.fs2_frontend_sq_threshold(eps)
    assert math.sqrt(T) <= eps < math.sqrt(math.nextafter(T, math.inf))
    from oracle import frontend_oracle as fe
    assert T == fe.sq_threshold(eps)                     # the restatement used by cluster_labels_two_phase
    s = T
    for _ in range(500):
        assert math.sqrt(s) <= eps
        s = math.nextafter(s, 0.0)
    s = math.nextafter(T, math.inf)
    for _ in range(500):
        assert math.sqrt(s) > eps
        s = math.nextafter(s, math.inf)
    rng = np.random.default_rng(3)
    for s in rng.uniform(0.0, 4.0 * eps * eps, 2000):
        assert (math.sqrt(s) <= eps) == (s <= T)
