"""TEST INFRASTRUCTURE ONLY -- freezes outputs of the UNMODIFIED reference into tests/golden/*.npz.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden
The reference has no tests or golden vectors of its own, so these files are the
pin: every array below is produced by executing cy-rae/fast-slam's own code (HAL stubbed, NUM_THREAD=1)
with numpy 2.3.5 / scipy 1.18.1 as installed here.  The .npz files are committed; this script is the
recipe.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from oracle import scenarios as sc    # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def record_trajectory(P: int, stream, np_seed: int, lcap: int, init=None):
    np.random.seed(np_seed)
    flt = rh.new_filter(P)
    if init is not None:
        rh.set_state(flt, init["x"], init["y"], init["yaw"], init["w"], init["count"], init["lm"])
    S = len(stream)
    mmax = max(1, max(len(m) for _, _, m in stream))
    rec = dict(
        rotation=np.zeros(S), translation=np.zeros(S), nmeas=np.zeros(S, np.int32),
        meas=np.full((S, mmax, 2), np.nan), noise=np.zeros((S, P)), u0=np.full(S, np.nan),
        resampled=np.zeros(S, np.bool_), assoc=np.full((S, mmax, P), -9, np.int32),
        resample_idx=np.zeros((S, P), np.int32), estimate=np.zeros((S, 3)),
        x=np.zeros((S, P)), y=np.zeros((S, P)), yaw=np.zeros((S, P)), w=np.zeros((S, P)),
        counts=np.zeros((S, P), np.int32), lm=np.full((S, P, lcap, 6), np.nan),
    )
    if init is not None:
        st0 = rh.get_state(flt, lcap)
        for k in ("x", "y", "yaw", "w", "counts", "lm"):
            rec["init_" + k] = st0[k]
    for s, (rot, tr, meas) in enumerate(stream):
        r = rh.iterate_recorded(flt, rot, tr, meas)
        st = rh.get_state(flt, lcap)
        assert st["counts"].max() <= lcap
        rec["rotation"][s], rec["translation"][s], rec["nmeas"][s] = rot, tr, len(meas)
        if meas:
            rec["meas"][s, :len(meas)] = np.array(meas)
            rec["assoc"][s, :len(meas)] = r["assoc"]
        rec["noise"][s] = r["noise"]
        rec["u0"][s] = r["u0"]
        rec["resampled"][s] = r["resampled"]
        rec["resample_idx"][s] = r["resample_idx"]
        rec["estimate"][s] = r["estimate"]
        for k in ("x", "y", "yaw", "w", "counts", "lm"):
            rec[k][s] = st[k]
    return rec


def stage_kats():
    ref = rh.load_reference()
    rng = np.random.default_rng(2024)
    out = {}
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())  # noqa: E731

    # geometry_utils.py:14-23
    n = 300
    a = rng.normal(0, 3, (n, 2)); b = a + rng.normal(0, 0.4, (n, 2))
    cov = np.zeros((n, 2, 2))
    for i in range(n):
        s0, s1 = rng.uniform(0.002, 0.12, 2)
        off = rng.uniform(-0.9, 0.9) * np.sqrt(s0 * s1)
        cov[i] = [[s0, off], [off + rng.uniform(-1e-7, 1e-7), s1]]
    cov[:5] = np.array([[0.1, 0.0], [0.0, 0.1]])                      # fresh landmark (landmark.py:13)
    cov[5] = [[0.01, 0.02], [0.02, 0.01]]                             # indefinite -> NaN distance
    a[5] = (1.0, 2.0); b[5] = (1.3, 2.0)
    with np.errstate(all="ignore"):
        d = np.array([ref.GeometryUtils.mahalanobis_distance(a[i], b[i], cov[i]) for i in range(n)])
    out.update(maha_a=a, maha_b=b, maha_cov=cov, maha_d=d)

    # landmark_utils.py:92-117 (first match, Q2)
    cases = 120
    L = 12
    lm = np.zeros((cases, L, 6)); cnt = rng.integers(0, L + 1, cases).astype(np.int32)
    obs = np.zeros((cases, 2)); idx = np.zeros(cases, np.int32)
    for c in range(cases):
        pts = rng.uniform(-4, 4, (L, 2))
        if c % 3 == 0 and L > 3:
            pts[3] = pts[1] + rng.normal(0, 0.05, 2)                  # two landmarks inside one gate
        s = rng.uniform(0.001, 0.1, (L, 2))
        o = rng.uniform(-0.3, 0.3, L) * np.sqrt(s[:, 0] * s[:, 1])
        lm[c, :, 0:2] = pts; lm[c, :, 2] = s[:, 0]; lm[c, :, 3] = o; lm[c, :, 4] = o; lm[c, :, 5] = s[:, 1]
        tgt = rng.integers(0, L)
        obs[c] = pts[tgt] + rng.normal(0, 0.3, 2) * (1 if c % 4 else 8)
        lms = [ref.Landmark(lm[c, j, 0], lm[c, j, 1], lm[c, j, 2:6].reshape(2, 2).copy()) for j in range(cnt[c])]
        _, i = ref.LandmarkUtils.associate_landmarks(ref.Landmark(obs[c, 0], obs[c, 1]), lms)
        idx[c] = -1 if i is None else i
    out.update(assoc_lm=lm, assoc_count=cnt, assoc_obs=obs, assoc_idx=idx)

    # scipy.stats.multivariate_normal.pdf as called at fast_slam_2.py:156
    from scipy.stats import multivariate_normal
    n = 200
    Q = np.zeros((n, 2, 2)); nu = rng.normal(0, 0.05, (n, 2)); pdf = np.zeros(n)
    for i in range(n):
        s0, s1 = rng.uniform(0.001, 0.05, 2)
        off = rng.uniform(-0.8, 0.8) * np.sqrt(s0 * s1)
        Q[i] = [[s0, off + rng.uniform(-1e-9, 1e-9)], [off, s1]]
        pdf[i] = multivariate_normal.pdf(nu[i], mean=np.zeros(2), cov=Q[i])
    out.update(pdf_Q=Q, pdf_nu=nu, pdf_val=pdf)

    # fast_slam_2.py:69-87 (Q12): rotation != 0 drops the translation; yaw wrap
    P = 64
    flt = rh.new_filter(P)
    yaw0 = rng.uniform(-np.pi, np.pi, P); yaw0[:4] = [np.pi - 1e-4, -np.pi + 1e-4, 3.1415, -3.1415]
    x0 = rng.normal(0, 2, P); y0 = rng.normal(0, 2, P)
    mot = []
    for rot, tr, sigma in [(0.0, 0.018, 0.0055), (0.05, 0.3, 0.001), (0.0, 0.0, 0.0055), (-0.4, 0.0, 0.001),
                           (0.0, -0.25, 0.0055)]:
        noise = rng.normal(0, sigma, P); noise[:4] = [3e-4, -3e-4, 2e-4, -2e-4]
        for i, p in enumerate(flt.particles):
            p.x, p.y, p.yaw = float(x0[i]), float(y0[i]), float(yaw0[i])
        it = iter(noise)
        orig = np.random.normal
        np.random.normal = lambda *a, **k: np.float64(next(it))
        try:
            for i in range(P):
                flt._FastSLAM2__move_particle(i, rot, tr)
        finally:
            np.random.normal = orig
        mot.append(dict(rot=rot, tr=tr, noise=noise,
                        x=np.array([float(p.x) for p in flt.particles]),
                        y=np.array([float(p.y) for p in flt.particles]),
                        yaw=np.array([float(p.yaw) for p in flt.particles])))
    out.update(motion_x0=x0, motion_y0=y0, motion_yaw0=yaw0,
               motion_rot=np.array([m["rot"] for m in mot]), motion_tr=np.array([m["tr"] for m in mot]),
               motion_noise=np.stack([m["noise"] for m in mot]), motion_x=np.stack([m["x"] for m in mot]),
               motion_y=np.stack([m["y"] for m in mot]), motion_yaw=np.stack([m["yaw"] for m in mot]))

    # fast_slam_2.py:161-175, :212-223 (Q8, Q9, Q18); weights given as np.float64 and as Python floats
    norm_in, norm_out, norm_kind, neff_out = [], [], [], []
    P = 97
    flt = rh.new_filter(P)
    base = rng.uniform(0, 1, P) ** 6
    for case in range(8):
        w = base * rng.uniform(0.5, 1.5, P)
        kind = np.zeros(P, np.uint8)
        if case == 1:
            w = w * 1e-9                                      # total < 1e-5 -> reset to 1/N
        if case == 2:
            w[::3] = 1e-7                                     # below 1e-5: left un-normalised
        if case == 3:
            kind[:] = 1                                       # all exact Python floats -> compensated sum
        if case == 4:
            kind[:40] = 1                                     # compensated prefix then plain additions
        if case == 5:
            w = np.full(P, 1.0 / P); kind[:] = 1
        if case == 6:
            w = w / w.sum()
        if case == 7:
            w[:] = 0.0; w[11] = 0.7; w[50] = 0.3
        for i, p in enumerate(flt.particles):
            p.weight = float(w[i]) if kind[i] else np.float64(w[i])
        flt._FastSLAM2__normalize_weights()
        norm_in.append(w.copy()); norm_kind.append(kind)
        norm_out.append(np.array([float(p.weight) for p in flt.particles]))
        with quiet():
            neff_out.append(float(flt._FastSLAM2__calculate_effective_particles()))
    out.update(norm_in=np.stack(norm_in), norm_kind=np.stack(norm_kind), norm_out=np.stack(norm_out),
               neff_out=np.array(neff_out))

    # fast_slam_2.py:177-199 (Q10) resampling indices, and :201-210 (Q11) first arg-max
    res = {}
    for tag, P in [("n1", 1), ("n2", 2), ("n7", 7), ("n64", 64), ("n1000", 1000), ("n4096", 4096)]:
        flt = rh.new_filter(P)
        ws, u0s, idxs, ams = [], [], [], []
        for case in range(6):
            if case == 0:
                w = rng.uniform(0, 1, P)
            elif case == 1:
                w = rng.uniform(0, 1, P) ** 12
            elif case == 2:
                w = np.full(P, 1.0)
            elif case == 3:
                w = rng.uniform(0, 1, P); w[rng.uniform(0, 1, P) < 0.7] = 0.0; w[-1] = 0.5
            elif case == 4:
                w = np.full(P, 1e-12); w[P // 2] = 1.0
            else:
                w = np.exp(rng.normal(0, 4, P))
            w = w / w.sum()
            u0 = float(rng.uniform(0, 1.0 / P))
            if case == 2:
                u0 = 0.0
            for i, p in enumerate(flt.particles):
                p.weight = np.float64(w[i]); p._fs2_tag = i
            orig = np.random.uniform
            np.random.uniform = lambda *a, **k: u0
            try:
                flt._FastSLAM2__low_variance_resample()
            finally:
                np.random.uniform = orig
            idx = np.array([p._fs2_tag for p in flt.particles], np.int32)
            am = max(range(P), key=lambda i: flt.particles[i].weight)
            est = flt._FastSLAM2__estimate_robot_position()
            assert est[0] == flt.particles[am].x
            ws.append(w); u0s.append(u0); idxs.append(idx); ams.append(am)
            flt = rh.new_filter(P)
        res["res_%s_w" % tag] = np.stack(ws); res["res_%s_u0" % tag] = np.array(u0s)
        res["res_%s_idx" % tag] = np.stack(idxs); res["res_%s_argmax_after" % tag] = np.array(ams, np.int32)
    out.update(res)
    # arg-max with ties: first occurrence
    w = np.array([0.1, 0.3, 0.3, 0.05, 0.3]); flt = rh.new_filter(5)
    for i, p in enumerate(flt.particles):
        p.weight = float(w[i]); p.x = float(i)
    out.update(argmax_w=w, argmax_first=np.array(flt._FastSLAM2__estimate_robot_position()[0]))
    return out


def frontend_kats():
    """Outputs of the reference's own front-end (landmark_utils.py:20-37, cv2 4.13 / scipy 1.18 / sklearn 1.9)."""
    import cv2
    from fast_slam_b200.synthetic import room_scans
    ref = rh.load_reference()
    out = {}
    for tag, beams, fov in (("b180", 180, np.pi), ("b360", 360, 2 * np.pi), ("b1081", 1081, 1.5 * np.pi)):
        scans = room_scans(6, beams, fov, seed=7 + beams)
        K = 16
        meas = np.full((len(scans), K, 2), np.nan); k = np.zeros(len(scans), np.int32)
        nlines = np.zeros(len(scans), np.int32); lines = np.full((len(scans), 128, 2), np.nan, np.float32)
        geo = np.zeros((len(scans), 4), np.int32)
        for b, pts in enumerate(scans):
            m = ref.LandmarkUtils.get_measurements_to_landmarks(pts)
            k[b] = len(m)
            for j, mm in enumerate(m):
                meas[b, j] = (mm.distance, mm.yaw)
            fil = ref.LineFilter.filter(pts)
            img, w, h = ref.HoughTransformation._HoughTransformation__create_hough_transformation_image(fil)
            L = cv2.HoughLines(img, 1, np.pi / 180, 80)
            if L is not None:
                nlines[b] = len(L); lines[b, :len(L)] = L.reshape(-1, 2)
            geo[b] = (w, h, int((img > 0).sum()), 0)
        out.update({"%s_scans" % tag: scans, "%s_meas" % tag: meas, "%s_k" % tag: k, "%s_nlines" % tag: nlines,
                    "%s_lines" % tag: lines, "%s_geo" % tag: geo})
    pts = room_scans(1, 360, 2 * np.pi, seed=3)[0]
    out["lf_points"] = pts
    out["lf_sigma1"] = ref.LineFilter.filter(pts, sigma=1.0)
    out["lf_sigma2"] = ref.LineFilter.filter(pts, sigma=2.5)
    return out


def frontend_polar_kats():
    """Robot.scan_environment (models/robot.py:32-58, HAL stubbed with a recorded laser message) followed by the
    reference's get_measurements_to_landmarks: 180-beam laser, range limits that drop beams."""
    import importlib
    from fast_slam_b200.synthetic import room_ranges
    ref = rh.load_reference()
    hal = sys.modules["HAL"]
    sys.path.insert(0, rh.REFERENCE_ROOT)
    try:
        robot_mod = importlib.import_module("fast_slam_2.models.robot")
    finally:
        sys.path.remove(rh.REFERENCE_ROOT)
    angles = np.radians(np.arange(180) - 90)
    rng = np.random.default_rng(41)
    B, K = 8, 16
    values = np.zeros((B, 180)); npts = np.zeros(B, np.int32); pts = np.full((B, 180, 2), np.nan)
    meas = np.full((B, K, 2), np.nan); k = np.zeros(B, np.int32)
    min_range, max_range = 0.3, 6.0
    for b in range(B):
        pose = (rng.uniform(-2.5, 2.5), rng.uniform(-1.5, 1.5), rng.uniform(-np.pi, np.pi))
        values[b] = room_ranges(angles, pose, seed=500 + b)
        msg = types.SimpleNamespace(values=[float(v) for v in values[b]], minRange=min_range, maxRange=max_range, timeStamp=0)
        hal.getLaserData = lambda msg=msg: msg
        p = robot_mod.Robot.scan_environment()
        npts[b] = len(p); pts[b, :len(p)] = p
        m = ref.LandmarkUtils.get_measurements_to_landmarks(p)
        k[b] = len(m)
        for j, mm in enumerate(m):
            meas[b, j] = (mm.distance, mm.yaw)
    return dict(values=values, angles=angles, min_range=np.array(min_range), max_range=np.array(max_range), npts=npts,
                pts=pts, meas=meas, k=k)


def icp_cases():
    """pairs of 2-D point sets: the same room seen from two nearby poses, and a rotated / shifted random cloud"""
    from fast_slam_b200.synthetic import room_scan
    rng = np.random.default_rng(77)
    cases = []
    for k, beams in enumerate((180, 360, 360, 540)):
        a = (rng.uniform(-2, 2), rng.uniform(-1.5, 1.5), rng.uniform(-3, 3))
        b = (a[0] + rng.normal(0, 0.05), a[1] + rng.normal(0, 0.05), a[2] + rng.normal(0, 0.04))
        cases.append((room_scan(beams, 2 * np.pi, a, seed=10 + k), room_scan(beams, 2 * np.pi, b, seed=20 + k)))
    pts = rng.uniform(-3, 3, size=(300, 2))
    th = 0.12
    r = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    cases.append((pts, pts[rng.permutation(300)[:250]] @ r.T + [0.2, -0.1] + rng.normal(0, 0.005, (250, 2))))
    return cases


def icp_kats():
    """Outputs of the reference's own ICP.get_transformation (icp.py:13-58, scipy KDTree + numpy SVD)."""
    import importlib
    rh.load_reference()
    sys.path.insert(0, rh.REFERENCE_ROOT)
    try:
        icp_mod = importlib.import_module("fast_slam_2.algorithms.icp")
    finally:
        sys.path.remove(rh.REFERENCE_ROOT)
    out = {}
    cases = icp_cases()
    for k, (src, tgt) in enumerate(cases):
        r, t = icp_mod.ICP.get_transformation(src.copy(), tgt.copy())
        r1, t1 = icp_mod.ICP.get_transformation(src.copy(), tgt.copy(), max_iterations=3)
        out["c%d_source" % k], out["c%d_target" % k] = src, tgt
        out["c%d_rotation" % k], out["c%d_translation" % k] = r, t
        out["c%d_rotation_3it" % k], out["c%d_translation_3it" % k] = r1, t1
    out["n"] = np.array(len(cases))
    return out


def known_landmark_cases():
    """Per-particle landmark maps for the map-clustering KATs (row N1): list of (tag, [array [count_p][2]] * P)."""
    rng = np.random.default_rng(2024)
    cases = []
    # A: 20 particles x 6 landmark clouds (sigma 3 cm), strays 0.3-0.7 m off, maps of different length
    centres = np.array([[2.0, 1.0], [-1.5, 2.5], [0.0, -3.0], [4.0, 4.0], [-4.0, -1.0], [1.0, 5.5]])
    maps = []
    for p in range(20):
        m = [c + rng.normal(0, 0.03, 2) for c in centres[:rng.integers(4, 7)]]
        for _ in range(rng.integers(0, 3)):
            c = centres[rng.integers(0, 6)]
            a = rng.uniform(0, 2 * np.pi)
            m.append(c + rng.uniform(0.3, 0.7) * np.array([np.cos(a), np.sin(a)]))
        maps.append(np.array(m))
    cases.append(("clouds", maps))
    # B: two clouds 0.55 m apart (they chain into one cluster or not, point by point) and one far away
    maps = []
    for p in range(30):
        maps.append(np.array([[0, 0] + rng.normal(0, 0.05, 2), [0.55, 0] + rng.normal(0, 0.05, 2),
                              [5, 5] + rng.normal(0, 0.05, 2)]))
    cases.append(("touching", maps))
    # C: a 0.25 m lattice -- squared distances are exact multiples of 1/16, the eps boundary is hit exactly
    maps = []
    for p in range(8):
        ij = rng.integers(0, 9, size=(12, 2))
        maps.append(ij * 0.25)
    cases.append(("lattice", maps))
    # D: a lone point exactly eps from one point of each of two tight clusters: 3 neighbours < min_samples = 4,
    #    so it is a border point of both and takes the lower label (cluster "B", listed first)
    maps = []
    far = np.array([[10.0, 0.0], [0.0, 10.0], [-10.0, 0.0], [0.0, -10.0]])
    for p in range(10):
        maps.append(np.vstack([[2.0 - 0.001 * p, 0.0], [3.0 + 0.001 * p, 0.0], far + rng.normal(0, 0.01, far.shape)]))
    maps[4] = np.vstack([maps[4], [[2.5, 0.0]]])
    cases.append(("border", maps))
    # E: 2500 points scattered over 12 m x 12 m, 100 "particles": core / border / noise all occur
    pts = rng.uniform(-6, 6, size=(2500, 2))
    cases.append(("scatter", [pts[25 * p:25 * (p + 1)] for p in range(100)]))
    # F: fewer landmarks than particles -> min_samples = 0 -> the reference returns without clustering
    maps = [np.zeros((0, 2)) for _ in range(10)]
    maps[3] = np.array([[1.0, 1.0]]); maps[7] = np.array([[1.02, 1.0], [4.0, 4.0]])
    cases.append(("skip", maps))
    # G: nothing dense enough: every point is noise
    cases.append(("noise", [rng.uniform(-50, 50, size=(4, 2)) for _ in range(6)]))
    # H: negative coordinates, clouds straddling the origin and cell borders (multiples of 1/32 m)
    centres = np.array([[0.0, 0.0], [-0.5, -0.5], [-3.03125, 2.0], [1.0, -0.96875]])
    cases.append(("origin", [centres + rng.normal(0, 0.04, centres.shape) for _ in range(25)]))
    # I: a dense ridge 6 m long: one cluster through chaining, ends far beyond eps of each other
    maps = []
    for p in range(40):
        t = rng.uniform(0, 6, size=20)
        maps.append(np.stack([t, 0.3 * np.sin(t)], 1) + rng.normal(0, 0.02, (20, 2)))
    cases.append(("ridge", maps))
    return cases


def known_landmark_kats():
    """Outputs of the reference's own LandmarkUtils.update_known_landmarks (landmark_utils.py:120-144, sklearn 1.9)
    for the cases above, plus DBSCAN's labels for the same points (the reference only keeps the centroids)."""
    from sklearn.cluster import DBSCAN
    ref = rh.load_reference()
    Particle = ref.particle_mod.Particle
    out = {}
    cases = known_landmark_cases()
    # J: the map of a real run: final state of the drive trajectory (reference filter, 24 particles)
    tr = record_trajectory(24, sc.drive_stream(1, 60), np_seed=7, lcap=64)
    cnt = tr["counts"][-1]
    cases.append(("drive", [tr["lm"][-1][p, :cnt[p], 0:2] for p in range(len(cnt))]))
    for tag, maps in cases:
        parts = []
        for m in maps:
            q = Particle(0.0, 0.0, 0.0)
            q.landmarks = [ref.Landmark(float(x), float(y)) for x, y in np.asarray(m).reshape(-1, 2)]
            parts.append(q)
        sentinel = [ref.Landmark(123.0, 456.0)]
        ref.LandmarkUtils.known_landmarks = sentinel
        ref.LandmarkUtils.update_known_landmarks(parts)
        kl = ref.LandmarkUtils.known_landmarks
        skipped = kl is sentinel
        pts = np.concatenate([np.asarray(m, dtype=np.float64).reshape(-1, 2) for m in maps])
        ms = int((len(pts) / len(maps)) * 0.7)
        out["%s_pts" % tag] = pts
        out["%s_counts" % tag] = np.array([len(m) for m in maps], np.int32)
        out["%s_skipped" % tag] = np.array(skipped)
        out["%s_min_samples" % tag] = np.array(ms)
        out["%s_cent" % tag] = np.zeros((0, 2)) if skipped else np.array([[l.x, l.y] for l in kl]).reshape(-1, 2)
        out["%s_labels" % tag] = (DBSCAN(eps=0.5, min_samples=ms).fit(pts).labels_.astype(np.int64)
                                  if ms >= 1 else -np.ones(len(pts), np.int64))
    out["tags"] = np.array([t for t, _ in cases])
    return out


def main():
    if not rh.reference_available():
        raise SystemExit("needs the reference tree at %s" % rh.REFERENCE_ROOT)
    os.makedirs(GOLDEN, exist_ok=True)
    if sys.argv[1:] == ["icp"]:             # only the ICP KATs
        np.savez_compressed(os.path.join(GOLDEN, "icp_kats.npz"), **icp_kats())
        return
    if sys.argv[1:] == ["polar"]:           # only the laser-range front-end KATs
        np.savez_compressed(os.path.join(GOLDEN, "frontend_polar_kats.npz"), **frontend_polar_kats())
        return
    if sys.argv[1:] == ["known"]:           # only the map-clustering KATs (the other files stay as committed)
        np.savez_compressed(os.path.join(GOLDEN, "known_landmarks_kats.npz"), **known_landmark_kats())
        return
    # A: drive around the room, 0..4 observations per step, both motion branches, several resamples
    np.savez_compressed(os.path.join(GOLDEN, "traj_drive.npz"),
                        **record_trajectory(24, sc.drive_stream(1, 60), np_seed=7, lcap=64))
    # B: repeated observations of the same landmarks inside one step (sequential dependency, Q7)
    np.savez_compressed(os.path.join(GOLDEN, "traj_repeat.npz"),
                        **record_trajectory(16, sc.repeat_stream(3, 12), np_seed=11, lcap=48))
    # C: benchmark-shaped state: pre-populated 6x6 grid map, 8 observations per step, 2 novel
    init = sc.synthetic_state(1234, 12, 36, 48)
    stream = []
    for s in range(6):
        rot, tr = sc.synthetic_odometry(s + 7)
        obs = sc.synthetic_obs(1234, s, init["world"], 8, novel=2 if s % 2 else 0, max_range=6.0)
        stream.append((rot, tr, [tuple(o) for o in obs]))
    np.savez_compressed(os.path.join(GOLDEN, "traj_synth.npz"),
                        **record_trajectory(12, stream, np_seed=5, lcap=48, init=init))
    np.savez_compressed(os.path.join(GOLDEN, "stage_kats.npz"), **stage_kats())
    np.savez_compressed(os.path.join(GOLDEN, "frontend_kats.npz"), **frontend_kats())
    np.savez_compressed(os.path.join(GOLDEN, "known_landmarks_kats.npz"), **known_landmark_kats())
    np.savez_compressed(os.path.join(GOLDEN, "frontend_polar_kats.npz"), **frontend_polar_kats())
    np.savez_compressed(os.path.join(GOLDEN, "icp_kats.npz"), **icp_kats())
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
