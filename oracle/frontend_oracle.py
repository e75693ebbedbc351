"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's scan front-end (rows A11-A15 of SURVEY.md 8a):

    LandmarkUtils.get_measurements_to_landmarks        fast_slam_2/utils/landmark_utils.py:20-89
      LineFilter.filter                                fast_slam_2/algorithms/line_filter.py:12-21
      HoughTransformation.detect_line_intersections    fast_slam_2/algorithms/hough_transformation.py:14-145
      GeometryUtils.cluster_points                     fast_slam_2/utils/geometry_utils.py:26-62
      LandmarkUtils.__get_corners                      landmark_utils.py:66-89
      GeometryUtils.calculate_distance_and_angle       geometry_utils.py:65-74
    Robot.scan_environment (laser ranges -> points)    fast_slam_2/models/robot.py:32-58

The arithmetic lives in third-party code that is not under /root/reference: scipy.ndimage.gaussian_filter1d
(scipy 1.18.1), cv2.circle and cv2.HoughLines (opencv 4.13.0, modules/imgproc/src/hough.cpp HoughLinesStandard:
float32 sin/cos tables accumulated in float32, cvRound, 4-neighbour local maxima above the threshold, sorted by
votes then index), sklearn.cluster.DBSCAN with min_samples=1 (= connected components at eps, labels in order of
first appearance; sklearn 1.9.0).  Their published algorithms are restated here and pinned by executing the
reference in the build container (tests/golden/frontend_kats.npz, oracle/gen_golden.py).  Integer results (image
geometry, votes, peaks and their order, cluster labels, corner decisions) must be bit exact; the intersections go
through numpy's float32 cos/sin, so real-valued outputs agree to float32 rounding.
"""
from __future__ import annotations

import math

import numpy as np

PADDING = 20          # hough_transformation.py:10
SCALE = 100           # hough_transformation.py:11
HOUGH_THRESHOLD = 80  # hough_transformation.py:24
CLUSTER_EPS = 0.5     # landmark_utils.py:57
CORNER_THRESHOLD = 0.1  # landmark_utils.py:63
# cv2.circle(radius=2, thickness=-1): the 13 pixels with |dx| + |dy| <= 2
DISC = [(dx, dy) for dy in range(-2, 3) for dx in range(-2, 3) if abs(dx) + abs(dy) <= 2]


def scan_environment(values, angles, min_range: float, max_range: float):
    """models/robot.py:32-58 for any beam count: drop beams outside [min_range, max_range], the rest become
    (dist * cos(angle), dist * sin(angle)) with libm's cos / sin (math.cos / math.sin in the reference; its
    180-beam laser has angles[i] = np.radians(i - 90))."""
    pts = []
    for dist, ang in zip(values, angles):
        dist = float(dist)
        if dist < min_range or dist > max_range:
            continue
        pts.append([dist * math.cos(float(ang)), dist * math.sin(float(ang))])
    return np.array(pts, dtype=np.float64).reshape(-1, 2)


def gaussian_kernel1d(sigma: float, truncate: float = 4.0):
    """scipy.ndimage._filters._gaussian_kernel1d (order 0); radius = int(truncate * sigma + 0.5)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum(), radius


def line_filter(points, sigma: float = 0.1):
    """line_filter.py:12-21: gaussian_filter1d (mode='reflect') on both columns; identity at sigma = 0.1 (Q17)."""
    pts = np.asarray(points, dtype=np.float64)
    w, radius = gaussian_kernel1d(sigma)
    if radius == 0:
        return pts.copy()
    n = len(pts)
    out = np.empty_like(pts)
    for c in range(2):
        col = pts[:, c]
        for i in range(n):
            acc = 0.0
            for k in range(-radius, radius + 1):
                j = i + k
                # scipy 'reflect' = half-sample symmetric: (d c b a | a b c d | d c b a)
                while j < 0 or j >= n:
                    j = -j - 1 if j < 0 else 2 * n - 1 - j
                acc += w[k + radius] * col[j]     # correlate1d with the symmetric kernel
            out[i, c] = acc
    return out


def image_geometry(pts):
    """hough_transformation.py:52-66: scaled integer bounds (int() truncates toward zero), offsets, size."""
    min_x = int(np.min(pts[:, 0] * SCALE)); min_y = int(np.min(pts[:, 1] * SCALE))
    max_x = int(np.max(pts[:, 0] * SCALE)); max_y = int(np.max(pts[:, 1] * SCALE))
    off_x = (-min_x if min_x < 0 else 0) + PADDING
    off_y = (-min_y if min_y < 0 else 0) + PADDING
    return off_x, off_y, max_x + off_x + PADDING, max_y + off_y + PADDING


def rasterise(pts, off_x, off_y, width, height):
    """hough_transformation.py:68-71: a filled radius-2 disc per point (clipped by the image)."""
    img = np.zeros((height, width), dtype=np.uint8)
    for p in pts:
        x = int(p[0] * SCALE) + off_x
        y = int(p[1] * SCALE) + off_y
        for dx, dy in DISC:
            xx, yy = x + dx, y + dy
            if 0 <= xx < width and 0 <= yy < height:
                img[yy, xx] = 255
    return img


def hough_tables(numangle):
    """createTrigTable: float32 angle accumulated in float32, sin/cos evaluated in double, stored as float32."""
    theta = np.float32(np.pi / 180)
    tab_sin = np.zeros(numangle, np.float32)
    tab_cos = np.zeros(numangle, np.float32)
    ang = np.float32(0.0)
    for n in range(numangle):
        tab_sin[n] = np.float32(math.sin(float(ang)))
        tab_cos[n] = np.float32(math.cos(float(ang)))
        ang = np.float32(ang + theta)
    return tab_sin, tab_cos


NUMANGLE = 180


def hough_lines(img, threshold: int = HOUGH_THRESHOLD):
    """cv2.HoughLines(img, 1, pi/180, threshold) -> float32 [n][2] (rho, theta), strongest first."""
    h, w = img.shape
    numrho = 2 * (w + h) + 1
    tab_sin, tab_cos = hough_tables(NUMANGLE)
    acc = np.zeros((NUMANGLE + 2) * (numrho + 2), np.int32)
    ys, xs = np.nonzero(img)
    xf, yf = xs.astype(np.float32), ys.astype(np.float32)
    for n in range(NUMANGLE):
        r = np.rint(xf * tab_cos[n] + yf * tab_sin[n]).astype(np.int64) + (numrho - 1) // 2
        np.add.at(acc, (n + 1) * (numrho + 2) + r + 1, 1)
    a2 = acc.reshape(NUMANGLE + 2, numrho + 2)
    c = a2[1:-1, 1:-1]
    peak = (c > threshold) & (c > a2[1:-1, :-2]) & (c >= a2[1:-1, 2:]) & (c > a2[:-2, 1:-1]) & (c >= a2[2:, 1:-1])
    ns, rs = np.nonzero(peak)
    base = (ns + 1) * (numrho + 2) + rs + 1
    order = np.lexsort((base, -acc[base]))          # votes descending, then index ascending
    lines = np.zeros((len(base), 2), np.float32)
    thetaf = np.float32(np.pi / 180)
    for k, o in enumerate(order):
        n, r = int(ns[o]), int(rs[o])
        lines[k, 0] = np.float32((np.float32(r) - np.float32(numrho - 1) * np.float32(0.5)) * np.float32(1.0))
        lines[k, 1] = np.float32(np.float32(n) * thetaf)
    return lines, acc[base][order]


def intersections(lines, width, height):
    """hough_transformation.py:76-119 in the reference's float32 arithmetic (np.float32 scalars)."""
    out = []
    n = len(lines)
    lim = np.deg2rad(45)
    for i in range(n):
        for j in range(i + 1, n):
            rho1, theta1 = lines[i]
            rho2, theta2 = lines[j]
            d = np.float32(abs(theta1 - theta2))
            d = min(d, np.float32(np.float32(np.pi) - d) if False else np.float32(np.pi - d))
            if d < lim:
                continue
            a1, b1 = np.cos(theta1), np.sin(theta1)
            a2, b2 = np.cos(theta2), np.sin(theta2)
            det = a1 * b2 - a2 * b1
            if abs(det) > 1e-10:
                x = (b2 * rho1 - b1 * rho2) / det
                y = (a1 * rho2 - a2 * rho1) / det
                if 0 <= x < width and 0 <= y < height:
                    out.append((x, y))
    return np.array(out, dtype=np.float32).reshape(-1, 2)


def to_metres(px, off_x, off_y):
    """hough_transformation.py:142-145 (float32 - int, / int stay float32 under NEP 50)."""
    out = np.empty_like(px)
    out[:, 0] = (px[:, 0] - np.float32(off_x)) / np.float32(SCALE)
    out[:, 1] = (px[:, 1] - np.float32(off_y)) / np.float32(SCALE)
    return out


def cluster_labels(pts, eps: float = CLUSTER_EPS):
    """DBSCAN(eps, min_samples=1).labels_: connected components of the eps-graph, numbered by first appearance."""
    n = len(pts)
    lab = -np.ones(n, np.int64)
    p64 = pts.astype(np.float64)
    nxt = 0
    for i in range(n):
        if lab[i] >= 0:
            continue
        lab[i] = nxt
        stack = [i]
        while stack:
            a = stack.pop()
            d = np.sqrt(((p64 - p64[a]) ** 2).sum(1))
            for b in np.flatnonzero((d <= eps) & (lab < 0)):
                lab[b] = nxt
                stack.append(int(b))
        nxt += 1
    return lab


def centroids(pts, lab):
    """geometry_utils.py:49-60: per label (ascending), float32 mean accumulated sequentially in row order."""
    out = []
    for l in range(int(lab.max()) + 1 if len(lab) else 0):
        sel = pts[lab == l]
        s = np.zeros(2, np.float32)
        for row in sel:
            s = (s + row).astype(np.float32)
        out.append(s / np.float32(len(sel)))
    return np.array(out, dtype=np.float32).reshape(-1, 2)


def corners(cent, filtered, threshold: float = CORNER_THRESHOLD):
    """landmark_utils.py:66-89: keep a centroid if any scan point lies within the threshold."""
    keep = []
    f = np.asarray(filtered, dtype=np.float64)
    for c in cent:
        d = np.sqrt((np.float64(c[0]) - f[:, 0]) ** 2 + (np.float64(c[1]) - f[:, 1]) ** 2)
        keep.append(bool((d <= threshold).any()))
    return np.array(keep, dtype=bool)


def measurements(cent):
    """geometry_utils.py:72-73 on np.float32 coordinates: x**2 + y**2 in float32, sqrt/atan2 in double."""
    out = np.zeros((len(cent), 2))
    for k, c in enumerate(cent):
        q = np.float32(np.float32(c[0] * c[0]) + np.float32(c[1] * c[1]))
        out[k] = (math.sqrt(float(q)), math.atan2(float(c[1]), float(c[0])))
    return out


def get_measurements(points, sigma: float = 0.1, detail: bool = False):
    """LandmarkUtils.get_measurements_to_landmarks(points) -> float64 [K][2] (distance, yaw)."""
    filtered = line_filter(points, sigma)
    off_x, off_y, width, height = image_geometry(filtered)
    img = rasterise(filtered, off_x, off_y, width, height)
    lines, votes = hough_lines(img)
    px = intersections(lines, width, height)
    if len(px) == 0:
        res = np.zeros((0, 2))
        return (res, dict(lines=lines, votes=votes, geometry=(off_x, off_y, width, height))) if detail else res
    pm = to_metres(px, off_x, off_y)
    lab = cluster_labels(pm)
    cent = centroids(pm, lab)
    keep = corners(cent, filtered)
    res = measurements(cent[keep])
    if detail:
        return res, dict(lines=lines, votes=votes, geometry=(off_x, off_y, width, height), inter=pm, labels=lab,
                         centroids=cent, keep=keep, npix=int((img > 0).sum()))
    return res


# ---------------------------------------------------------------------------------------------------------------
# Restatements of the DEVICE's decompositions (csrc/fs2_frontend.cuh), kept next to the reference forms above so that
# the CPU suite can hold them equal: the kernels compute the same integers by another route, and these show the route.
# ---------------------------------------------------------------------------------------------------------------
def hough_lines_banded(img, threshold: int = HOUGH_THRESHOLD, band_rows: int = 10, row_words: int = 2304):
    """fe_vote_peaks: the accumulator in tiles of `band_rows` angle rows + one halo row either side (rows -1 and 180 are
    OpenCV's zero border and stay zero) by 2 * row_words - 2 rho cells + one halo cell either side; votes of the whole
    image into each tile, 4-neighbour maxima inside it; ranked like hough_lines.  Returns (lines, votes)."""
    h, w = img.shape
    numrho = 2 * (w + h) + 1
    half = (numrho - 1) // 2
    stride = numrho + 2
    tab_sin, tab_cos = hough_tables(NUMANGLE)
    ys, xs = np.nonzero(img)
    xf, yf = xs.astype(np.float32), ys.astype(np.float32)
    width = 2 * row_words - 2
    cand = []
    for n_first in range(0, NUMANGLE, band_rows):
        n0 = n_first - 1
        for r0 in range(0, numrho, width):
            wt = min(width, numrho - r0)
            acc = np.zeros((band_rows + 2, 2 * row_words), np.int64)
            for k in range(band_rows + 2):
                n = n0 + k
                if n < 0 or n >= NUMANGLE:
                    continue
                c = np.rint(xf * tab_cos[n] + yf * tab_sin[n]).astype(np.int64) + half - r0 + 1
                np.add.at(acc[k], c[(c >= 0) & (c <= wt + 1)], 1)
            assert acc.max() < 65536                      # the device's cells are 16 bits wide
            cen = acc[1:-1, 1:wt + 1]
            peak = ((cen > threshold) & (cen > acc[1:-1, 0:wt]) & (cen >= acc[1:-1, 2:wt + 2])
                    & (cen > acc[:-2, 1:wt + 1]) & (cen >= acc[2:, 1:wt + 1]))
            for k, c in zip(*np.nonzero(peak)):
                n = n0 + k + 1
                if n < NUMANGLE:
                    cand.append(((n + 1) * stride + (r0 + c) + 1, int(cen[k, c])))
    cand.sort(key=lambda t: (-t[1], t[0]))
    lines = np.zeros((len(cand), 2), np.float32)
    thetaf = np.float32(np.pi / 180)
    for i, (base, _) in enumerate(cand):
        n, r = base // stride - 1, base % stride - 1
        lines[i, 0] = np.float32(np.float32(r) - np.float32(numrho - 1) * np.float32(0.5))
        lines[i, 1] = np.float32(np.float32(n) * thetaf)
    return lines, np.array([v for _, v in cand], np.int64)


def sq_threshold(eps: float) -> float:
    """fe_sq_threshold (csrc/fs2.cu): the largest double whose correctly rounded square root is <= eps."""
    s = eps * eps
    while math.sqrt(s) > eps:
        s = math.nextafter(s, 0.0)
    while math.sqrt(math.nextafter(s, math.inf)) <= eps:
        s = math.nextafter(s, math.inf)
    return s


def cluster_labels_two_phase(pts, eps: float = CLUSTER_EPS, order=None):
    """fe_intersect_cluster's connected components: (1) every point takes its first neighbour j < i as parent, flatten;
    (2) one sweep over the pairs within eps (in any order: `order` permutes them) with the points' labels frozen merges
    the roots (atomicMin hooking: the larger root under the smaller, a displaced parent linked in turn); flatten.  Distances
    are compared squared against sq_threshold(eps).  Returns first-appearance labels like cluster_labels."""
    p = np.asarray(pts, np.float32).astype(np.float64)
    n = len(p)
    T = sq_threshold(eps)
    d2 = (p[:, None, 0] - p[None, :, 0]) ** 2 + (p[:, None, 1] - p[None, :, 1]) ** 2
    near = d2 <= T
    lab = np.arange(n)
    for i in range(n):
        nz = np.flatnonzero(near[i, :i])
        if len(nz):
            lab[i] = nz[0]

    def find(i):
        while lab[i] != i:
            i = lab[i]
        return i

    for i in range(n):
        lab[i] = find(i)
    frozen = lab.copy()
    pairs = [(i, j) for i in range(n) for j in range(i) if near[i, j]]
    if order is not None:
        pairs = [pairs[k] for k in order.permutation(len(pairs))]
    for i, j in pairs:
        la, lb = frozen[i], lab[j]
        if la == lb:
            continue
        a, b = max(la, lb), min(la, lb)
        if lab[a] == b:
            continue
        while a != b:
            if a < b:
                a, b = b, a
            old = lab[a]
            lab[a] = min(old, b)
            if old == a:
                break
            a = old
    for i in range(n):
        lab[i] = find(i)
    first = {}
    return np.array([first.setdefault(int(r), len(first)) for r in lab], np.int64)
