"""Scan front-end: the reference's LandmarkUtils / GeometryUtils / LineFilter surface (fast_slam_2/utils/
landmark_utils.py, utils/geometry_utils.py, algorithms/line_filter.py) over the batched CUDA front-end
(fs2_frontend in include/fs2.h, kernels in csrc/fs2_frontend.cuh).  No CPU implementation of the pipeline lives
here: without the CUDA library / a GPU the calls raise."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import check
from .models import Landmark, Measurement


def frontend_batch(scans, sigma: float = 0.1, device: int | None = None):
    """scans: array [B][N][2] (x, y) in the robot frame.  Returns (meas [B][Kmax][2], k [B], status [B])."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.Fs2Error("fast_slam_b200 front-end needs a CUDA device; there is no CPU fallback")
    L = _lib.load()
    scans = np.ascontiguousarray(scans, dtype=np.float64)
    assert scans.ndim == 3 and scans.shape[2] == 2
    B, N = scans.shape[:2]
    kmax = L.fs2_frontend_max_measurements()
    meas = np.zeros((B, kmax, 2))
    k = np.zeros(B, np.int32)
    status = np.zeros(B, np.int32)
    dev = torch.cuda.current_device() if device is None else int(device)
    torch.zeros(1, device="cuda:%d" % dev)
    pd = C.POINTER(C.c_double)
    pi = C.POINTER(C.c_int32)
    check(L.fs2_frontend(scans.ctypes.data_as(pd), B, N, float(sigma), dev, meas.ctypes.data_as(pd), k.ctypes.data_as(pi),
                         status.ctypes.data_as(pi), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs2_frontend")
    return meas, k, status


def frontend_batch_polar(ranges, angles, min_range: float, max_range: float, sigma: float = 0.1, device: int | None = None):
    """Laser messages in, measurements out: Robot.scan_environment (models/robot.py:32-58) fused in front of the
    batched front-end.  ranges: [B][N] beam ranges, angles: [N] beam angles (radians).  Beams outside
    [min_range, max_range] are dropped on the device, so every scan keeps its own number of points.
    Returns (meas [B][Kmax][2], k [B], status [B])."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.Fs2Error("fast_slam_b200 front-end needs a CUDA device; there is no CPU fallback")
    L = _lib.load()
    ranges = np.ascontiguousarray(ranges, dtype=np.float64)
    angles = np.ascontiguousarray(angles, dtype=np.float64)
    assert ranges.ndim == 2 and angles.shape == (ranges.shape[1],)
    B, N = ranges.shape
    kmax = L.fs2_frontend_max_measurements()
    meas = np.zeros((B, kmax, 2))
    k = np.zeros(B, np.int32)
    status = np.zeros(B, np.int32)
    dev = torch.cuda.current_device() if device is None else int(device)
    torch.zeros(1, device="cuda:%d" % dev)
    pd = C.POINTER(C.c_double)
    pi = C.POINTER(C.c_int32)
    check(L.fs2_frontend_polar(ranges.ctypes.data_as(pd), angles.ctypes.data_as(pd), B, N, float(min_range), float(max_range),
                               float(sigma), dev, meas.ctypes.data_as(pd), k.ctypes.data_as(pi), status.ctypes.data_as(pi),
                               C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs2_frontend_polar")
    return meas, k, status


class LineFilter:
    """line_filter.py:6-21: Gaussian filter along the point index, x and y separately (scipy gaussian_filter1d,
    reflect boundary).  Inside ``get_measurements_to_landmarks`` the filter is the first kernel of the batched
    front-end; this entry point runs that kernel alone (fs2_line_filter)."""

    @staticmethod
    def filter(points, sigma=0.1, device: int | None = None):
        pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 2))
        if int(4.0 * sigma + 0.5) == 0:        # radius 0: the gaussian is the identity (quirk Q17)
            return pts.copy()
        import torch
        if not torch.cuda.is_available():
            raise _lib.Fs2Error("fast_slam_b200 front-end needs a CUDA device; there is no CPU fallback")
        L = _lib.load()
        out = np.empty_like(pts)
        dev = torch.cuda.current_device() if device is None else int(device)
        torch.zeros(1, device="cuda:%d" % dev)
        pd = C.POINTER(C.c_double)
        check(L.fs2_line_filter(pts.ctypes.data_as(pd), 1, len(pts), float(sigma), dev, out.ctypes.data_as(pd),
                                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs2_line_filter")
        return out


class HoughTransformation:
    """hough_transformation.py:5-145: Hough image, cv2.HoughLines and the intersections of the detected lines, as
    the front-end's kernels compute them (fs2_hough_intersections)."""

    @staticmethod
    def detect_line_intersections(points, device: int | None = None):
        """points: [N][2] (already filtered) -> list of (x, y) in metres (np.float32, like the reference)."""
        import torch
        if not torch.cuda.is_available():
            raise _lib.Fs2Error("fast_slam_b200 front-end needs a CUDA device; there is no CPU fallback")
        L = _lib.load()
        pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 2))
        cap = L.fs2_hough_max_intersections()
        inter = np.zeros((1, cap, 2), np.float32)
        n = np.zeros(1, np.int32)
        status = np.zeros(1, np.int32)
        dev = torch.cuda.current_device() if device is None else int(device)
        torch.zeros(1, device="cuda:%d" % dev)
        check(L.fs2_hough_intersections(pts.ctypes.data_as(C.POINTER(C.c_double)), 1, len(pts), dev,
                                        inter.ctypes.data_as(C.POINTER(C.c_float)), n.ctypes.data_as(C.POINTER(C.c_int32)),
                                        status.ctypes.data_as(C.POINTER(C.c_int32)),
                                        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs2_hough_intersections")
        return [(inter[0, i, 0], inter[0, i, 1]) for i in range(int(n[0]))]


class ICP:
    """icp.py:5-90: rigid alignment of two 2-D point sets (the reference keeps it for odometry from scans and does not
    call it in its loop).  One thread block per pair on the device (fs2_icp)."""

    @staticmethod
    def get_transformation_batch(sources, targets, max_iterations: int = 100, threshold: float = 1e-5, device: int | None = None):
        """sources [B][Ns][2], targets [B][Nt][2] -> (rotations [B][2][2], translations [B][2], iterations [B])"""
        import torch
        if not torch.cuda.is_available():
            raise _lib.Fs2Error("fast_slam_b200 ICP needs a CUDA device; there is no CPU fallback")
        L = _lib.load()
        src = np.ascontiguousarray(sources, dtype=np.float64)
        tgt = np.ascontiguousarray(targets, dtype=np.float64)
        assert src.ndim == 3 and tgt.ndim == 3 and src.shape[0] == tgt.shape[0] and src.shape[2] == tgt.shape[2] == 2
        B = src.shape[0]
        rot = np.zeros((B, 2, 2)); tr = np.zeros((B, 2)); it = np.zeros(B, np.int32)
        dev = torch.cuda.current_device() if device is None else int(device)
        torch.zeros(1, device="cuda:%d" % dev)
        pd = C.POINTER(C.c_double)
        check(L.fs2_icp(src.ctypes.data_as(pd), tgt.ctypes.data_as(pd), B, src.shape[1], tgt.shape[1], int(max_iterations),
                        float(threshold), dev, rot.ctypes.data_as(pd), tr.ctypes.data_as(pd),
                        it.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs2_icp")
        return rot, tr, it

    @staticmethod
    def get_transformation(source_points, target_points, max_iterations=100, threshold=1e-5):
        """icp.py:13-58 -> (rotation matrix [2][2], translation vector [2])"""
        rot, tr, _ = ICP.get_transformation_batch(np.asarray(source_points, dtype=np.float64)[None],
                                                  np.asarray(target_points, dtype=np.float64)[None], max_iterations, threshold)
        return rot[0], tr[0]

    @staticmethod
    def best_fit_transform(source_points, target_points):
        """icp.py:60-90 for matched pairs (row i of one belongs to row i of the other): 2 x 2 algebra on the host, in
        the closed form the kernel uses (rotation by atan2(H01 - H10, H00 + H11))."""
        s = np.asarray(source_points, dtype=np.float64).reshape(-1, 2)
        t = np.asarray(target_points, dtype=np.float64).reshape(-1, 2)
        cs, ct = s.mean(axis=0), t.mean(axis=0)
        h = (s - cs).T @ (t - ct)
        th = math.atan2(h[0, 1] - h[1, 0], h[0, 0] + h[1, 1])
        r = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
        return r, ct - r @ cs


class GeometryUtils:
    """geometry_utils.py:8-74: the scalar helpers the reference's callers use on the host."""

    @staticmethod
    def mahalanobis_distance(position_a, position_b, covariance_matrix) -> float:
        # geometry_utils.py:14-23 in the closed form the kernels use (adj/det, one reciprocal)
        c = np.asarray(covariance_matrix, dtype=np.float64)
        det = c[0, 0] * c[1, 1] - c[0, 1] * c[1, 0]
        if det == 0.0:
            raise np.linalg.LinAlgError("Singular matrix")
        r = 1.0 / det
        dx, dy = float(position_b[0]) - float(position_a[0]), float(position_b[1]) - float(position_a[1])
        t0 = dx * (c[1, 1] * r) + dy * (-c[1, 0] * r)
        t1 = dx * (-c[0, 1] * r) + dy * (c[0, 0] * r)
        with np.errstate(invalid="ignore"):
            return float(np.sqrt(t0 * dx + t1 * dy))

    @staticmethod
    def cluster_points(point_lists, eps: float, min_samples: int, device: int | None = None, with_members: bool = False):
        """geometry_utils.py:26-62: DBSCAN(eps, min_samples) on the device (fs2_cluster_points), one centroid per
        cluster in label order.  Returns a list of (x, y) arrays like the reference."""
        import torch
        if not torch.cuda.is_available():
            raise _lib.Fs2Error("fast_slam_b200 clustering needs a CUDA device; there is no CPU fallback")
        L = _lib.load()
        pts = np.ascontiguousarray(np.asarray(point_lists, dtype=np.float64).reshape(-1, 2))
        n = len(pts)
        cap = max(n, 1)
        cent = np.zeros((cap, 2))
        mem = np.zeros(cap, np.int64)
        k = C.c_int32(0)
        dev = torch.cuda.current_device() if device is None else int(device)
        torch.zeros(1, device="cuda:%d" % dev)
        check(L.fs2_cluster_points(pts.ctypes.data_as(C.POINTER(C.c_double)), n, float(eps), int(min_samples), dev, cap,
                                   cent.ctypes.data_as(C.POINTER(C.c_double)), mem.ctypes.data_as(C.POINTER(C.c_int64)),
                                   C.byref(k), None), "fs2_cluster_points")
        out = [cent[i].copy() for i in range(k.value)]
        return (out, mem[:k.value].copy()) if with_members else out

    @staticmethod
    def calculate_distance_and_angle(x: float, y: float):
        return math.sqrt(x ** 2 + y ** 2), math.atan2(y, x)                      # geometry_utils.py:72-74


class LandmarkUtils:
    """landmark_utils.py:13-144 (front-end and association entry points)."""
    known_landmarks: list = []                                                    # landmark_utils.py:18

    @staticmethod
    def get_measurements_to_landmarks(scanned_points) -> list:
        meas, k, _ = frontend_batch(np.asarray(scanned_points, dtype=np.float64)[None])
        return [Measurement(float(d), float(a)) for d, a in meas[0, :k[0]]]

    @staticmethod
    def get_measurements_batch(scans, sigma: float = 0.1):
        """Many scans at once (BASELINE.json config 5): list of float64 [K_b][2] arrays."""
        meas, k, _ = frontend_batch(scans, sigma)
        return [meas[b, :k[b]].copy() for b in range(len(k))]

    @staticmethod
    def get_measurements_from_laser(values, min_range: float, max_range: float, angles=None) -> list:
        """One laser message (HAL.getLaserData(): .values, .minRange, .maxRange) -> list[Measurement]:
        Robot.scan_environment (models/robot.py:32-58) + get_measurements_to_landmarks in one device call.
        angles default to the reference's 180-beam layout, radians(i - 90)."""
        values = np.asarray(values, dtype=np.float64)
        if angles is None:
            angles = np.radians(np.arange(len(values)) - 90)
        meas, k, _ = frontend_batch_polar(values[None], angles, min_range, max_range)
        return [Measurement(float(d), float(a)) for d, a in meas[0, :k[0]]]

    @staticmethod
    def update_known_landmarks(particles):
        """landmark_utils.py:120-144: cluster every particle's landmarks (DBSCAN, eps 0.5, min_samples = 70 % of
        the average map length) into ``known_landmarks``.  ``FastSLAM2.particles`` is clustered where it lives, in
        device memory (fs2_known_landmarks); a plain list of Particle objects is uploaded as points."""
        store = getattr(particles, "store", None)
        if store is not None:
            res = store.known_landmarks()
            if res is None:
                return
            cent = res[0]
        else:
            pts = [(lm.x, lm.y) for p in particles for lm in p.landmarks]
            min_samples = int(len(pts) / len(particles) * 0.7)
            if min_samples < 1:
                return
            cent = GeometryUtils.cluster_points(pts, eps=0.5, min_samples=min_samples)
        LandmarkUtils.known_landmarks = [Landmark(float(c[0]), float(c[1])) for c in cent]

    @staticmethod
    def associate_landmarks(observed_landmark, particle_landmarks):
        """landmark_utils.py:92-117: first landmark in list order inside the Mahalanobis gate (host helper; the
        filter itself associates on the device)."""
        from . import config
        for i, lm in enumerate(particle_landmarks):
            d = GeometryUtils.mahalanobis_distance(lm.as_vector(), observed_landmark.as_vector(), lm.cov)
            if d < config.MAXIMUM_LANDMARK_DISTANCE:
                return lm, i
        return None, None
