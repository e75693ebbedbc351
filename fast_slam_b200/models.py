"""Data model of the reference (fast_slam_2/models/{point,directed_point,landmark,particle,measurement}.py)
as the API surface of the HBM store: plain value classes for what the caller constructs (Measurement,
Landmark, Point) and read-only views for what the filter owns (Particle and its landmarks)."""
from __future__ import annotations

import numpy as np

_DEFAULT_COV = ((0.1, 0.0), (0.0, 0.1))   # landmark.py:13


class Point:
    """2-D point (point.py:9-33)."""
    __slots__ = ("x", "y")

    def __init__(self, x: float, y: float):
        self.x, self.y = x, y

    def as_vector(self):
        return np.array([self.x, self.y])

    def to_dict(self):
        return {"x": self.x, "y": self.y}


class DirectedPoint(Point):
    """2-D pose (directed_point.py:9-28)."""
    __slots__ = ("yaw",)

    def __init__(self, x: float, y: float, yaw: float):
        super().__init__(x, y)
        self.yaw = yaw

    def to_dict(self):
        d = super().to_dict()
        d["yaw"] = self.yaw
        return d


class Landmark(Point):
    """Landmark mean + 2x2 covariance, default 0.1*I (landmark.py:13-21)."""
    __slots__ = ("cov",)

    def __init__(self, x: float, y: float, cov=None):
        super().__init__(x, y)
        self.cov = np.array(_DEFAULT_COV) if cov is None else cov

    def __str__(self):
        return f"Landmark ID: x: {self.x}, y: {self.y}, Covariance: {self.cov}"


class Measurement:
    """Range/bearing of an observed landmark from the robot (measurement.py:9-23)."""
    __slots__ = ("distance", "yaw")

    def __init__(self, distance: float, yaw: float):
        self.distance, self.yaw = distance, yaw

    def as_vector(self):
        return np.array([self.distance, self.yaw])


class Particle(DirectedPoint):
    """One particle as the reference exposes it (particle.py:11-27): pose, weight, landmark list.
    Instances handed out by FastSLAM2.particles are snapshots of the HBM store; values are Python
    floats (JSON-serialisable, serializer.py:39)."""
    __slots__ = ("weight", "landmarks")

    def __init__(self, x: float, y: float, yaw: float, weight: float = None, landmarks=None):
        super().__init__(x, y, yaw)
        if weight is None:
            from . import config
            weight = 1.0 / config.NUM_PARTICLES
        self.weight = weight
        self.landmarks = [] if landmarks is None else landmarks

    def __str__(self):
        return f"Particle: x: {self.x}, y: {self.y}, yaw: {self.yaw}, weight: {self.weight}, landmarks: {self.landmarks}"


class _LandmarkList:
    """list[Landmark] view of one particle's rows of the downloaded map block."""

    def __init__(self, rows):
        self._rows = rows          # ndarray [count][6]

    def __len__(self):
        return self._rows.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        r = self._rows[i]
        return Landmark(float(r[0]), float(r[1]), np.array([[r[2], r[3]], [r[4], r[5]]]))

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class ParticleSet:
    """Sequence view returned by ``FastSLAM2.particles``: looks like list[Particle] (jde_robots_main.py:52,59;
    landmark_utils.py:126-131; serializer.py:39) over one host snapshot of the store, taken lazily."""

    def __init__(self, snapshot_fn, store=None):
        self._fn = snapshot_fn
        self._snap = None
        self.store = store          # the device store behind the view (LandmarkUtils.update_known_landmarks uses it)

    def _s(self):
        if self._snap is None:
            self._snap = self._fn()
        return self._snap

    def __len__(self):
        return len(self._s()["x"])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        s = self._s()
        n = int(s["counts"][i])
        return Particle(float(s["x"][i]), float(s["y"][i]), float(s["yaw"][i]), float(s["w"][i]),
                        _LandmarkList(s["lm"][i, :n]))

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    # bulk accessors for callers that do not want 10^6 Python objects
    def poses(self):
        s = self._s()
        return np.stack([s["x"], s["y"], s["yaw"]], axis=1)

    def weights(self):
        return self._s()["w"]

    def landmark_points(self):
        """All landmarks of all particles as an [n][2] array (what update_known_landmarks collects)."""
        s = self._s()
        mask = np.arange(s["lm"].shape[1])[None, :] < s["counts"][:, None]
        return s["lm"][mask][:, :2]
