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
    """list[Landmark] view of one particle's map.  The rows ([count][6]) are either given or fetched from the store
    on first use (one particle's map, fs2_download_particles) -- len() never touches the map."""

    def __init__(self, rows=None, fetch=None, count=None):
        self._rows = rows          # ndarray [count][6] or None until fetched
        self._fetch = fetch
        self._n = int(rows.shape[0] if rows is not None else count)

    def _r(self):
        if self._rows is None:
            self._rows = self._fetch()
        return self._rows

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        r = self._r()[i]
        return Landmark(float(r[0]), float(r[1]), np.array([[r[2], r[3]], [r[4], r[5]]]))

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class ParticleSet:
    """Sequence view returned by ``FastSLAM2.particles``: looks like list[Particle] (jde_robots_main.py:52,59;
    landmark_utils.py:126-131; serializer.py:39).  Host copies are taken lazily and PER FIELD: poses, weights and
    map lengths come from one small read (28 bytes per particle, no map traffic) -- all ``Serializer`` needs every
    loop iteration -- and a particle's landmarks are fetched for that particle alone when they are first touched.
    Only ``landmark_points()`` reads whole maps, in bounded chunks.

    ``store`` must offer download(maps=False) and download_particles(sel); ``epoch`` (optional) returns a counter that
    changes whenever the store is stepped: a view outlives its step only as long as nothing new has to be fetched."""

    def __init__(self, store, epoch=None):
        self.store = store          # the device store behind the view (LandmarkUtils.update_known_landmarks uses it)
        self._epoch_fn = epoch
        self._epoch = epoch() if epoch is not None else None
        self._snap = None

    def _check_fresh(self):
        if self._epoch_fn is not None and self._epoch_fn() != self._epoch:
            raise RuntimeError("this FastSLAM2.particles view belongs to an earlier filter step; read fast_slam.particles again")

    def _s(self):
        if self._snap is None:
            self._check_fresh()
            self._snap = self.store.download(maps=False)
        return self._snap

    def __len__(self):
        return len(self._s()["x"])

    def _rows_of(self, i, n):
        def fetch():
            self._check_fresh()
            return self.store.download_particles([i])["lm"][0, :n].copy()
        return fetch

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        s = self._s()
        if i < 0:
            i += len(s["x"])
        n = int(s["counts"][i])
        return Particle(float(s["x"][i]), float(s["y"][i]), float(s["yaw"][i]), float(s["w"][i]),
                        _LandmarkList(fetch=self._rows_of(int(i), n), count=n))

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    # bulk accessors for callers that do not want 10^6 Python objects
    def poses(self, max_particles=None):
        """[n][3] array of (x, y, yaw); with max_particles, every ceil(P / max_particles)-th particle."""
        s = self._s()
        a = np.stack([s["x"], s["y"], s["yaw"]], axis=1)
        if max_particles is not None and len(a) > int(max_particles) > 0:
            a = a[::-(-len(a) // int(max_particles))]
        return a

    def weights(self):
        return self._s()["w"]

    def counts(self):
        return self._s()["counts"]

    def landmark_points(self, chunk: int = 4096):
        """All landmarks of all particles as an [n][2] array (what update_known_landmarks collects), read from the
        store ``chunk`` particles at a time."""
        s = self._s()
        out = []
        P = len(s["x"])
        for lo in range(0, P, chunk):
            sel = np.arange(lo, min(P, lo + chunk))
            self._check_fresh()
            blk = self.store.download_particles(sel)
            mask = np.arange(blk["lm"].shape[1])[None, :] < blk["counts"][:, None]
            out.append(blk["lm"][mask][:, :2])
        return np.concatenate(out) if out else np.zeros((0, 2))
