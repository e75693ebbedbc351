"""TEST INFRASTRUCTURE ONLY -- drives the *unmodified* reference (cy-rae/fast-slam) in this container.

Used by ``oracle/gen_golden.py`` to freeze golden vectors under ``tests/golden/`` and by the
oracle-validation tests (only when ``/root/reference`` exists, i.e. never on the GPU box).
Nothing under ``fast_slam_b200/`` may import this module.

What it does (SURVEY.md section 8c "Harness rules"):
  * stubs the simulator module ``HAL`` so ``import fast_slam_2`` works
    (reference ``fast_slam_2/__init__.py:16-17`` imports Robot/EvaluationUtils which need it);
  * patches the constants that the reference binds by name at import
    (``fast_slam_2/algorithms/fast_slam_2.py:8``, ``fast_slam_2/models/particle.py:1``);
  * forces ``NUM_THREAD = 1`` so the thread pool is FIFO and the global ``np.random`` draw order is
    particle-index order (reference ``fast_slam_2.py:42-45``);
  * records every ``np.random.normal`` / ``np.random.uniform`` draw, every
    ``LandmarkUtils.associate_landmarks`` result (observation-major, particle-minor) and the
    resampling indices (particles are tagged before the step; ``deepcopy`` keeps the tag).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("FS2_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fast_slam_2"))


_loaded = None


def load_reference():
    """Import the reference package (HAL stubbed) and return a namespace of its modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    # The repo ships its own drop-in package that is also called ``fast_slam_2``; make sure the
    # reference's one wins inside this process.
    for name in [m for m in sys.modules if m == "fast_slam_2" or m.startswith("fast_slam_2.")]:
        del sys.modules[name]
    sys.modules.setdefault("HAL", types.ModuleType("HAL"))
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import fast_slam_2  # noqa: F401
        import fast_slam_2.algorithms.fast_slam_2 as alg
        import fast_slam_2.models.particle as particle_mod
        import fast_slam_2.utils.landmark_utils as lu_mod
        import fast_slam_2.utils.geometry_utils as gu_mod
        import fast_slam_2.algorithms.line_filter as lf_mod
        import fast_slam_2.algorithms.hough_transformation as ht_mod
        from fast_slam_2.models.landmark import Landmark
        from fast_slam_2.models.measurement import Measurement
    finally:
        sys.path.remove(REFERENCE_ROOT)
    assert os.path.realpath(alg.__file__).startswith(os.path.realpath(REFERENCE_ROOT))
    _loaded = types.SimpleNamespace(
        alg=alg, particle_mod=particle_mod, lu_mod=lu_mod, gu_mod=gu_mod, lf_mod=lf_mod,
        ht_mod=ht_mod, Landmark=Landmark, Measurement=Measurement,
        FastSLAM2=alg.FastSLAM2, LandmarkUtils=lu_mod.LandmarkUtils,
        GeometryUtils=gu_mod.GeometryUtils, LineFilter=lf_mod.LineFilter,
        HoughTransformation=ht_mod.HoughTransformation,
    )
    return _loaded


def unload_reference():
    """Drop the reference's ``fast_slam_2`` from ``sys.modules`` (so the repo's drop-in can load)."""
    global _loaded
    for name in [m for m in sys.modules if m == "fast_slam_2" or m.startswith("fast_slam_2.")]:
        del sys.modules[name]
    _loaded = None


class _Recorder:
    """Wraps np.random.normal/uniform and LandmarkUtils.associate_landmarks for one run."""

    def __init__(self, ref):
        self.ref = ref
        self.normals: list[float] = []
        self.uniforms: list[float] = []
        self.assoc: list[int] = []

    def __enter__(self):
        self._normal, self._uniform = np.random.normal, np.random.uniform
        self._assoc = self.ref.LandmarkUtils.associate_landmarks

        def normal(*a, **k):
            v = self._normal(*a, **k)
            self.normals.append(float(v))
            return v

        def uniform(*a, **k):
            v = self._uniform(*a, **k)
            self.uniforms.append(float(v))
            return v

        def associate(observed, landmarks):
            lm, idx = self._assoc(observed, landmarks)
            self.assoc.append(-1 if idx is None else int(idx))
            return lm, idx

        np.random.normal, np.random.uniform = normal, uniform
        # fast_slam_2.py:12 imported the class object, so patching the attribute on the class is
        # seen by FastSLAM2.__update_particle (fast_slam_2.py:104).
        self.ref.LandmarkUtils.associate_landmarks = staticmethod(associate)
        return self

    def __exit__(self, *exc):
        np.random.normal, np.random.uniform = self._normal, self._uniform
        self.ref.LandmarkUtils.associate_landmarks = staticmethod(self._assoc)
        return False


def new_filter(num_particles: int):
    """FastSLAM2() of the reference with NUM_PARTICLES patched (fast_slam_2.py:20-31)."""
    ref = load_reference()
    ref.alg.NUM_PARTICLES = num_particles
    ref.particle_mod.NUM_PARTICLES = num_particles
    ref.alg.NUM_THREAD = 1
    return ref.FastSLAM2()


def set_state(flt, x, y, yaw, w, counts, lm):
    """Overwrite the particle set of a reference filter from SoA arrays (lm: [P][L][6])."""
    ref = load_reference()
    for i, p in enumerate(flt.particles):
        p.x, p.y, p.yaw, p.weight = float(x[i]), float(y[i]), float(yaw[i]), float(w[i])
        p.landmarks = [
            ref.Landmark(float(lm[i, j, 0]), float(lm[i, j, 1]),
                         np.array([[lm[i, j, 2], lm[i, j, 3]], [lm[i, j, 4], lm[i, j, 5]]], dtype=np.float64))
            for j in range(int(counts[i]))
        ]


def get_state(flt, lcap: int | None = None):
    """SoA snapshot of a reference filter: x, y, yaw, w [P], counts [P], lm [P][lcap][6] (NaN padded)."""
    parts = flt.particles
    P = len(parts)
    counts = np.array([len(p.landmarks) for p in parts], dtype=np.int32)
    L = int(counts.max()) if lcap is None else lcap
    lm = np.full((P, max(L, 1), 6), np.nan)
    for i, p in enumerate(parts):
        for j, l in enumerate(p.landmarks):
            c = np.asarray(l.cov, dtype=np.float64)
            lm[i, j] = (l.x, l.y, c[0, 0], c[0, 1], c[1, 0], c[1, 1])
    return dict(
        x=np.array([float(p.x) for p in parts]), y=np.array([float(p.y) for p in parts]),
        yaw=np.array([float(p.yaw) for p in parts]), w=np.array([float(p.weight) for p in parts]),
        counts=counts, lm=lm,
    )


def iterate_recorded(flt, rotation: float, translation: float, measurements):
    """One reference ``iterate`` (fast_slam_2.py:33-67) with everything recorded.

    measurements: iterable of (distance, yaw).
    Returns dict(noise[P], u0 (nan if no resample), resampled (bool), assoc[M][P] int32,
    resample_idx[P] int32 (identity if not resampled), estimate[3]).
    """
    ref = load_reference()
    P = len(flt.particles)
    meas = [ref.Measurement(float(d), float(a)) for d, a in measurements]
    for i, p in enumerate(flt.particles):
        p._fs2_tag = i
    with _Recorder(ref) as rec, contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        est = flt.iterate(rotation, translation, meas)
    assert len(rec.normals) == P, (len(rec.normals), P)
    assoc = np.array(rec.assoc, dtype=np.int32).reshape(len(meas), P) if meas else np.zeros((0, P), np.int32)
    resampled = len(rec.uniforms) == 1
    idx = np.array([p._fs2_tag for p in flt.particles], dtype=np.int32)
    return dict(
        noise=np.array(rec.normals), u0=rec.uniforms[0] if resampled else float("nan"),
        resampled=resampled, assoc=assoc, resample_idx=idx,
        estimate=np.array([float(v) for v in est]),
    )
