"""FastSLAM2 -- the reference's class (fast_slam_2/algorithms/fast_slam_2.py:15-223) over the HBM store.

Same constructor (no arguments, reads config), same ``iterate(rotation, translation, measurements)``,
same ``particles`` attribute; the Python-object particle set and the NUM_THREAD thread pool are replaced
by one call into the C ABI per step (fs2_step_host) -- or, in RNG = "reference" mode, by the stage entry
points with the random draws taken from the global np.random in the reference's order."""
from __future__ import annotations

import numpy as np

from . import config
from .models import ParticleSet
from .store import DeviceFilter
from . import _lib


def _hash_uniform(seed: int, step: int) -> float:
    """splitmix64 -> [0, 1): the resampling start point of RNG = "device" mode."""
    z = (seed * 0x9E3779B97F4A7C15 + step * 0xBF58476D1CE4E5B9 + 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    z ^= z >> 31
    return (z >> 11) / 9007199254740992.0


class FastSLAM2:
    """Drop-in for fast_slam_2.FastSLAM2 (jde_robots_main.py:13,38)."""

    def __init__(self):
        self._n = int(config.NUM_PARTICLES)                       # fast_slam_2.py:25-31
        self._rng_mode = str(config.RNG)
        self._seed = int(config.SEED)
        self._store = DeviceFilter(
            self._n, int(config.LANDMARK_CAPACITY), device=config.DEVICE,
            translation_noise=float(config.TRANSLATION_NOISE), rotation_noise=float(config.ROTATION_NOISE),
            measurement_noise=np.asarray(config.MEASUREMENT_NOISE, dtype=np.float64),
            max_landmark_distance=float(config.MAXIMUM_LANDMARK_DISTANCE), seed=self._seed)
        self._step = 0
        self._particles = None
        self.last = None          # diagnostics of the last step (neff, resampled, total)

    # the reference exposes a plain list attribute; here it is a lazy snapshot of the store
    @property
    def particles(self):
        if self._particles is None:
            self._particles = ParticleSet(self._store, epoch=lambda: self._step)
        return self._particles

    @property
    def store(self) -> DeviceFilter:
        return self._store

    def iterate(self, rotation: float, translation: float, measurements) -> tuple[float, float, float]:
        """One filter step (fast_slam_2.py:33-67): move, update per measurement, normalise, resample when
        Neff < N/2, return the pose of the heaviest particle."""
        n = len(measurements)
        obs = np.empty((n, 2))
        for k, m in enumerate(measurements):
            obs[k, 0] = m.distance                                 # measurement.py:15-16
            obs[k, 1] = m.yaw
        rotation, translation = float(rotation), float(translation)
        st = self._store
        if self._rng_mode == "reference":
            est = self._iterate_reference_stream(rotation, translation, obs)
        else:
            u0 = _hash_uniform(self._seed, self._step) / self._n   # U(0, 1/N), fast_slam_2.py:183
            r = st.step(rotation, translation, obs, noise=None, u0=u0, step_index=self._step,
                        want_assoc=False, want_ancestor=False)
            self.last = r
            if r["resampled"]:
                print("\nRESAMPLING")                               # fast_slam_2.py:63
            est = r["estimate"]
        self._step += 1
        self._particles = None
        return float(est[0]), float(est[1]), float(est[2])

    def _iterate_reference_stream(self, rotation, translation, obs):
        st = self._store
        sigma = float(config.ROTATION_NOISE) if rotation != 0 else float(config.TRANSLATION_NOISE)
        noise = np.random.normal(0, sigma, self._n)                # == N scalar draws in index order (Q15)
        st.motion_update(rotation, translation, obs, noise=noise)
        st.weight_total()
        st.normalize()
        stats = st.stats.cpu().numpy()
        resampled = bool(stats[_lib.STAT_NEFF] < self._n / 2)       # fast_slam_2.py:62
        if resampled:
            print("\nRESAMPLING")
            u0 = float(np.random.uniform(0, 1 / self._n))           # fast_slam_2.py:183
            anc = st.resample_indices(u0)
            st.gather(anc)
            st.estimate()
            stats = st.stats.cpu().numpy()
        self.last = dict(neff=float(stats[_lib.STAT_NEFF]), resampled=resampled, total=float(stats[_lib.STAT_TOTAL]))
        return stats[_lib.STAT_EST_X:_lib.STAT_EST_YAW + 1]
