"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/fs2_oracle.c (the CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import
this.  The product path (fast_slam_b200/) never does and fails loudly without its CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfs2_oracle.so")
_SOURCES = ["fs2_oracle.c"]

ST_SINGULAR_LM, ST_SINGULAR_Q, ST_PDF_FAILED, ST_MAP_FULL = 1, 2, 4, 8


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, s) for s in _SOURCES if os.path.exists(os.path.join(_HERE, s))]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-std=c11", "-o", _SO] + srcs + ["-lm"]
        subprocess.run(cmd, check=True, cwd=_HERE)
    return _SO


_lib = None


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _bp(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        D, I64, I32 = C.c_double, C.c_int64, C.c_int
        PD, PI, PB = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        L.fs2o_wrap_pi.restype = D
        L.fs2o_wrap_pi.argtypes = [D]
        L.fs2o_mahalanobis.restype = D
        L.fs2o_mahalanobis.argtypes = [D] * 8 + [C.POINTER(C.c_int)]
        L.fs2o_associate.restype = I32
        L.fs2o_associate.argtypes = [D, D, PD, I32, I32, D]
        L.fs2o_motion.restype = None
        L.fs2o_motion.argtypes = [I64, PD, PD, PD, D, D, PD]
        L.fs2o_mvn_pdf2.restype = I32
        L.fs2o_mvn_pdf2.argtypes = [D] * 5 + [PD]
        L.fs2o_update.restype = None
        L.fs2o_update.argtypes = [I64, PD, PD, PD, PD, PI, PD, I32, PD, I32, PD, D, PI, PI, PB]
        L.fs2o_weight_total.restype = D
        L.fs2o_weight_total.argtypes = [I64, PD, PB, C.POINTER(C.c_int)]
        L.fs2o_normalize.restype = D
        L.fs2o_normalize.argtypes = [I64, PD, PB]
        L.fs2o_neff.restype = D
        L.fs2o_neff.argtypes = [I64, PD]
        L.fs2o_resample_indices.restype = I32
        L.fs2o_resample_indices.argtypes = [I64, PD, D, PI]
        L.fs2o_cumsum_seq.restype = None
        L.fs2o_cumsum_seq.argtypes = [I64, PD, PD]
        L.fs2o_gather.restype = None
        L.fs2o_gather.argtypes = [I64, PI, I32, PD, PD, PD, PD, PI, PD, PB, PD, PD, PD, PD, PI, PD, PB]
        L.fs2o_argmax.restype = I64
        L.fs2o_argmax.argtypes = [I64, PD]
        L.fs2o_step.restype = I32
        L.fs2o_step.argtypes = [I64, I32, PD, PD, PD, PD, PI, PD, PB, D, D, PD, PD, I32, D, PD, D,
                                PI, PI, PI, PD, PI, PD, PD]
        L.fs2o_step_noresample.restype = None
        L.fs2o_step_noresample.argtypes = [I64, I32, PD, PD, PD, PD, PI, PD, D, D, PD, PD, I32, PD, D, PD]
        L.fs2o_num_threads.restype = I32
        L.fs2o_set_num_threads.restype = I32
        L.fs2o_set_num_threads.argtypes = [I32]
        _lib = L
    return _lib


DEFAULT_R = np.array([0.001, 0.0, 0.0, 0.001])  # config.py:15
DEFAULT_GATE = 8.0                                  # config.py:18


def wrap_pi(a: float) -> float:
    return lib().fs2o_wrap_pi(float(a))


def mahalanobis(a, b, cov):
    s = C.c_int(0)
    cov = np.asarray(cov, dtype=np.float64)
    d = lib().fs2o_mahalanobis(float(a[0]), float(a[1]), float(b[0]), float(b[1]),
                               float(cov[0, 0]), float(cov[0, 1]), float(cov[1, 0]), float(cov[1, 1]), C.byref(s))
    return d, bool(s.value)


def mvn_pdf2(nu, q):
    out = C.c_double(0.0)
    q = np.asarray(q, dtype=np.float64)
    ok = lib().fs2o_mvn_pdf2(float(nu[0]), float(nu[1]), float(q[0, 0]), float(q[1, 0]), float(q[1, 1]), C.byref(out))
    return (out.value if ok else None)


def resample_indices(w, u0):
    w = np.ascontiguousarray(w, dtype=np.float64)
    idx = np.empty(len(w), dtype=np.int32)
    stuck = lib().fs2o_resample_indices(len(w), _dp(w), float(u0), _ip(idx))
    return idx, bool(stuck)


def cumsum_seq(w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    c = np.empty_like(w)
    lib().fs2o_cumsum_seq(len(w), _dp(w), _dp(c))
    return c


def neff(w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    return lib().fs2o_neff(len(w), _dp(w))


class OracleFilter:
    """SoA particle set + the oracle's step.  Mirrors FastSLAM2 (fast_slam_2.py:20-31) at P particles."""

    def __init__(self, num_particles: int, lcap: int, R=DEFAULT_R, gate=DEFAULT_GATE, track_pytypes=True):
        P = int(num_particles)
        self.P, self.lcap = P, int(lcap)
        self.R = np.ascontiguousarray(R, dtype=np.float64).reshape(4).copy()
        self.gate = float(gate)
        self.x = np.zeros(P)
        self.y = np.zeros(P)
        self.yaw = np.zeros(P)
        self.w = np.full(P, 1.0 / P)                       # particle.py:19
        self.count = np.zeros(P, dtype=np.int32)           # particle.py:20
        self.lm = np.zeros((P, self.lcap, 6))
        self.status = np.zeros(P, dtype=np.int32)
        self.wkind = np.ones(P, dtype=np.uint8) if track_pytypes else None
        self._sp = None

    # -- state exchange -------------------------------------------------------------------------
    def set_state(self, x, y, yaw, w, count, lm_p_l_6=None, lm=None):
        self.x[:] = x; self.y[:] = y; self.yaw[:] = yaw; self.w[:] = w; self.count[:] = count
        if lm_p_l_6 is not None:                           # [P][L][6], NaN padded -> [P][lcap][6]
            L = lm_p_l_6.shape[1]
            self.lm[:, :L, :] = np.nan_to_num(lm_p_l_6, nan=0.0)
        elif lm is not None:
            self.lm[:] = lm
        if self.wkind is not None:
            self.wkind[:] = 0

    def state(self):
        return dict(x=self.x.copy(), y=self.y.copy(), yaw=self.yaw.copy(), w=self.w.copy(),
                    counts=self.count.copy(), lm=self.lm.copy())

    def lm_p_l_6(self, L=None):
        L = int(self.count.max()) if L is None else L
        out = self.lm[:, :max(L, 1), :].copy()
        mask = np.arange(max(L, 1))[None, :] >= self.count[:, None]
        out[mask] = np.nan
        return out

    # -- stages ---------------------------------------------------------------------------------
    def motion(self, rotation, translation, noise):
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        lib().fs2o_motion(self.P, _dp(self.x), _dp(self.y), _dp(self.yaw), float(rotation), float(translation), _dp(noise))

    def update(self, obs):
        obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1, 2)
        M = obs.shape[0]
        assoc = np.empty((M, self.P), dtype=np.int32)
        lib().fs2o_update(self.P, _dp(self.x), _dp(self.y), _dp(self.yaw), _dp(self.w), _ip(self.count),
                          _dp(self.lm), self.lcap, _dp(obs), M, _dp(self.R), self.gate, _ip(assoc),
                          _ip(self.status), _bp(self.wkind))
        return assoc

    def normalize(self):
        return lib().fs2o_normalize(self.P, _dp(self.w), _bp(self.wkind))

    def neff(self):
        return lib().fs2o_neff(self.P, _dp(self.w))

    def step(self, rotation, translation, obs, noise, u0):
        """One filter step (fast_slam_2.py:33-67).  Returns dict like ref_harness.iterate_recorded."""
        P = self.P
        obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1, 2)
        M = obs.shape[0]
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        assoc = np.empty((M, P), dtype=np.int32)
        idx = np.empty(P, dtype=np.int32)
        if self._sp is None:
            self._sp = (np.empty(4 * P), np.empty(P, dtype=np.int32), np.empty_like(self.lm))
        sp, sc, sl = self._sp
        out = np.zeros(6)
        lib().fs2o_step(P, self.lcap, _dp(self.x), _dp(self.y), _dp(self.yaw), _dp(self.w), _ip(self.count),
                        _dp(self.lm), _bp(self.wkind), float(rotation), float(translation), _dp(noise),
                        _dp(obs), M, float(0.0 if u0 is None or np.isnan(u0) else u0), _dp(self.R), self.gate,
                        _ip(assoc), _ip(idx), _ip(self.status), _dp(sp), _ip(sc), _dp(sl), _dp(out))
        return dict(assoc=assoc, resample_idx=idx, estimate=out[:3].copy(), neff=out[3],
                    resampled=bool(out[4]), total=out[5])

    def step_noresample(self, rotation, translation, obs, noise):
        obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1, 2)
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        out = np.zeros(6)
        lib().fs2o_step_noresample(self.P, self.lcap, _dp(self.x), _dp(self.y), _dp(self.yaw), _dp(self.w),
                                   _ip(self.count), _dp(self.lm), float(rotation), float(translation),
                                   _dp(noise), _dp(obs), obs.shape[0], _dp(self.R), self.gate, _dp(out))
        return out
