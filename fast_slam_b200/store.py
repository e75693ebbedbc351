"""DeviceFilter -- the HBM particle/landmark store and the stage entry points of libfs2.so, from Python.

This is the thin layer north_star asks for: PyTorch supplies the CUDA context, streams, scratch tensors
and (in dist.py) torch.distributed; every computation is a call into the C ABI (include/fs2.h).  There
is no CPU implementation here -- without the CUDA library the constructor raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Fs2Config, Fs2Ptrs, Fs2StepResult, check


class _DevArray:
    """Zero-copy view of device memory owned by the C library (``__cuda_array_interface__`` v2)."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        self._owner = owner


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


class DeviceFilter:
    """Owns one fs2 handle = one GPU's shard of the particle set."""

    def __init__(self, num_particles: int, landmark_capacity: int, device: int | None = None,
                 translation_noise: float = 0.0055, rotation_noise: float = 0.001,
                 measurement_noise=((0.001, 0.0), (0.0, 0.001)), max_landmark_distance: float = 8.0,
                 seed: int = 0, flags: int = 0, global_particles: int = 0, global_offset: int = 0, spare_slots: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise _lib.Fs2Error("fast_slam_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._torch = torch
        self._L = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device="cuda:%d" % self.device)  # make sure the primary context exists
        cfg = Fs2Config()
        cfg.num_particles = int(num_particles)
        cfg.global_particles = int(global_particles)
        cfg.global_offset = int(global_offset)
        cfg.landmark_capacity = int(landmark_capacity)
        cfg.device = self.device
        cfg.flags = int(flags)
        cfg.spare_slots = int(spare_slots)
        cfg.translation_noise = float(translation_noise)
        cfg.rotation_noise = float(rotation_noise)
        r = np.asarray(measurement_noise, dtype=np.float64).reshape(4)
        for i in range(4):
            cfg.measurement_noise[i] = float(r[i])
        cfg.max_landmark_distance = float(max_landmark_distance)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.cfg = cfg
        h = C.c_void_p()
        check(self._L.fs2_create(C.byref(cfg), C.byref(h)), "fs2_create")
        self._h = h
        self.P = int(num_particles)
        self.Pglobal = int(global_particles) if global_particles else self.P
        self.lcap = int(landmark_capacity)
        p = Fs2Ptrs()
        check(self._L.fs2_get_ptrs(self._h, C.byref(p)), "fs2_get_ptrs")
        self.ptrs = p
        dev = "cuda:%d" % self.device
        as_t = lambda ptr, shape, ts: torch.as_tensor(_DevArray(ptr, shape, ts, self), device=dev)  # noqa: E731
        self.x = as_t(p.x, (self.P,), "<f8")
        self.y = as_t(p.y, (self.P,), "<f8")
        self.yaw = as_t(p.yaw, (self.P,), "<f8")
        self.w = as_t(p.w, (self.P,), "<f8")
        self.count = as_t(p.count, (self.P,), "<i4")
        self.status = as_t(p.status, (self.P,), "<i4")
        self.noise = as_t(p.noise, (self.P,), "<f8")
        self.cumsum = as_t(p.cumsum, (self.Pglobal,), "<f8")
        self.ancestor = as_t(p.ancestor, (self.P,), "<i4")
        self.stats = as_t(p.stats, (_lib.FS2_STATS_LEN,), "<f8")
        # raw map storage [slot][lcap][6]; particle p's map is slot p only until the first resample
        # (copy-on-resample permutes a private slot table) -- use download() for canonical maps.
        self.lm_raw = as_t(p.lm, (self.P + int(spare_slots), self.lcap, 6), "<f8")

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.fs2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launches(self) -> int:
        return int(self._L.fs2_launch_count(self._h))

    # -- state exchange ---------------------------------------------------------------------------
    def reset(self):
        check(self._L.fs2_reset(self._h, self._stream()), "fs2_reset")

    def upload(self, x=None, y=None, yaw=None, w=None, count=None, lm=None):
        c = lambda a, dt: None if a is None else np.ascontiguousarray(a, dtype=dt)  # noqa: E731
        x, y, yaw, w, lm = (c(a, np.float64) for a in (x, y, yaw, w, lm))
        count = c(count, np.int32)
        if lm is not None:
            assert lm.shape == (self.P, self.lcap, 6), lm.shape
        check(self._L.fs2_upload_state(self._h, _pd(x), _pd(y), _pd(yaw), _pd(w), _pi(count), _pd(lm), self._stream()),
              "fs2_upload_state")

    def download(self, maps: bool = True):
        P = self.P
        out = dict(x=np.empty(P), y=np.empty(P), yaw=np.empty(P), w=np.empty(P), counts=np.empty(P, np.int32),
                   status=np.empty(P, np.int32))
        lm = np.empty((P, self.lcap, 6)) if maps else None
        check(self._L.fs2_download_state(self._h, _pd(out["x"]), _pd(out["y"]), _pd(out["yaw"]), _pd(out["w"]),
                                         _pi(out["counts"]), _pd(lm), _pi(out["status"]), self._stream()),
              "fs2_download_state")
        if maps:
            out["lm"] = lm
        return out

    def download_particles(self, sel):
        sel = np.ascontiguousarray(sel, dtype=np.int64)
        n = len(sel)
        out = dict(x=np.empty(n), y=np.empty(n), yaw=np.empty(n), w=np.empty(n), counts=np.empty(n, np.int32),
                   lm=np.empty((n, self.lcap, 6)))
        check(self._L.fs2_download_particles(self._h, sel.ctypes.data_as(C.POINTER(C.c_int64)), n, _pd(out["x"]),
                                             _pd(out["y"]), _pd(out["yaw"]), _pd(out["w"]), _pi(out["counts"]),
                                             _pd(out["lm"]), self._stream()), "fs2_download_particles")
        return out

    def host_state(self, lcap=None):
        """Snapshot in the layout of the golden files: lm [P][lcap][6], NaN beyond each particle's count."""
        st = self.download()
        L = self.lcap if lcap is None else lcap
        lm = st["lm"][:, :L, :].copy()
        lm[np.arange(L)[None, :] >= st["counts"][:, None]] = np.nan
        st["lm"] = lm
        return st

    # -- stages -----------------------------------------------------------------------------------
    def _obs(self, obs):
        obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1, 2)
        return obs, obs.shape[0]

    def draw_noise(self, sigma: float, step: int, out=None):
        ptr = C.c_void_p(out.data_ptr()) if out is not None else None
        check(self._L.fs2_draw_noise(self._h, float(sigma), int(step), ptr, self._stream()), "fs2_draw_noise")
        return self.noise if out is None else out

    def _noise_ptr(self, noise):
        if noise is None:
            return None
        torch = self._torch
        if not torch.is_tensor(noise):
            noise = torch.as_tensor(np.ascontiguousarray(noise, dtype=np.float64), device="cuda:%d" % self.device)
        assert noise.dtype == torch.float64 and noise.is_cuda and noise.numel() == self.P and noise.is_contiguous()
        self._keep = noise
        return C.c_void_p(noise.data_ptr())

    def motion(self, rotation: float, translation: float, noise=None):
        check(self._L.fs2_motion(self._h, float(rotation), float(translation), self._noise_ptr(noise), self._stream()), "fs2_motion")

    def _assoc_buf(self, M, want):
        if not want or M == 0:
            return None
        return self._torch.full((M, self.P), -9, dtype=self._torch.int32, device="cuda:%d" % self.device)

    def update(self, obs, want_assoc: bool = False):
        obs, M = self._obs(obs)
        a = self._assoc_buf(M, want_assoc)
        check(self._L.fs2_update(self._h, _pd(obs), M, C.c_void_p(a.data_ptr()) if a is not None else None, self._stream()), "fs2_update")
        return a

    def motion_update(self, rotation, translation, obs, noise=None, want_assoc: bool = False):
        obs, M = self._obs(obs)
        a = self._assoc_buf(M, want_assoc)
        check(self._L.fs2_motion_update(self._h, float(rotation), float(translation), self._noise_ptr(noise), _pd(obs), M,
                                        C.c_void_p(a.data_ptr()) if a is not None else None, self._stream()), "fs2_motion_update")
        return a

    def weight_total(self):
        check(self._L.fs2_weight_total(self._h, self._stream()), "fs2_weight_total")

    def normalize(self, total=None):
        """total: optional 1-element float64 CUDA tensor (the all-gathered total when sharded)."""
        ptr = None
        if total is not None:
            assert total.is_cuda and total.dtype == self._torch.float64
            self._keep_total = total
            ptr = C.c_void_p(total.data_ptr())
        check(self._L.fs2_normalize(self._h, ptr, self._stream()), "fs2_normalize")

    def finish_step(self, u0: float, ancestor=None):
        """Normalise, Neff, resample if the device decides so, estimate -- one asynchronous chain (fs2_finish_step).
        Read ``self.stats`` afterwards (STAT_RESAMPLED, STAT_NEFF, STAT_EST_*)."""
        ptr = C.c_void_p(ancestor.data_ptr()) if ancestor is not None else None
        check(self._L.fs2_finish_step(self._h, float(u0), ptr, self._stream()), "fs2_finish_step")

    def estimate(self):
        check(self._L.fs2_estimate(self._h, self._stream()), "fs2_estimate")

    def known_landmarks(self, eps: float = 0.5, frac: float = 0.7, min_samples: int = 0, max_clusters: int = 4096):
        """LandmarkUtils.update_known_landmarks (landmark_utils.py:120-144) on the maps in device memory.
        Returns (centroids [K][2], members [K], info) or None when the reference returns early."""
        cent = np.zeros((max_clusters, 2))
        mem = np.zeros(max_clusters, np.int64)
        k = C.c_int32(0)
        info = _lib.Fs2KlInfo()
        check(self._L.fs2_known_landmarks(self._h, float(eps), float(frac), int(min_samples), int(max_clusters), _pd(cent),
                                          mem.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(k), C.byref(info), self._stream()),
              "fs2_known_landmarks")
        if k.value < 0:
            return None
        return cent[:k.value].copy(), mem[:k.value].copy(), {f: getattr(info, f) for f, _ in info._fields_}

    def resample_indices(self, u0: float, w_all=None, m_begin: int = 0, m_count: int | None = None, out=None):
        torch = self._torch
        w_all = self.w if w_all is None else w_all
        n = w_all.numel()
        m_count = self.P if m_count is None else m_count
        if out is None:
            out = torch.empty(m_count, dtype=torch.int32, device="cuda:%d" % self.device)
        check(self._L.fs2_resample_indices(self._h, C.c_void_p(w_all.data_ptr()), n, float(u0), int(m_begin), int(m_count),
                                           C.c_void_p(out.data_ptr()), self._stream()), "fs2_resample_indices")
        return out

    def gather(self, ancestor):
        assert ancestor.dtype == self._torch.int32 and ancestor.numel() == self.P
        check(self._L.fs2_gather(self._h, C.c_void_p(ancestor.data_ptr()), self._stream()), "fs2_gather")

    # -- whole step (FastSLAM2.iterate) ---------------------------------------------------------------
    def step(self, rotation, translation, obs, noise=None, u0=0.0, step_index: int = 0, want_assoc: bool = True,
             want_ancestor: bool = True):
        """One filter step through fs2_step_host.  noise: host array [P] (reference-stream mode) or None
        (device generator).  Returns a dict shaped like the oracle's / the reference harness'."""
        torch = self._torch
        obs, M = self._obs(obs)
        a = self._assoc_buf(M, want_assoc)
        anc = torch.empty(self.P, dtype=torch.int32, device="cuda:%d" % self.device) if want_ancestor else None
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
        res = Fs2StepResult()
        u0 = 0.0 if u0 is None or (isinstance(u0, float) and np.isnan(u0)) else float(u0)
        check(self._L.fs2_step_host(self._h, float(rotation), float(translation), _pd(obs), M, _pd(nz), int(step_index), u0,
                                    C.c_void_p(a.data_ptr()) if a is not None else None,
                                    C.c_void_p(anc.data_ptr()) if anc is not None else None, C.byref(res), self._stream()),
              "fs2_step_host")
        out = dict(estimate=np.array([res.x, res.y, res.yaw]), neff=res.neff, total=res.total, resampled=bool(res.resampled))
        if want_assoc:
            out["assoc"] = a.cpu().numpy() if a is not None else np.zeros((0, self.P), np.int32)
        if want_ancestor:
            out["resample_idx"] = anc.cpu().numpy() if res.resampled else np.arange(self.P, dtype=np.int32)
        return out
