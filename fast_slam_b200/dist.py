"""Particles sharded over the GPUs of one node (north_star "partitioning").

One process per GPU (torchrun), rank r owns the contiguous global block [r*P, (r+1)*P).  Motion, association,
EKF and weighting need no communication.  Per step:

  1. all_gather of the per-shard weight totals (8 B per rank)      -> every rank normalises with the same total
  2. all_gather of the per-shard stats block (64 B per rank)       -> global Neff, global first arg-max + its pose
  3. only when Neff < N/2 (fast_slam_2.py:62):
       all_gather of the normalised weights (8 B per particle)     -> every rank runs the SAME exact-rounding scan
                                                                      (fs2_resample_indices), so the global
                                                                      systematic resample equals the 1-GPU result
                                                                      index for index
       all_to_all of the surviving particles a rank needs from the others (pose + weight + count + map,
       packed by fs2_pack_records, NCCL P2P over NVLink), then the local copy-on-resample gather takes its
       ancestors from the local store or from the received records (fs2_gather_ext).

Because systematic resampling keeps order, ancestors are non-decreasing in the slot index: the particles a
rank needs from rank r are a sorted run, and everything about who-sends-what follows from the global ancestor
array that every rank computes identically -- no request messages are needed.

The planning functions below are plain tensor code (CPU or CUDA) and are unit-tested with the gloo backend.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


def migration_plan(anc_all, P: int, world: int, rank: int):
    """Who sends what, from the global ancestor array (int tensor [P*world], non-decreasing).

    Returns (send_ids, recv_ids, local_anc):
      send_ids[d]  sorted unique GLOBAL ids owned by `rank` that rank d needs (empty for d == rank)
      recv_ids[r]  sorted unique GLOBAL ids owned by rank r that `rank` needs (empty for r == rank)
      local_anc    int32 [P]: ancestor of each of my slots as a LOCAL index (< P), or P + position in the
                   staging order = concatenation of recv_ids[0], recv_ids[1], ... (skipping `rank`)
    """
    import torch
    lo, hi = rank * P, (rank + 1) * P
    send_ids = []
    for d in range(world):
        if d == rank:
            send_ids.append(anc_all.new_empty(0))
            continue
        a = anc_all[d * P:(d + 1) * P]
        a = a[(a >= lo) & (a < hi)]
        send_ids.append(torch.unique_consecutive(a))
    mine = anc_all[lo:hi]
    owner = torch.div(mine, P, rounding_mode="floor")
    local_anc = (mine - lo).to(torch.int32)
    recv_ids = []
    off = 0
    for r in range(world):
        if r == rank:
            recv_ids.append(anc_all.new_empty(0))
            continue
        sel = owner == r
        ids = torch.unique_consecutive(mine[sel])
        recv_ids.append(ids)
        if ids.numel():
            pos = torch.searchsorted(ids, mine[sel])
            local_anc[sel] = (P + off + pos).to(torch.int32)
        off += int(ids.numel())
    return send_ids, recv_ids, local_anc


def combine_stats(stats_all, P: int, world: int):
    """Global weight statistics from the all-gathered per-shard stats blocks [world][FS2_STATS_LEN].
    Sum of squares in rank order; first arg-max = largest weight, ties to the lowest LOGICAL (global) index
    (fast_slam_2.py:208) -- every shard reports the logical id of its own first arg-max (STAT_ARGMAX_ID).
    Returns dict(neff, argmax_global, estimate[3], sumsq)."""
    n = P * world
    sumsq = 0.0
    for r in range(world):
        sumsq = sumsq + float(stats_all[r][_lib.STAT_SUMSQ])
    neff = float(n) if sumsq < 1.0 / n else 1.0 / sumsq                     # fast_slam_2.py:220-223
    best = 0
    for r in range(1, world):
        wr, wb = float(stats_all[r][_lib.STAT_WMAX]), float(stats_all[best][_lib.STAT_WMAX])
        if wr > wb or (wr == wb and float(stats_all[r][_lib.STAT_ARGMAX_ID]) < float(stats_all[best][_lib.STAT_ARGMAX_ID])):
            best = r
    s = stats_all[best]
    return dict(neff=neff, sumsq=sumsq, argmax_global=int(s[_lib.STAT_ARGMAX_ID]),
                estimate=np.array([float(s[_lib.STAT_EST_X]), float(s[_lib.STAT_EST_Y]), float(s[_lib.STAT_EST_YAW])]))


def gather_rows(local, n_local: int, world: int, dev):
    """all-gather of a different number of rows per rank: the counts first, then the rows padded to the largest
    count.  Returns (rows of all ranks in rank order, counts)."""
    import torch
    import torch.distributed as dist
    cnt = torch.tensor([n_local], dtype=torch.int64, device=dev)
    cnts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnts, cnt)
    cnts = [int(c) for c in cnts.cpu()]
    m = max(cnts)
    if m == 0:
        return local[:0], cnts
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=dev)
    pad[:n_local] = local[:n_local]
    allp = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=dev)
    dist.all_gather_into_tensor(allp, pad)
    return torch.cat([allp[r * m:r * m + cnts[r]] for r in range(world)]), cnts


class ShardedFilter:
    """FastSLAM2.iterate over world_size GPUs; construct after torch.distributed is initialised (nccl)."""

    def __init__(self, particles_per_gpu: int, landmark_capacity: int, seed: int = 0, **cfg):
        import torch
        import torch.distributed as dist
        from .store import DeviceFilter
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.P = int(particles_per_gpu)
        self.N = self.P * self.world
        import os
        # placed: offspring stay on their ancestor's GPU, only the overflow migrates (csrc/fs2_place.cuh) -- the default;
        # p2p: logical slot m lives on GPU m // P, peer gather; pull: peer pull + staged records; nccl: all_to_all
        self.mode = os.environ.get("FS2_DIST", "placed")
        # spare map slots for the peer gather: a particle that survives only on another GPU keeps its slot for the
        # round, so in the worst case (every local particle needed remotely, none locally) P spare slots are needed --
        # the memory a double-buffered gather would take anyway.  FS2_SPARE_FRAC trades memory for staged fallbacks.
        frac = float(os.environ.get("FS2_SPARE_FRAC", "1.0"))
        spare = int(self.P * frac) if (self.mode in ("p2p", "placed") and self.world > 1) else 0
        self.store = DeviceFilter(self.P, landmark_capacity, seed=seed, global_particles=self.N,
                                  global_offset=self.rank * self.P, spare_slots=spare, **cfg)
        dev = self.store.x.device
        self.dev = dev
        self._tot_all = torch.empty(self.world, dtype=torch.float64, device=dev)
        self._stats_all = torch.empty((self.world, _lib.FS2_STATS_LEN), dtype=torch.float64, device=dev)
        self._w_all = torch.empty(self.N, dtype=torch.float64, device=dev)
        self._anc_all = torch.empty(self.N, dtype=torch.int32, device=dev)
        self.rstride = 8 + 6 * self.store.lcap
        self.last = None
        self._bufs = {}
        self._barrier = torch.zeros(1, dtype=torch.int32, device=dev)
        self.p2p = False
        if self.mode in ("p2p", "pull", "placed") and self.world <= 16:
            self.p2p = self._open_peers()
        self.fallbacks = 0
        self.placed = False
        self.migrated_total = [0, 0]                         # offspring that changed GPU (whole job), maps this rank pulled
        if self.mode == "placed":
            if self.p2p:
                check(self.store._L.fs2_place_enable(self.store._h, self.store._stream()), "fs2_place_enable")
                from .store import _DevArray
                ids = self.store._L.fs2_place_logical_ids(self.store._h)
                self.logical_ids = torch.as_tensor(_DevArray(ids, (self.P,), "<i4", self.store), device=dev)
                self._place = torch.arange(self.N, dtype=torch.int32, device=dev)     # logical particle -> rank * P + local index
                self._place_new = torch.empty_like(self._place)
                self.placed = True
            else:
                self.mode = "nccl"                           # no peer access between these GPUs: staged exchange instead

    def _open_peers(self) -> bool:
        """Map every shard's store into this process (CUDA IPC).  All ranks must agree, so the outcome is reduced."""
        torch, dist, st = self.torch, self.dist, self.store
        mine = (C.c_ubyte * 448)()
        ok = st._L.fs2_ipc_export(st._h, C.byref(mine)) == 0
        t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=self.dev)
        allh = torch.empty(448 * self.world, dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(allh, t)
        if ok:
            buf = (C.c_ubyte * (448 * self.world)).from_buffer_copy(bytes(allh.cpu().numpy().tobytes()))
            ok = st._L.fs2_ipc_open_peers(st._h, C.byref(buf), self.world, self.rank) == 0
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    # ------------------------------------------------------------------------------------------------
    def _global_stats(self):
        st = self.store
        self.dist.all_gather_into_tensor(self._stats_all.view(-1), st.stats)
        return combine_stats(self._stats_all.cpu().numpy(), self.P, self.world)      # one host sync

    def step(self, rotation, translation, obs, u0, step_index, events=None):
        torch, dist, st = self.torch, self.dist, self.store
        sigma = st.cfg.rotation_noise if rotation != 0 else st.cfg.translation_noise
        st.draw_noise(sigma, step_index)
        if events is not None:
            events[0].record()
        st.motion_update(rotation, translation, obs)
        if events is not None:
            events[1].record()
        st.weight_total()
        dist.all_gather_into_tensor(self._tot_all, st.stats[_lib.STAT_TOTAL:_lib.STAT_TOTAL + 1])
        total = self._tot_all.sum().reshape(1)            # identical on every rank (same inputs, same kernel)
        st.normalize(total)
        g = self._global_stats()
        resampled = g["neff"] < self.N / 2                # fast_slam_2.py:62
        if resampled:
            self.resample(u0)
            st.estimate()
            g2 = self._global_stats()
            g["estimate"], g["argmax_global"] = g2["estimate"], g2["argmax_global"]
        g["resampled"] = resampled
        self.last = g
        return resampled

    def logical_order(self):
        """Logical (reference-order) id of every local particle, as a host array; increasing.  Rank r holds the ids
        r * P .. (r + 1) * P - 1 until the first resample of a placed filter, any P of them afterwards."""
        if self.placed:
            return self.logical_ids.cpu().numpy().astype(np.int64)
        return np.arange(self.rank * self.P, (self.rank + 1) * self.P, dtype=np.int64)

    def _gather_rows(self, local, n_local: int):
        return gather_rows(local, n_local, self.world, self.dev)

    def known_landmarks(self, eps: float = 0.5, frac: float = 0.7, max_clusters: int = 4096):
        """LandmarkUtils.update_known_landmarks (landmark_utils.py:120-144) over all shards: the result of clustering
        the unsharded filter's maps, identical on every rank.  Every shard counts its own maps into its grid, the
        occupied tiles (a few MB) are all-gathered and merged, the cell-level clustering runs replicated; only the
        points near a decision boundary travel.  Returns (centroids, members, info) or None (min_samples < 1)."""
        torch, dist, st = self.torch, self.dist, self.store
        L, h = st._L, st._h
        n_local = C.c_int64(0)
        check(L.fs2_kl_shard_begin(h, float(eps), C.byref(n_local), st._stream()), "fs2_kl_shard_begin")
        mine = torch.tensor([n_local.value], dtype=torch.int64, device=self.dev)
        alln = torch.empty(self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(alln, mine)
        alln = [int(v) for v in alln.cpu()]
        total = sum(alln)
        min_samples = int(total / self.N * frac)                          # landmark_utils.py:129-130
        if min_samples < 1:
            return None
        n_tiles = C.c_int32(0)
        index_offset = sum(alln[:self.rank])
        if self.placed:
            # a point's global index follows the LOGICAL particle order: exclusive prefix of the map lengths in that order
            cnt_all = torch.empty(self.N, dtype=torch.int32, device=self.dev)
            dist.all_gather_into_tensor(cnt_all, st.count)
            cnt_log = torch.index_select(cnt_all, 0, self._place).to(torch.int64)
            cum = torch.cumsum(cnt_log, 0) - cnt_log
            bases = torch.index_select(cum, 0, self.logical_ids).contiguous()
            check(L.fs2_kl_shard_set_bases(h, C.c_void_p(bases.data_ptr()), st._stream()), "fs2_kl_shard_set_bases")
            index_offset = 0
        check(L.fs2_kl_shard_count(h, index_offset, C.byref(n_tiles), st._stream()), "fs2_kl_shard_count")
        words = L.fs2_kl_record_bytes() // 8
        rec = torch.empty((max(n_tiles.value, 1), words), dtype=torch.int64, device=self.dev)
        check(L.fs2_kl_shard_export(h, C.c_void_p(rec.data_ptr()), n_tiles.value, st._stream()), "fs2_kl_shard_export")
        allrec, _ = self._gather_rows(rec, n_tiles.value)
        allrec = allrec.contiguous()
        involved = C.c_int64(0)
        check(L.fs2_kl_shard_merge(h, C.c_void_p(allrec.data_ptr()) if allrec.numel() else None, int(allrec.shape[0]),
                                   min_samples, C.byref(involved), st._stream()), "fs2_kl_shard_merge")
        pts_all = torch.empty((0, 3), dtype=torch.float64, device=self.dev)
        if involved.value:
            cap = 1 << 16
            while True:
                pts = torch.empty((cap, 3), dtype=torch.float64, device=self.dev)
                n = C.c_int64(0)
                check(L.fs2_kl_shard_extract(h, C.c_void_p(pts.data_ptr()), cap, C.byref(n), st._stream()), "fs2_kl_shard_extract")
                if n.value <= cap:
                    break
                cap = int(n.value)
            pts_all, _ = self._gather_rows(pts, int(n.value))
            pts_all = pts_all.contiguous()
        cent = np.zeros((max_clusters, 2))
        mem = np.zeros(max_clusters, np.int64)
        k = C.c_int32(0)
        info = _lib.Fs2KlInfo()
        check(L.fs2_kl_shard_finish(h, C.c_void_p(pts_all.data_ptr()) if pts_all.numel() else None, int(pts_all.shape[0]), total,
                                    max_clusters, cent.ctypes.data_as(C.POINTER(C.c_double)), mem.ctypes.data_as(C.POINTER(C.c_int64)),
                                    C.byref(k), C.byref(info), st._stream()), "fs2_kl_shard_finish")
        return cent[:k.value].copy(), mem[:k.value].copy(), {f: getattr(info, f) for f, _ in info._fields_}

    def _buffer(self, name: str, rows: int):
        """Persistent, geometrically grown staging buffers (a fresh multi-GB cudaMalloc per resample costs ms)."""
        torch = self.torch
        buf = self._bufs.get(name)
        if buf is None or buf.shape[0] < rows:
            cap = max(rows, int(1.5 * (buf.shape[0] if buf is not None else 0)), 1024)
            buf = torch.empty((cap, self.rstride), dtype=torch.float64, device=self.dev)
            self._bufs[name] = buf
        return buf[:rows]

    def resample(self, u0: float):
        """Global systematic resample + migration of the survivors' maps (fast_slam_2.py:177-199)."""
        torch, dist, st = self.torch, self.dist, self.store
        import os, time
        prof = os.environ.get("FS2_DIST_PROFILE")
        def tick(tag):
            if prof:
                torch.cuda.synchronize()
                self._prof.append((tag, time.perf_counter()))
        self._prof = []
        tick("start")
        dist.all_gather_into_tensor(self._w_all, st.w)
        tick("allgather_w")
        if self.placed:
            # the reference's running sum walks the particles in LOGICAL order
            w_log = torch.index_select(self._w_all, 0, self._place)
            st.resample_indices(u0, w_all=w_log, m_begin=0, m_count=self.N, out=self._anc_all)     # logical ancestors
            tick("scan")
            info = (C.c_int64 * 2)()
            check(st._L.fs2_place_resample(st._h, C.c_void_p(self._anc_all.data_ptr()), C.c_void_p(self._place.data_ptr()),
                                           C.c_void_p(self._place_new.data_ptr()), info, st._stream()), "fs2_place_resample")
            tick("plan_gather")
            dist.all_reduce(self._barrier)                # nobody publishes before everybody has finished reading
            check(st._L.fs2_place_commit(st._h, st._stream()), "fs2_place_commit")
            self._place, self._place_new = self._place_new, self._place
            tick("commit")
            self.migrated = (int(info[0]), int(info[1]))
            self.migrated_total[0] += int(info[0]); self.migrated_total[1] += int(info[1])
            if prof:
                print("rank %d resample ms:" % self.rank,
                      " ".join("%s=%.2f" % (a, 1e3 * (b - c)) for (a, b), (_, c) in zip(self._prof[1:], self._prof[:-1])),
                      "moved_offspring_all_ranks=%d maps_pulled_here=%d" % (info[0], info[1]), flush=True)
            return None
        st.resample_indices(u0, w_all=self._w_all, m_begin=0, m_count=self.N, out=self._anc_all)
        tick("scan")
        if self.p2p and self.mode == "p2p":
            # one set of kernels: local copy-on-resample copies and NVLink pulls of remote ancestors together
            rc = st._L.fs2_gather_p2p(st._h, C.c_void_p(self._anc_all.data_ptr()), st._stream())
            if rc == 0:
                tick("gather_p2p")
                dist.all_reduce(self._barrier)            # nobody publishes before everybody has finished reading
                check(st._L.fs2_gather_commit(st._h, st._stream()), "fs2_gather_commit")
                tick("commit")
                if prof:
                    mine = self._anc_all[self.rank * self.P:(self.rank + 1) * self.P]
                    remote = mine[(mine // self.P) != self.rank]
                    local = mine[(mine // self.P) == self.rank]
                    self.migrated = (0, int(remote.numel()))
                    print("rank %d resample ms:" % self.rank,
                          " ".join("%s=%.2f" % (a, 1e3 * (b - c)) for (a, b), (_, c) in zip(self._prof[1:], self._prof[:-1])),
                          "remote=%d unique_remote=%d local_unique=%d" % (remote.numel(), torch.unique(remote).numel(), torch.unique(local).numel()),
                          flush=True)
                return None
            if rc != -3:
                check(rc, "fs2_gather_p2p")
            self.fallbacks += 1                           # not enough free slots this round: pull into records instead
        send_ids, recv_ids, local_anc = migration_plan(self._anc_all, self.P, self.world, self.rank)
        n_send = [int(t.numel()) for t in send_ids]       # host sync: split sizes of the all_to_all
        n_recv = [int(t.numel()) for t in recv_ids]
        tick("plan")
        recv = self._buffer("recv", sum(n_recv))
        if self.p2p:
            # pull what I need straight out of the owners' stores (peer loads over NVLink, my own kernel) ...
            off = 0
            keep = []
            for r in range(self.world):
                if n_recv[r]:
                    ids = recv_ids[r].to(torch.int64).contiguous()
                    keep.append(ids)
                    check(st._L.fs2_pull_records(st._h, r, C.c_void_p(ids.data_ptr()), n_recv[r],
                                                 C.c_void_p(recv[off:].data_ptr()), st._stream()), "fs2_pull_records")
                    off += n_recv[r]
            tick("pull")
            # ... and nobody rewrites its store before everybody has finished reading (stream-ordered barrier)
            dist.all_reduce(self._barrier)
            tick("barrier")
            send, sel = None, keep
        else:
            sel = torch.cat(send_ids).to(torch.int64) - self.rank * self.P
            send = self._buffer("send", int(sel.numel()))
            if sel.numel():
                check(st._L.fs2_pack_records(st._h, C.c_void_p(sel.data_ptr()), int(sel.numel()), C.c_void_p(send.data_ptr()),
                                             st._stream()), "fs2_pack_records")
            tick("pack")
            dist.all_to_all_single(recv, send, output_split_sizes=n_recv, input_split_sizes=n_send)
            tick("all_to_all")
        check(st._L.fs2_gather_ext(st._h, C.c_void_p(local_anc.data_ptr()), C.c_void_p(recv.data_ptr()) if recv.numel() else None,
                                   int(recv.shape[0]), st._stream()), "fs2_gather_ext")
        tick("gather")
        if prof and self.rank == 0:
            t0 = self._prof[0][1]
            print("resample breakdown ms:", " ".join("%s=%.2f" % (a, 1e3 * (b - c)) for (a, b), (_, c) in zip(self._prof[1:], self._prof[:-1])),
                  "sent=%d recv=%d" % (sum(n_send), sum(n_recv)), flush=True)
        self.migrated = (sum(n_send), sum(n_recv))
        self._keep = (send, recv, local_anc, sel)         # alive until the stream has consumed them
        return local_anc
