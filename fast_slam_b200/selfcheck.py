"""Sharded filter against ONE GPU holding all particles: same observations, same counter-based motion noise, same
resampling start points -- every resampling decision, the global estimate and every particle (pose, weight, map) must
agree.  Used by tests/sharded_check.py and, after the timed region, by bench.py at N > 1.
Needs torch.distributed (nccl) initialised; rank 0 also builds the single-GPU filter."""
from __future__ import annotations

import numpy as np


def sharded_equals_single(P: int, L: int, lcap: int, M: int, seed: int, steps: int, novel_every: int = 5, rtol: float = 1e-12):
    """Returns dict(resamples, migrated, mode) -- `migrated` = maps that crossed GPUs over the run, summed over ranks.
    Raises AssertionError (on rank 0) at the first difference; the other ranks are told and raise too."""
    import torch
    import torch.distributed as dist
    from . import DeviceFilter
    from .dist import ShardedFilter
    from .filter import _hash_uniform
    from .synthetic import fill_synthetic_device, synthetic_obs, synthetic_odometry
    rank, world = dist.get_rank(), dist.get_world_size()
    N = P * world
    sh = ShardedFilter(P, lcap, seed=seed)
    world_pts = fill_synthetic_device(sh.store, L, seed)
    single = None
    if rank == 0:
        single = DeviceFilter(N, lcap, seed=seed)
        fill_synthetic_device(single, L, seed)
    nres = moved = 0
    err = None
    for s in range(steps):
        rot, tr = synthetic_odometry(s)
        obs = synthetic_obs(seed, s, world_pts, M, novel=2 if s % novel_every == novel_every - 1 else 0, max_range=9.0)
        u0 = _hash_uniform(seed, s) / N
        res = sh.step(rot, tr, obs, u0, s)
        nres += int(res)
        if res:
            if sh.placed:
                moved += int(sh.migrated[1])                     # maps this rank pulled out of other GPUs' stores
            else:
                mine = sh._anc_all[rank * P:(rank + 1) * P]
                moved += int((torch.div(mine, P, rounding_mode="floor") != rank).sum().item())
        st = sh.store.download()
        st["ids"] = sh.logical_order()
        gathered = [None] * world if rank == 0 else None
        dist.gather_object({k: st[k] for k in ("x", "y", "yaw", "w", "counts", "lm", "ids")}, gathered, dst=0)
        if rank == 0:
            try:
                r1 = single.step(rot, tr, obs, noise=None, u0=u0, step_index=s, want_assoc=False, want_ancestor=False)
                assert r1["resampled"] == res, ("resample decision", s, r1["neff"], sh.last["neff"])
                assert abs(r1["neff"] - sh.last["neff"]) <= 1e-9 * r1["neff"], ("neff", s)
                assert np.allclose(r1["estimate"], sh.last["estimate"], rtol=rtol, atol=0), ("estimate", s)
                ref = single.download()
                ids = np.concatenate([g["ids"] for g in gathered])
                assert np.array_equal(np.sort(ids), np.arange(N)), "logical ids are not a permutation"
                for g in gathered:
                    assert (np.diff(g["ids"]) > 0).all(), "a shard's local order is not the logical order"
                order = np.argsort(ids)                          # physical (rank, local) -> logical
                for k in ("x", "y", "yaw", "w", "counts"):
                    got = np.concatenate([g[k] for g in gathered])[order]
                    if k == "counts":
                        assert np.array_equal(got, ref[k]), (s, k)
                    else:
                        assert np.allclose(got, ref[k], rtol=rtol, atol=0), (s, k)
                got = np.concatenate([g["lm"] for g in gathered])[order]
                mask = np.arange(lcap)[None, :] < ref["counts"][:, None]
                assert np.allclose(got[mask], ref["lm"][mask], rtol=rtol, atol=0), (s, "lm")
            except AssertionError as e:                          # tell the others before raising: they sit in a collective
                err = "step %d: %r" % (s, e.args)
        flag = torch.tensor([1 if err else 0], device="cuda")
        dist.broadcast(flag, 0)
        if int(flag.item()):
            raise AssertionError(err or "rank 0 found a difference between the sharded and the single-GPU filter")
    tot = torch.tensor([moved], device="cuda", dtype=torch.int64)
    dist.all_reduce(tot)
    out = dict(resamples=nres, migrated=int(tot.item()), mode=sh.mode, p2p=bool(sh.p2p), fallbacks=sh.fallbacks,
               sharded=sh, single=single, world_points=world_pts)
    return out
