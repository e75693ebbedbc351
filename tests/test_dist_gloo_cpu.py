"""CPU, world_size 2 and 3 over gloo: the sharded resample's host logic (fast_slam_b200/dist.py).  Every
rank computes the global ancestors, derives who-sends-what without any request message, exchanges the
packed particles with all_to_all, and gathers from (local store + received records); the result must be
the single-process resample, particle for particle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fast_slam_b200.dist import combine_stats, migration_plan
from oracle import fs2_oracle as fo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_global(P, world, lcap, seed, skew):
    rng = np.random.default_rng(seed)
    N = P * world
    w = rng.uniform(0, 1, N) ** skew
    if seed % 2:
        w[: N // 2] *= 1e-6          # almost all weight on the upper shards: heavy migration
    w /= w.sum()
    count = rng.integers(0, lcap + 1, N).astype(np.int32)
    lm = rng.normal(0, 3, (N, lcap, 6))
    pose = rng.normal(0, 1, (N, 3))
    return w, count, lm, pose


def _worker(rank, world, port, P, lcap, seed, skew, u0, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, count, lm, pose = _make_global(P, world, lcap, seed, skew)
        lo, hi = rank * P, (rank + 1) * P
        # --- what ShardedFilter.resample does, with numpy standing in for the device kernels ---
        w_local = torch.as_tensor(w[lo:hi].copy())
        w_all = torch.empty(P * world, dtype=torch.float64)
        dist.all_gather_into_tensor(w_all, w_local)
        assert np.array_equal(w_all.numpy(), w)
        anc, _ = fo.resample_indices(w_all.numpy(), u0)                    # the device scan is bit-identical to this
        anc_all = torch.as_tensor(anc.astype(np.int64))
        send_ids, recv_ids, local_anc = migration_plan(anc_all, P, world, rank)
        rstride = 8 + 6 * lcap
        sel = torch.cat(send_ids).numpy() - lo
        send = np.zeros((len(sel), rstride))
        for r, p in enumerate(sel):                                        # fs2_pack_records
            g = lo + p
            send[r, 0:3] = pose[g]; send[r, 3] = w[g]; send[r, 4] = count[g]
            send[r, 8:8 + 6 * count[g]] = lm[g, :count[g]].ravel()
        n_send = [int(t.numel()) for t in send_ids]
        n_recv = [int(t.numel()) for t in recv_ids]
        recv = torch.zeros((sum(n_recv), rstride), dtype=torch.float64)
        dist.all_to_all_single(recv, torch.as_tensor(send), output_split_sizes=n_recv, input_split_sizes=n_send)
        recv = recv.numpy()
        new_pose = np.zeros((P, 3)); new_w = np.zeros(P); new_count = np.zeros(P, np.int32); new_lm = np.zeros((P, lcap, 6))
        for m in range(P):                                                 # fs2_gather_ext
            a = int(local_anc[m])
            if a < P:
                g = lo + a
                new_pose[m], new_w[m], new_count[m] = pose[g], w[g], count[g]
                new_lm[m, :count[g]] = lm[g, :count[g]]
            else:
                r = recv[a - P]
                n = int(r[4])
                new_pose[m], new_w[m], new_count[m] = r[0:3], r[3], n
                new_lm[m, :n] = r[8:8 + 6 * n].reshape(n, 6)
        # --- single-process truth ---
        for m in range(P):
            g = anc[lo + m]
            assert np.array_equal(new_pose[m], pose[g]) and new_w[m] == w[g] and new_count[m] == count[g], (rank, m)
            assert np.array_equal(new_lm[m, :count[g]], lm[g, :count[g]])
        # nothing is sent twice, and only what is needed
        for d in range(world):
            assert len(np.unique(send_ids[d].numpy())) == send_ids[d].numel()
        out[rank] = (sum(n_send), sum(n_recv))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,P,seed,skew", [(2, 300, 0, 4), (2, 257, 1, 10), (3, 128, 2, 6), (3, 100, 3, 1)])
def test_sharded_resample_equals_single_process(world, P, seed, skew):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    u0 = float(np.random.default_rng(seed).uniform(0, 1.0 / (P * world)))
    mp.spawn(_worker, args=(world, port, P, 6, seed, skew, u0, out), nprocs=world, join=True)
    sent = sum(v[0] for v in out.values()); got = sum(v[1] for v in out.values())
    assert sent == got
    if seed % 2:
        assert sent > 0            # the skewed case really migrates particles


def test_migration_plan_is_consistent_between_ranks():
    rng = np.random.default_rng(5)
    world, P = 4, 64
    anc = torch.as_tensor(np.sort(rng.integers(0, world * P, world * P)))
    plans = [migration_plan(anc, P, world, r) for r in range(world)]
    for r in range(world):
        for d in range(world):
            assert torch.equal(plans[r][0][d], plans[d][1][r])       # what r sends to d is what d expects from r
        la = plans[r][2]
        assert la.dtype == torch.int32 and la.numel() == P
        staged = sum(int(t.numel()) for t in plans[r][1])
        assert int(la.max()) < P + staged


def test_combine_stats_first_argmax_and_neff():
    from fast_slam_b200 import _lib
    P, world = 10, 3
    s = np.zeros((world, _lib.FS2_STATS_LEN))
    s[:, 1] = [0.01, 0.02, 0.03]; s[:, 3] = [0.2, 0.5, 0.5]; s[:, 4] = [3, 7, 1]
    s[:, 5] = [1, 2, 3]
    s[:, _lib.STAT_ARGMAX_ID] = [3, 17, 21]                          # contiguous shards: rank * P + local index
    g = combine_stats(s, P, world)
    assert g["argmax_global"] == 17 and g["estimate"][0] == 2.0     # tie between ranks 1 and 2: lower global index
    assert g["neff"] == pytest.approx(1 / 0.06)
    # freely placed shards: the tie goes to the lower LOGICAL id, whichever rank holds it
    s[:, _lib.STAT_ARGMAX_ID] = [3, 25, 12]
    g = combine_stats(s, P, world)
    assert g["argmax_global"] == 12 and g["estimate"][0] == 3.0
    s[:, 1] = 0.001
    assert combine_stats(s, P, world)["neff"] == 30.0                # sum w^2 < 1/N -> N (fast_slam_2.py:220)


def _gather_rows_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fast_slam_b200.dist import gather_rows
        # rank r contributes r * 2 rows (rank 0: none) of width 3, values identify (rank, row)
        n = rank * 2
        local = torch.zeros((max(n, 1) + 3, 3), dtype=torch.float64)       # over-allocated, only n rows count
        for i in range(n):
            local[i] = torch.tensor([rank, i, 100.0 * rank + i])
        rows, cnts = gather_rows(local, n, world, torch.device("cpu"))
        empty, c0 = gather_rows(local, 0, world, torch.device("cpu"))
        out[rank] = (rows.numpy().copy(), cnts, int(empty.shape[0]), c0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_rows_of_different_lengths(world):
    """the ragged all-gather behind ShardedFilter.known_landmarks (tiles and involved points differ per rank)"""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gather_rows_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    want = np.array([[r, i, 100.0 * r + i] for r in range(world) for i in range(2 * r)], dtype=np.float64).reshape(-1, 3)
    for r in range(world):
        rows, cnts, nempty, c0 = out[r]
        assert cnts == [2 * q for q in range(world)] and np.array_equal(rows, want)
        assert nempty == 0 and c0 == [0] * world
