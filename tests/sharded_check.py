"""Run under torchrun on >= 2 GPUs (tests/test_gpu_sharded.py): the sharded filter against one GPU holding
all particles, same observations, same counter-based noise, same resampling start points."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_slam_b200 import DeviceFilter, _lib          # noqa: E402
from fast_slam_b200.dist import ShardedFilter          # noqa: E402
from fast_slam_b200.filter import _hash_uniform        # noqa: E402
from fast_slam_b200.synthetic import fill_synthetic_device, synthetic_obs, synthetic_odometry  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    P, L, lcap, M, seed, steps = 1 << 14, 64, 96, 16, 99, 25
    N = P * world
    sh = ShardedFilter(P, lcap, seed=seed)
    world_pts = fill_synthetic_device(sh.store, L, seed)
    single = None
    if rank == 0:
        single = DeviceFilter(N, lcap, seed=seed)
        fill_synthetic_device(single, L, seed)
    nres = moved = 0
    for s in range(steps):
        rot, tr = synthetic_odometry(s)
        obs = synthetic_obs(seed, s, world_pts, M, novel=2 if s % 5 == 4 else 0, max_range=9.0)
        u0 = _hash_uniform(seed, s) / N
        res = sh.step(rot, tr, obs, u0, s)
        nres += int(res)
        if res:
            mine = sh._anc_all[rank * P:(rank + 1) * P]
            moved += int((torch.div(mine, P, rounding_mode="floor") != rank).sum().item())
        st = sh.store.download()
        gathered = [None] * world if rank == 0 else None
        dist.gather_object({k: st[k] for k in ("x", "y", "yaw", "w", "counts", "lm")}, gathered, dst=0)
        if rank == 0:
            r1 = single.step(rot, tr, obs, noise=None, u0=u0, step_index=s, want_assoc=False, want_ancestor=False)
            assert r1["resampled"] == res, (s, r1["neff"], sh.last["neff"])
            assert abs(r1["neff"] - sh.last["neff"]) <= 1e-9 * r1["neff"]
            assert np.allclose(r1["estimate"], sh.last["estimate"], rtol=1e-12, atol=0)
            ref = single.download()
            for k in ("x", "y", "yaw", "w", "counts"):
                got = np.concatenate([g[k] for g in gathered])
                if k == "counts":
                    assert np.array_equal(got, ref[k]), (s, k)
                else:
                    assert np.allclose(got, ref[k], rtol=1e-12, atol=0), (s, k)
            got = np.concatenate([g["lm"] for g in gathered])
            mask = np.arange(lcap)[None, :] < ref["counts"][:, None]
            assert np.allclose(got[mask], ref["lm"][mask], rtol=1e-12, atol=0), (s, "lm")
    tot = torch.tensor([moved], device="cuda")
    dist.all_reduce(tot)
    if rank == 0:
        assert nres >= 2, nres
        print("SHARDED_OK world=%d steps=%d resamples=%d migrated=%d mode=%s p2p=%s fallbacks=%d" % (world, steps, nres, int(tot.item()), sh.mode, sh.p2p, sh.fallbacks))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
