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
    from fast_slam_b200.selfcheck import sharded_equals_single
    chk = sharded_equals_single(P, L, lcap, M, seed, steps)
    sh, single, nres = chk["sharded"], chk["single"], chk["resamples"]
    # ---- map clustering over the shards (row N1) = the single filter's, here all at cell level ...
    kl = sh.known_landmarks()
    if rank == 0:
        ref = single.known_landmarks()
        assert kl is not None and ref is not None
        assert np.array_equal(kl[1], ref[1]) and np.array_equal(kl[0], ref[0]), "sharded known_landmarks differs"
        assert kl[2]["n_points"] == ref[2]["n_points"] and kl[2]["clusters"] == ref[2]["clusters"] >= L
    # ... and on a small scattered map where cores, border points and noise are decided point by point
    from oracle import known_landmarks_oracle as ko
    p2, l2 = 48, 32
    rng = np.random.default_rng(5)
    cnt_all = rng.integers(20, l2 + 1, size=p2 * world).astype(np.int32)
    lm_all = np.zeros((p2 * world, l2, 6))
    lm_all[:, :, 0:2] = rng.uniform(-5, 5, size=(p2 * world, l2, 2)) + [100.0, -40.0]
    sh2 = ShardedFilter(p2, l2, seed=1)
    sh2.store.upload(count=cnt_all[rank * p2:(rank + 1) * p2], lm=lm_all[rank * p2:(rank + 1) * p2])
    kl2 = sh2.known_landmarks()
    maps = [lm_all[p, :cnt_all[p], 0:2] for p in range(p2 * world)]
    want = ko.update_known_landmarks(maps)
    assert kl2 is not None and want is not None and kl2[2]["involved_points"] > 0
    assert np.array_equal(kl2[1], want[1]), (kl2[1], want[1])
    assert np.allclose(kl2[0], want[0], rtol=0, atol=1e-11)
    assert kl2[2]["noise_points"] == int(cnt_all.sum() - want[1].sum())
    assert world > 4 or kl2[2]["noise_points"] > 0      # (8 ranks put 10^4 points on the same 10 m square: none is noise)
    alls = [None] * world
    dist.all_gather_object(alls, (kl2[0].tobytes(), kl2[1].tobytes()))
    assert all(a == alls[0] for a in alls), "ranks disagree on the clustering"
    sh2.store.close()
    if rank == 0:
        assert nres >= 2, nres
        print("SHARDED_OK world=%d steps=%d resamples=%d migrated=%d mode=%s p2p=%s fallbacks=%d" % (world, steps, nres, chk["migrated"], sh.mode, sh.p2p, sh.fallbacks))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
