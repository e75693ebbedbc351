"""Randomised parity of the fused motion + update kernel against the C restatement (oracle/fs2_oracle.c): random maps
(clustered landmarks with small, default 0.1 I and large covariances, correlated and near-singular ones), random
poses, observations that hit landmarks once, several times, in clusters of overlapping gates, or nothing; maps near
their capacity; 0..40 observations per step; both the speculative and the forced-sequential path.  Association indices
and status words must be equal, state to 1e-8.
    python scripts/update_stress.py [cases] [seed]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_slam_b200 import DeviceFilter               # noqa: E402
from oracle import fs2_oracle as fo                   # noqa: E402


def max_rel(a, b, floor=1e-300):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(a), floor))) if a.size else 0.0


def make(rng):
    P = int(rng.integers(1, 70))
    L = int(rng.integers(0, 90))
    lcap = L + int(rng.choice([0, 1, 2, 8, 40]))
    lcap = max(lcap, 1)
    M = int(rng.choice([0, 1, 2, 5, 16, 31, 32, 33, 40]))
    spread = float(rng.choice([0.5, 2.0, 6.0, 15.0]))
    world = rng.uniform(-spread, spread, (max(L, 1), 2))
    near = np.hypot(world[:, 0], world[:, 1]) < 0.4          # a laser has a minimum range; at r -> 0 the bearing
    world[near] += 0.8 * np.sign(world[near] + 1e-9)         # Jacobian ~ 1/r makes Q arbitrarily ill-conditioned
    lm = np.zeros((P, lcap, 6))
    cnt = np.full(P, L, np.int32)
    if rng.random() < 0.3 and L > 2:
        cnt = rng.integers(0, L + 1, P).astype(np.int32)
    kind = rng.integers(0, 4, size=L)
    for j in range(L):
        lm[:, j, 0:2] = world[j] + rng.normal(0, 0.02, (P, 2))
        if kind[j] == 0:
            a = rng.uniform(0.002, 0.006, P); c = rng.uniform(0.002, 0.006, P); b = rng.uniform(-0.001, 0.001, P)
        elif kind[j] == 1:
            a = np.full(P, 0.1); c = np.full(P, 0.1); b = np.zeros(P)
        elif kind[j] == 2:
            a = rng.uniform(0.2, 2.0, P); c = rng.uniform(0.2, 2.0, P); b = rng.uniform(-0.1, 0.1, P)
        else:                                            # strongly correlated, sometimes not symmetric
            a = rng.uniform(0.01, 0.05, P); c = rng.uniform(0.01, 0.05, P); b = 0.95 * np.sqrt(a * c) * rng.choice([-1, 1])
        lm[:, j, 2] = a; lm[:, j, 5] = c; lm[:, j, 3] = b; lm[:, j, 4] = b * (1.0 + (rng.random() < 0.2) * 1e-3)
    x = rng.normal(0, 0.05, P); y = rng.normal(0, 0.05, P); yaw = rng.normal(0, 0.02, P)
    w = rng.uniform(0.1, 1.0, P); w /= w.sum()
    obs = []
    for _ in range(M):
        r = rng.random()
        if L and r < 0.6:
            t = world[rng.integers(0, L)] + rng.normal(0, 0.03, 2)
        elif L and r < 0.75 and obs:
            t = None
            obs.append(obs[rng.integers(0, len(obs))] + rng.normal(0, 0.005, 2))          # nearly the same observation again
            continue
        else:
            t = rng.uniform(-spread * 1.2, spread * 1.2, 2)
        if np.hypot(*t) < 0.4:
            t = t + 0.8 * np.sign(t + 1e-9)
        obs.append(np.array([np.hypot(*t), np.arctan2(t[1], t[0])]))
    obs = np.array(obs).reshape(-1, 2)
    return P, lcap, x, y, yaw, w, cnt, lm, obs


def main(cases=None, seed=None):
    if cases is None:
        cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    if seed is None:
        seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    bad = 0
    t0 = time.time()
    for k in range(cases):
        P, lcap, x, y, yaw, w, cnt, lm, obs = make(rng)
        flags = int(rng.random() < 0.2)
        f = DeviceFilter(P, lcap, flags=flags)
        f.upload(x, y, yaw, w, cnt, lm)
        o = fo.OracleFilter(P, lcap)
        o.set_state(x, y, yaw, w, cnt, lm=lm)
        ok = True
        why = ""
        for step in range(2):
            rot, tr = ((0.03, 0.0) if rng.random() < 0.3 else (0.0, 0.02))
            noise = rng.normal(0, 0.003, P)
            ob = obs if step == 0 else obs[rng.permutation(len(obs))]
            a = f.motion_update(rot, tr, ob, noise=noise, want_assoc=True)
            a = a.cpu().numpy() if len(ob) else np.zeros((0, P), np.int32)
            o.motion(rot, tr, noise)
            ao = o.update(ob) if len(ob) else np.zeros((0, P), np.int32)
            st = f.download()
            if not np.array_equal(a, ao):
                ok, why = False, "assoc step %d (%d of %d differ)" % (step, int((a != ao).sum()), a.size)
                break
            if not (np.array_equal(st["counts"], o.count) and np.array_equal(st["status"], o.status)):
                ok, why = False, "counts/status step %d" % step
                break
            mask = np.arange(lcap)[None, :] < o.count[:, None]
            # a weight below ~1e-250 may have passed through the subnormal range inside the reference's sequential
            # product w *= l_k (the kernel multiplies the likelihoods of a round as a tree): there the reference's own
            # digits are rounding noise, so such weights are compared absolutely
            worst = max(max_rel(o.x, st["x"]), max_rel(o.y, st["y"]), max_rel(o.yaw, st["yaw"]), max_rel(o.w, st["w"], floor=1e-250),
                        max_rel(o.lm[mask], st["lm"][mask], floor=1e-7))      # off-diagonals of ~1e-9 are cancellation noise
            if not worst < 1e-8:
                rl = np.abs(o.lm - st["lm"]) / np.maximum(np.abs(o.lm), 1e-7) * mask[:, :, None]
                p_, j_, c_ = np.unravel_index(np.argmax(rl), rl.shape)
                iw = int(np.argmax(np.abs(o.w - st["w"]) / np.maximum(np.abs(o.w), 1e-250)))
                print("   worst weight: oracle %.17g device %.17g" % (o.w[iw], st["w"][iw]))
                ok, why = False, ("state step %d rel %.3g (pose/weight %.2g; worst landmark entry [%d][%d][%d]: oracle %.17g device %.17g, "
                                  "its row %s)" % (step, worst, max(max_rel(o.x, st["x"]), max_rel(o.y, st["y"]), max_rel(o.yaw, st["yaw"]), max_rel(o.w, st["w"])),
                                                   p_, j_, c_, o.lm[p_, j_, c_], st["lm"][p_, j_, c_], np.array2string(o.lm[p_, j_], precision=6)))
                break
        f.close()
        if not ok and os.environ.get("FS2_STRESS_DUMP"):
            os.makedirs("gpurun_out", exist_ok=True)
            np.savez("gpurun_out/update_bad_%d.npz" % k, x=x, y=y, yaw=yaw, w=w, cnt=cnt, lm=lm, obs=obs, w_dev=st["w"], w_ora=o.w)
        if not ok:
            bad += 1
            print("MISMATCH case %d: P %d lcap %d M %d flags %d: %s" % (k, P, lcap, len(obs), flags, why), flush=True)
    print("update_stress: %d cases, %d mismatches, %.1f s" % (cases, bad, time.time() - t0))
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
