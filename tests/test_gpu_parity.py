"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against
the CPU oracle on identical inputs and identical random draws, and against the golden vectors frozen from
the reference.  Bars (north_star): association and resampling indices bit-exact; poses, landmark means /
covariances and weights within 1e-5 relative in fp64 -- the tests below hold them to 1e-9."""
import numpy as np
import pytest

from oracle import fs2_oracle as fo
from oracle import scenarios as sc
from tests.util import load_golden, max_rel, replay_trajectory

pytestmark = pytest.mark.gpu

RTOL = 1e-9   # far inside north_star's 1e-5 (fp64); differences are libm-level (atan2/exp/log/sincos)


def _device_filter(*a, **k):
    from fast_slam_b200 import DeviceFilter
    return DeviceFilter(*a, **k)


def _pair(P, L, lcap, seed, shuffle=False, flags=0):
    init = sc.synthetic_state(seed, P, L, lcap, shuffle=shuffle)
    f = _device_filter(P, lcap, flags=flags)
    f.upload(init["x"], init["y"], init["yaw"], init["w"], init["count"], init["lm"])
    o = fo.OracleFilter(P, lcap)
    o.set_state(init["x"], init["y"], init["yaw"], init["w"], init["count"], lm=init["lm"])
    return init, f, o


def _compare_state(f, o, rtol=RTOL):
    st = f.download()
    np.testing.assert_array_equal(st["counts"], o.count)
    worst = 0.0
    for k, ref in (("x", o.x), ("y", o.y), ("yaw", o.yaw), ("w", o.w)):
        worst = max(worst, max_rel(ref, st[k]))
    mask = np.arange(o.lcap)[None, :] < o.count[:, None]
    worst = max(worst, max_rel(o.lm[mask], st["lm"][mask], floor=1e-9))
    assert worst <= rtol, worst
    np.testing.assert_array_equal(st["status"], o.status)
    return worst


@pytest.mark.parametrize("flags", [0, 1], ids=["speculative", "sequential"])
@pytest.mark.parametrize("P,L,lcap,M,novel,shuffle", [
    (600, 64, 96, 16, 0, False),      # cfg2 shape: every observation matches a distinct landmark
    (600, 64, 96, 16, 4, True),       # with new landmarks appended, shuffled map order
    (257, 256, 320, 32, 0, False),    # cfg3 shape per particle
    (257, 256, 320, 32, 4, True),
    (300, 36, 38, 8, 4, False),       # capacity reached: two of four appends dropped (FS2_ST_MAP_FULL)
    (100, 100, 128, 3, 1, False),     # MP=4 template
    (64, 49, 64, 40, 3, False),       # more than 32 observations: two batches
    (33, 0, 16, 5, 0, False),         # empty maps: everything is appended
])
def test_motion_update_parity(P, L, lcap, M, novel, shuffle, flags):
    if L == 0:
        init = dict(x=np.zeros(P), y=np.zeros(P), yaw=np.zeros(P), w=np.full(P, 1.0 / P),
                    count=np.zeros(P, np.int32), lm=np.zeros((P, lcap, 6)), world=sc.grid_world(36))
        f = _device_filter(P, lcap, flags=flags)
        f.upload(init["x"], init["y"], init["yaw"], init["w"], init["count"], init["lm"])
        o = fo.OracleFilter(P, lcap)
        o.set_state(init["x"], init["y"], init["yaw"], init["w"], init["count"], lm=init["lm"])
    else:
        init, f, o = _pair(P, L, lcap, seed=100 + P, shuffle=shuffle, flags=flags)
    rng = np.random.default_rng(P)
    for step in range(3):
        rot, tr = (0.02, 0.0) if step == 1 else (0.0, 0.01)
        obs = sc.synthetic_obs(5, step, init["world"], M, novel=novel, max_range=9.0)
        noise = rng.normal(0, 0.001 if rot else 0.0055, P)
        a = f.motion_update(rot, tr, obs, noise=noise, want_assoc=True).cpu().numpy()
        o.motion(rot, tr, noise)
        ao = o.update(obs)
        np.testing.assert_array_equal(a, ao, err_msg="association indices, step %d" % step)
        _compare_state(f, o)
    if novel == 0 and L > 0:
        assert (ao >= 0).mean() > 0.97    # a noisy far observation may legitimately miss the gate (Q3)
    f.close()


def test_sequential_dependencies_inside_a_step():
    """Several observations of one step hit the same landmark, and landmarks appended by observation k
    are matched by k' > k (quirk Q7): the speculative path must commit in order."""
    P, lcap = 64, 64
    f = _device_filter(P, lcap)
    o = fo.OracleFilter(P, lcap)
    rng = np.random.default_rng(0)
    for step, (rot, tr, meas) in enumerate(sc.repeat_stream(9, 8, reps=3)):
        noise = rng.normal(0, 0.001 if rot else 0.0055, P)
        obs = np.array(meas)
        a = f.motion_update(rot, tr, obs, noise=noise, want_assoc=True).cpu().numpy()
        o.motion(rot, tr, noise)
        ao = o.update(obs)
        np.testing.assert_array_equal(a, ao, err_msg="step %d" % step)
        _compare_state(f, o)
    assert o.count.max() <= 12     # 6 world landmarks, most observations re-match
    f.close()


def test_many_matches_per_observation_takes_the_sequential_fallback():
    """> 4 landmarks inside one observation's gate that all get touched: match list exhausted."""
    P, lcap, L = 40, 32, 12
    rng = np.random.default_rng(5)
    lm = np.zeros((P, lcap, 6))
    lm[:, :L, 0] = 2.0 + rng.normal(0, 0.05, (P, L))
    lm[:, :L, 1] = 1.0 + rng.normal(0, 0.05, (P, L))
    lm[:, :L, 2] = 0.1; lm[:, :L, 5] = 0.1
    x = rng.normal(0, 0.01, P); y = rng.normal(0, 0.01, P); yaw = rng.normal(0, 0.01, P)
    w = np.full(P, 1.0 / P); cnt = np.full(P, L, np.int32)
    f = _device_filter(P, lcap)
    f.upload(x, y, yaw, w, cnt, lm)
    o = fo.OracleFilter(P, lcap)
    o.set_state(x, y, yaw, w, cnt, lm=lm)
    obs = np.array([(np.hypot(2.0, 1.0) + rng.normal(0, 0.05), np.arctan2(1.0, 2.0) + rng.normal(0, 0.02)) for _ in range(10)])
    a = f.update(obs, want_assoc=True).cpu().numpy()
    ao = o.update(obs)
    np.testing.assert_array_equal(a, ao)
    _compare_state(f, o, rtol=1e-8)
    f.close()


def test_singular_and_indefinite_covariances():
    """np.linalg.inv raising inside associate skips the update (status bit); an indefinite covariance
    just never matches unless its quadratic form happens to be positive."""
    P, lcap = 32, 8
    lm = np.zeros((P, lcap, 6))
    lm[:, 0] = (5.0, 5.0, 0.1, 0.0, 0.0, 0.1)
    lm[:, 1] = (2.0, 0.0, 0.0, 0.0, 0.0, 0.0)         # singular: stops the scan for observations reaching it
    lm[:, 2] = (2.0, 0.1, 0.01, 0.02, 0.02, 0.01)     # indefinite
    lm[:, 3] = (2.0, 0.0, 0.1, 0.0, 0.0, 0.1)
    cnt = np.full(P, 4, np.int32)
    z = np.zeros(P)
    f = _device_filter(P, lcap)
    f.upload(z, z, z, np.full(P, 1.0 / P), cnt, lm)
    o = fo.OracleFilter(P, lcap)
    o.set_state(z, z, z, np.full(P, 1.0 / P), cnt, lm=lm)
    obs = np.array([(np.hypot(5, 5), np.arctan2(5, 5)), (2.0, 0.0), (3.0, -2.0)])
    a = f.update(obs, want_assoc=True).cpu().numpy()
    ao = o.update(obs)
    np.testing.assert_array_equal(a, ao)
    assert (ao[0] == 0).all() and (ao[1] == -2).all()
    _compare_state(f, o)
    f.close()


def _skewed_state(P, L, lcap, seed):
    """Maps of very different lengths (many empty or tiny, some at capacity) and a few singular / indefinite
    covariances: one screener warp of the update kernel then takes several times longer over its particle than
    its neighbours -- the hand-over between the kernel's two warp roles must not depend on them keeping pace."""
    init = sc.synthetic_state(seed, P, L, lcap)
    rng = np.random.default_rng(seed)
    kind = rng.integers(0, 4, P)
    cnt = np.where(kind == 0, rng.integers(0, 3, P), np.where(kind == 1, rng.integers(0, L + 1, P), np.where(kind == 2, L, lcap)))
    cnt = cnt.astype(np.int32)
    lm = init["lm"].copy()
    # slots beyond the generated map: copies of earlier landmarks (several matches per observation)
    for j in range(L, lcap):
        lm[:, j] = lm[:, j - L]
        lm[:, j, 0:2] += rng.normal(0, 0.05, (P, 2))
    sing = rng.choice(P, size=max(4, P // 200), replace=False)
    for i in sing:
        j = int(rng.integers(0, max(1, cnt[i])))
        lm[i, j, 2:6] = 0.0 if i % 2 else (0.01, 0.02, 0.02, 0.01)     # singular / indefinite
    init["lm"], init["count"] = lm, cnt
    return init


@pytest.mark.parametrize("P,L,lcap,M", [(8192 + 37, 100, 128, 24), (20000, 36, 48, 32)])
def test_skewed_map_sizes_many_tickets_per_block(P, L, lcap, M, monkeypatch):
    """More than 16 tickets per thread block with strongly non-uniform work per particle: the warp-specialised
    kernel against the oracle AND against the single-role kernel (FS2_KERNEL=v3) on the same state."""
    init = _skewed_state(P, L, lcap, seed=4242 + P)
    f = _device_filter(P, lcap)
    monkeypatch.setenv("FS2_KERNEL", "v3")
    f3 = _device_filter(P, lcap)
    monkeypatch.delenv("FS2_KERNEL")
    o = fo.OracleFilter(P, lcap)
    for flt in (f, f3):
        flt.upload(init["x"], init["y"], init["yaw"], init["w"], init["count"], init["lm"])
    o.set_state(init["x"], init["y"], init["yaw"], init["w"], init["count"], lm=init["lm"])
    rng = np.random.default_rng(P)
    for step in range(4):
        rot, tr = (0.02, 0.0) if step == 2 else (0.0, 0.01)
        obs = sc.synthetic_obs(9, step, init["world"], M, novel=2, max_range=9.0)
        noise = rng.normal(0, 0.001 if rot else 0.0055, P)
        a = f.motion_update(rot, tr, obs, noise=noise, want_assoc=True).cpu().numpy()
        a3 = f3.motion_update(rot, tr, obs, noise=noise, want_assoc=True).cpu().numpy()
        o.motion(rot, tr, noise)
        ao = o.update(obs)
        np.testing.assert_array_equal(a, ao, err_msg="association indices (warp-specialised), step %d" % step)
        np.testing.assert_array_equal(a3, ao, err_msg="association indices (single-role), step %d" % step)
        _compare_state(f, o, rtol=1e-8)
        _compare_state(f3, o, rtol=1e-8)
    f.close()
    f3.close()


@pytest.mark.parametrize("name,P,lcap", [("traj_drive.npz", 24, 64), ("traj_repeat.npz", 16, 48)])
def test_golden_trajectory_from_origin(name, P, lcap):
    """Whole steps (fs2_step_host) on the reference's own recorded draws, against the reference's outputs."""
    g = load_golden(name)
    f = _device_filter(P, lcap)
    worst, nres = replay_trajectory(g, f, lcap, rtol=RTOL)
    assert nres >= 1
    f.close()


def test_golden_trajectory_synthetic_state():
    g = load_golden("traj_synth.npz")
    f = _device_filter(12, 48)
    f.upload(g["init_x"], g["init_y"], g["init_yaw"], g["init_w"], g["init_counts"], np.nan_to_num(g["init_lm"]))
    replay_trajectory(g, f, 48, rtol=RTOL)
    f.close()


def test_multi_step_with_resampling_vs_oracle():
    """40 whole steps at 3000 particles x 64 landmarks x 16 observations: every association index and every
    resampling index equals the oracle's, state within RTOL, several resamples on the way."""
    P, L, lcap, M = 3000, 64, 96, 16
    init, f, o = _pair(P, L, lcap, seed=77)
    o.wkind[:] = 0
    rng = np.random.default_rng(1)
    nres = 0
    for step in range(40):
        rot, tr = sc.synthetic_odometry(step)
        obs = sc.synthetic_obs(77, step, init["world"], M, novel=2 if step % 8 == 7 else 0, max_range=9.0)
        noise = rng.normal(0, 0.001 if rot else 0.0055, P)
        u0 = float(rng.uniform(0, 1.0 / P))
        g = f.step(rot, tr, obs, noise=noise, u0=u0)
        r = o.step(rot, tr, obs, noise, u0)
        np.testing.assert_array_equal(g["assoc"], r["assoc"], err_msg="step %d" % step)
        assert g["resampled"] == r["resampled"], step
        np.testing.assert_array_equal(g["resample_idx"], r["resample_idx"], err_msg="step %d" % step)
        nres += int(r["resampled"])
        assert max_rel(r["estimate"], g["estimate"]) < RTOL
        assert abs(g["neff"] - r["neff"]) <= 1e-9 * r["neff"]
        if step % 5 == 4 or r["resampled"]:
            _compare_state(f, o)
    assert nres >= 2
    f.close()


@pytest.mark.parametrize("P,w_reset", [(3000, False), (4099, True)])
def test_deferred_map_copies_ride_in_the_next_update(P, w_reset):
    """After a resample the offspring's deep copies (fast_slam_2.py:192-196) are written by the NEXT update kernel:
    leaders stream their map, their followers share the match lists (csrc/fs2_update_ws.cuh, DEFER).  Here the state
    is NOT read back after a resampling step (a download would make the copies on the spot), so the steps that follow
    a resample run the deferred form; associations of every particle are compared on every step, the state a few steps
    later.  w_reset concentrates the weight on a handful of particles now and then: lineages of hundreds of offspring
    (many groups of eight, leaders that get a real copy)."""
    import torch
    L, lcap, M = 64, 96, 16
    init, f, o = _pair(P, L, lcap, seed=78)
    o.wkind[:] = 0
    rng = np.random.default_rng(2)
    nres = after_res = 0
    was_res = False
    for step in range(36):
        if w_reset and step % 6 == 3 and not was_res:
            w = np.full(P, 1e-6)
            w[rng.choice(P, 5, replace=False)] = 1.0
            f.w.copy_(torch.as_tensor(w, device=f.w.device))     # (upload() would reset the rest of the state)
            o.w[:] = w
        rot, tr = sc.synthetic_odometry(step)
        obs = sc.synthetic_obs(78, step, init["world"], M, novel=2 if step % 8 == 7 else 0, max_range=9.0)
        noise = rng.normal(0, 0.001 if rot else 0.0055, P)
        u0 = float(rng.uniform(0, 1.0 / P))
        g = f.step(rot, tr, obs, noise=noise, u0=u0)
        r = o.step(rot, tr, obs, noise, u0)
        np.testing.assert_array_equal(g["assoc"], r["assoc"], err_msg="step %d" % step)
        assert g["resampled"] == r["resampled"], step
        np.testing.assert_array_equal(g["resample_idx"], r["resample_idx"], err_msg="step %d" % step)
        after_res += int(was_res)
        if was_res and not r["resampled"]:
            _compare_state(f, o)                     # the step after a resample: leaders and followers updated
        was_res = bool(r["resampled"])
        nres += int(was_res)
    assert nres >= 3 and after_res >= 3
    _compare_state(f, o)
    f.close()


@pytest.mark.parametrize("M", [1, 2, 5, 31])
def test_few_observations_keep_the_warp_together(M):
    """Fewer than 32 observations leave idle lanes in every applier warp.  Round 2 found two ways for them to run a turn
    apart from the busy ones (a ticket taken on different turns; state kept in warp-uniform registers): whole steps with
    resampling and deferred copies at 1, 2, 5 and 31 observations against the oracle, every association of every particle."""
    import torch
    P, L, lcap = 6000, 64, 96
    init, f, o = _pair(P, L, lcap, seed=80 + M)
    o.wkind[:] = 0
    rng = np.random.default_rng(M)
    nres = 0
    was_res = False
    for step in range(16):
        rot, tr = sc.synthetic_odometry(step)
        obs = sc.synthetic_obs(80 + M, step, init["world"], M, novel=1 if (step % 5 == 4 and M > 1) else 0, max_range=9.0)
        noise = rng.normal(0, 0.001 if rot else 0.0055, P)
        u0 = float(rng.uniform(0, 1.0 / P))
        if step % 4 == 1 and not was_res:          # concentrate the weight: a resample with long lineages follows
            w = np.full(P, 1e-6)
            w[rng.choice(P, 40, replace=False)] = 1.0
            f.w.copy_(torch.as_tensor(w, device=f.w.device))     # (upload() would reset the rest of the state)
            o.w[:] = w
        g = f.step(rot, tr, obs, noise=noise, u0=u0)
        r = o.step(rot, tr, obs, noise, u0)
        np.testing.assert_array_equal(g["assoc"], r["assoc"], err_msg="step %d" % step)
        assert g["resampled"] == r["resampled"], step
        np.testing.assert_array_equal(g["resample_idx"], r["resample_idx"], err_msg="step %d" % step)
        if was_res and not r["resampled"]:
            _compare_state(f, o)
        was_res = bool(r["resampled"])
        nres += int(was_res)
    assert nres >= 2
    _compare_state(f, o)
    f.close()


def test_deferred_map_copies_equal_immediate_copies(monkeypatch):
    """The same stream with FS2_DEFER=0 (copies made by the resample) and with the default: bit-identical states; and the
    other consumers of a map whose copy is still owed (motion-only step, map clustering, download of single particles)."""
    import torch
    from fast_slam_b200.synthetic import fill_synthetic_device, synthetic_obs, synthetic_odometry
    from fast_slam_b200.filter import _hash_uniform
    P, L, lcap, M = 20000, 49, 64, 24
    outs = []
    for defer in ("0", "1"):
        monkeypatch.setenv("FS2_DEFER", defer)
        f = _device_filter(P, lcap, seed=5)
        world = fill_synthetic_device(f, L, 5)
        nres = 0
        kl = None
        for s in range(24):
            rot, tr = synthetic_odometry(s)
            obs = synthetic_obs(5, s, world, M, novel=1 if s % 6 == 5 else 0, max_range=12.0)
            g = f.step(rot, tr, obs, u0=_hash_uniform(5, s) / P, step_index=s, want_assoc=False, want_ancestor=False)
            nres += int(g["resampled"])
            if g["resampled"] and nres == 2:
                f.draw_noise(0.0055, 1000 + s)
                f.motion(0.0, 0.01)                   # no observations: nothing streams the maps, the copies are made first
            if g["resampled"] and nres == 3:
                kl = f.known_landmarks()
            if g["resampled"] and nres == 4:
                part = f.download_particles(np.arange(0, P, 97))
        assert nres >= 4
        st = f.download()
        st["kl"], st["part"] = kl, part
        outs.append(st)
        f.close()
    a, b = outs
    for k in ("x", "y", "yaw", "w", "counts", "status"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    mask = np.arange(lcap)[None, :] < a["counts"][:, None]
    np.testing.assert_array_equal(a["lm"][mask], b["lm"][mask])
    np.testing.assert_array_equal(a["kl"][0], b["kl"][0])
    np.testing.assert_array_equal(a["kl"][1], b["kl"][1])
    for k in ("x", "w", "counts"):
        np.testing.assert_array_equal(a["part"][k], b["part"][k])
    pm = np.arange(lcap)[None, :] < a["part"]["counts"][:, None]
    np.testing.assert_array_equal(a["part"]["lm"][pm], b["part"]["lm"][pm])


def _weights(kind, n, rng):
    if kind == "uniform":
        w = rng.uniform(0, 1, n)
    elif kind == "skewed":
        w = rng.uniform(0, 1, n) ** 12
    elif kind == "lognormal":
        w = np.exp(rng.normal(0, 5, n))
    elif kind == "equal":
        w = np.full(n, 1.0)
    elif kind == "dyadic":           # exact ties in the running sum's rounding
        w = rng.integers(1, 1 << 20, n).astype(np.float64) * 2.0 ** -60
        w[::7] = 2.0 ** -54
        w[0] = 0.25
    elif kind == "sparse":
        w = rng.uniform(0, 1, n)
        w[rng.uniform(0, 1, n) < 0.9] = 0.0
        w[-1] = 0.3
    elif kind == "leading_zeros":
        w = rng.uniform(0, 1, n)
        w[: n // 2] = 0.0
    elif kind == "one_heavy":
        w = np.full(n, 1e-13)
        w[n // 3] = 1.0
    else:
        raise ValueError(kind)
    if kind != "dyadic":
        w = w / w.sum()
    else:                            # scale by a power of two (keeps the ties) so that the total is >= 1:
        w = w * 2.0 ** np.ceil(-np.log2(w.sum()))   # the reference's walk spins for ever on a total < u
    return w


@pytest.mark.parametrize("n", [1, 2, 5, 1000, 1024, 1025, 4096, 100003, 1 << 20])
@pytest.mark.parametrize("kind", ["uniform", "skewed", "lognormal", "equal", "dyadic", "sparse", "leading_zeros", "one_heavy"])
def test_resample_scan_is_bit_exact(n, kind):
    """The parallel scan reproduces the SEQUENTIAL fp64 running sum bit for bit (quirk Q10), hence the
    resampling indices; checked stage-wise on identical weights."""
    import torch
    rng = np.random.default_rng(n * 31 + len(kind))
    w = _weights(kind, n, rng)
    u0 = float(rng.uniform(0, 1.0 / n))
    f = _device_filter(n, 1)
    wt = torch.as_tensor(w, device="cuda")
    idx = f.resample_indices(u0, w_all=wt).cpu().numpy()
    cum = f.cumsum.cpu().numpy()
    ref_cum = fo.cumsum_seq(w)
    assert np.array_equal(cum.view(np.int64), ref_cum.view(np.int64)), "running sum differs in %d places" % (cum != ref_cum).sum()
    ref_idx, _ = fo.resample_indices(w, u0)
    np.testing.assert_array_equal(idx, ref_idx)
    f.close()


def test_resample_kats_from_the_reference():
    import torch
    k = load_golden("stage_kats.npz")
    for tag in ["n1", "n2", "n7", "n64", "n1000", "n4096"]:
        W, U, I = k["res_%s_w" % tag], k["res_%s_u0" % tag], k["res_%s_idx" % tag]
        f = _device_filter(W.shape[1], 1)
        for c in range(len(U)):
            idx = f.resample_indices(float(U[c]), w_all=torch.as_tensor(W[c], device="cuda")).cpu().numpy()
            np.testing.assert_array_equal(idx, I[c], err_msg="%s case %d" % (tag, c))
        f.close()


def test_resample_negative_and_nan_weights_take_the_literal_path():
    import torch
    rng = np.random.default_rng(4)
    n = 3000
    w = rng.uniform(0, 1, n); w /= w.sum()
    w[100] = -1e-4
    f = _device_filter(n, 1)
    idx = f.resample_indices(1e-5, w_all=torch.as_tensor(w, device="cuda")).cpu().numpy()
    ref, _ = fo.resample_indices(w, 1e-5)
    np.testing.assert_array_equal(idx, ref)
    f.close()


def test_resample_slot_range_matches_whole():
    """A shard asks only for its own destination slots [m_begin, m_begin+m_count) of the global resample."""
    import torch
    rng = np.random.default_rng(8)
    n = 50000
    w = rng.uniform(0, 1, n) ** 4; w /= w.sum()
    wt = torch.as_tensor(w, device="cuda")
    f = _device_filter(n, 1)
    whole = f.resample_indices(3e-6, w_all=wt).cpu().numpy()
    part = f.resample_indices(3e-6, w_all=wt, m_begin=12500, m_count=12500).cpu().numpy()
    np.testing.assert_array_equal(part, whole[12500:25000])
    f.close()


def test_normalize_neff_argmax_kats():
    """fast_slam_2.py:161-175, 212-223, 201-210 against the reference's frozen outputs.  The device total
    is a tree sum, the reference's a sequential/compensated one: equal to rounding, so weights agree to
    1e-15 relative; given the reference's total the division is bit-exact."""
    import torch
    k = load_golden("stage_kats.npz")
    for c in range(len(k["norm_in"])):
        w = k["norm_in"][c]
        P = len(w)
        f = _device_filter(P, 1)
        f.upload(w=w)
        f.weight_total()
        f.normalize()
        st = f.download(maps=False)
        assert max_rel(k["norm_out"][c], st["w"]) < 1e-14, c
        stats = f.stats.cpu().numpy()
        assert abs(stats[2] - k["neff_out"][c]) <= 1e-12 * k["neff_out"][c]
        # bit-exact division when handed the oracle's total
        tot = fo.lib().fs2o_weight_total(P, fo._dp(np.ascontiguousarray(w)), fo._bp(np.ascontiguousarray(k["norm_kind"][c])), None)
        f.upload(w=w)
        f.normalize(total=torch.tensor([tot], dtype=torch.float64, device="cuda"))
        np.testing.assert_array_equal(f.download(maps=False)["w"], k["norm_out"][c])
        f.close()


def test_argmax_first_occurrence_on_ties():
    P = 5000
    w = np.full(P, 0.1)
    w[[777, 1234, 4000]] = 0.5
    x = np.arange(P, dtype=np.float64)
    f = _device_filter(P, 1)
    f.upload(x=x, w=w)
    f.weight_total(); f.normalize()
    stats = f.stats.cpu().numpy()
    assert int(stats[4]) == 777 and stats[5] == 777.0
    f.close()


def test_gather_is_a_deep_copy_in_ancestor_order():
    """deepcopy of the survivors incl. weight and map (fast_slam_2.py:196-199) via copy-on-resample slots."""
    import torch
    P, L, lcap = 2000, 25, 28
    init, f, o = _pair(P, L, lcap, seed=3)
    rng = np.random.default_rng(2)
    w = rng.uniform(0, 1, P) ** 8; w /= w.sum()
    f.upload(init["x"], init["y"], init["yaw"], w, rng.integers(0, L + 1, P).astype(np.int32), init["lm"])
    before = f.download()
    for rep in range(3):                 # repeated resamples permute the slot table further
        anc = np.sort(rng.choice(P, size=P, p=w)).astype(np.int32)
        f.gather(torch.as_tensor(anc, device="cuda"))
        after = f.download()
        for key in ("x", "y", "yaw", "w", "counts"):
            np.testing.assert_array_equal(after[key], before[key][anc])
        for m in range(0, P, 7):
            n = before["counts"][anc[m]]
            np.testing.assert_array_equal(after["lm"][m, :n], before["lm"][anc[m], :n])
        before = after
    # copies are independent: updating one offspring must not touch its siblings
    obs = sc.synthetic_obs(3, 0, init["world"], 4, max_range=5.0)
    f.update(obs)
    o2 = fo.OracleFilter(P, lcap)
    o2.set_state(before["x"], before["y"], before["yaw"], before["w"], before["counts"], lm=before["lm"])
    o2.update(obs)
    _compare_state(f, o2)
    f.close()


def test_device_noise_generator():
    import torch
    P = 1 << 16
    f = _device_filter(P, 1, seed=42)
    a = f.draw_noise(0.0055, 7).clone()
    b = f.draw_noise(0.0055, 7).clone()
    c = f.draw_noise(0.0055, 8).clone()
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(float(a.mean())) < 5 * 0.0055 / np.sqrt(P)
    assert abs(float(a.std()) / 0.0055 - 1) < 0.02
    # a shard draws exactly the numbers of its global index range
    g = _device_filter(P // 4, 1, seed=42, global_particles=P, global_offset=P // 2)
    s = g.draw_noise(0.0055, 7)
    assert torch.equal(s, a[P // 2: P // 2 + P // 4])
    f.close(); g.close()


def test_full_size_particles_are_independent_samples():
    """cfg3 shape at 2^17 particles x 256 landmarks x 32 observations (device-generated state): 512 sampled
    particles are pulled out before and after one fused motion+update launch and re-done by the oracle."""
    import torch
    P, L, lcap, M = 1 << 17, 256, 320, 32
    from bench import make_synthetic_filter, synthetic_step_inputs
    f, world = make_synthetic_filter(P, L, lcap, seed=1234)
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(P, 512, replace=False))
    before = f.download_particles(sel)
    rot, tr, obs = synthetic_step_inputs(1234, 3, world, M, novel=4)
    f.draw_noise(0.0055, 3)
    noise = f.noise[torch.as_tensor(sel, device="cuda")].cpu().numpy()
    a = f.motion_update(rot, tr, obs, want_assoc=True)[:, torch.as_tensor(sel, device="cuda")].cpu().numpy()
    after = f.download_particles(sel)
    o = fo.OracleFilter(len(sel), lcap)
    o.set_state(before["x"], before["y"], before["yaw"], before["w"], before["counts"], lm=before["lm"])
    o.motion(rot, tr, noise)
    ao = o.update(obs)
    np.testing.assert_array_equal(a, ao)
    # a novel point either starts a landmark or, inside the 2.53 m gate of one started by an earlier
    # novel point of the same step, matches that one (index >= L)
    assert (ao[:M - 4] >= 0).mean() > 0.97 and ((ao[M - 4:] == -1) | (ao[M - 4:] >= L)).all()
    np.testing.assert_array_equal(after["counts"], o.count)
    for k, ref in (("x", o.x), ("y", o.y), ("yaw", o.yaw), ("w", o.w)):
        assert max_rel(ref, after[k]) < RTOL
    mask = np.arange(lcap)[None, :] < o.count[:, None]
    assert max_rel(o.lm[mask], after["lm"][mask], floor=1e-9) < RTOL
    f.close()


def test_full_size_stream_with_resampling():
    """BASELINE.json config 3 at its full size, 2^20 particles x 256 landmarks x 32 observations (16 GB of maps), three
    steps of the bench stream.  Per step: 384 sampled particles are re-done by the oracle (association exact, state
    to 1e-9); the weight total is the sum of the per-particle weights; the resampling indices of all 2^20 particles
    are compared with the oracle's walk bit for bit; and after the gather sampled offspring are deep copies of their
    ancestors' pre-resample state."""
    import torch
    from bench import make_synthetic_filter, synthetic_step_inputs
    from fast_slam_b200 import _lib
    from fast_slam_b200.filter import _hash_uniform
    P, L, lcap, M = 1 << 20, 256, 320, 32
    f, world = make_synthetic_filter(P, L, lcap, seed=1234)
    rng = np.random.default_rng(1)
    resamples = 0
    for s in range(3):
        sel = np.sort(rng.choice(P, 384, replace=False))
        tsel = torch.as_tensor(sel, device="cuda")
        before = f.download_particles(sel)
        rot, tr, obs = synthetic_step_inputs(1234, s, world, M)
        f.draw_noise(0.001 if rot != 0 else 0.0055, s)
        noise = f.noise[tsel].cpu().numpy()
        a = f.motion_update(rot, tr, obs, want_assoc=True)[:, tsel].cpu().numpy()
        after = f.download_particles(sel)
        o = fo.OracleFilter(len(sel), lcap)
        o.set_state(before["x"], before["y"], before["yaw"], before["w"], before["counts"], lm=before["lm"])
        o.motion(rot, tr, noise)
        np.testing.assert_array_equal(a, o.update(obs))
        np.testing.assert_array_equal(after["counts"], o.count)
        for k, ref in (("x", o.x), ("y", o.y), ("yaw", o.yaw), ("w", o.w)):
            assert max_rel(ref, after[k]) < RTOL, (s, k)
        mask = np.arange(lcap)[None, :] < o.count[:, None]
        assert max_rel(o.lm[mask], after["lm"][mask], floor=1e-9) < RTOL
        # weights: total, normalisation, effective sample size
        w_raw = f.w.cpu().numpy().copy()
        f.weight_total()
        f.normalize()
        stats = f.stats.cpu().numpy()
        assert abs(stats[_lib.STAT_TOTAL] - w_raw.sum()) <= 1e-12 * w_raw.sum()
        w = f.w.cpu().numpy().copy()
        total = stats[_lib.STAT_TOTAL]
        assert total >= 1e-5                                                 # (else every weight becomes 1/N, :168-170)
        w_rule = np.where(w_raw < 1e-5, w_raw, w_raw / total)                # fast_slam_2.py:173: small weights stay
        np.testing.assert_array_equal(w, w_rule)
        sumsq = np.square(w).sum()
        neff = P if sumsq < 1.0 / P else 1.0 / sumsq                         # fast_slam_2.py:220-223
        assert abs(stats[_lib.STAT_NEFF] - neff) <= 1e-9 * neff
        assert int(stats[_lib.STAT_ARGMAX]) == int(np.argmax(w))
        if stats[_lib.STAT_NEFF] < P / 2:
            resamples += 1
            u0 = _hash_uniform(1234, s) / P
            anc = f.resample_indices(u0)
            want, stuck = fo.resample_indices(w, u0)
            assert not stuck
            anc_h = anc.cpu().numpy()
            np.testing.assert_array_equal(anc_h, want)                      # all 2^20 indices, bit for bit
            m_sel = np.sort(rng.choice(P, 256, replace=False))
            parents = f.download_particles(anc_h[m_sel].astype(np.int64))
            f.gather(anc)
            kids = f.download_particles(m_sel)
            for k in ("x", "y", "yaw", "w", "counts"):
                np.testing.assert_array_equal(kids[k], parents[k])
            mk = np.arange(lcap)[None, :] < parents["counts"][:, None]
            np.testing.assert_array_equal(kids["lm"][mk], parents["lm"][mk])
    assert resamples >= 1
    f.close()
