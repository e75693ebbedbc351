"""GPU: the reference-facing Python API (fast_slam_2.FastSLAM2) end to end."""
import contextlib
import io

import numpy as np
import pytest

from oracle import scenarios as sc
from tests.util import load_golden, max_rel

pytestmark = pytest.mark.gpu


def test_drop_in_replays_the_reference_trajectory_from_the_same_seed():
    """np.random.seed(7); FastSLAM2(); iterate(...) x 60 -- the exact calls gen_golden.py made on the
    reference -- with config.RNG = "reference": same draws, same resampling decisions, same estimates."""
    import fast_slam_2
    from fast_slam_2 import FastSLAM2, Measurement, config
    assert fast_slam_2.config is config
    g = load_golden("traj_drive.npz")
    config.NUM_PARTICLES, config.LANDMARK_CAPACITY, config.RNG = 24, 64, "reference"
    try:
        np.random.seed(7)
        f = FastSLAM2()
        nres = 0
        with contextlib.redirect_stdout(io.StringIO()) as out:
            for s, (rot, tr, meas) in enumerate(sc.drive_stream(1, 60)):
                est = f.iterate(rot, tr, [Measurement(d, a) for d, a in meas])
                assert isinstance(est, tuple) and len(est) == 3 and all(isinstance(v, float) for v in est)
                assert max_rel(g["estimate"][s], np.array(est)) < 1e-9, s
                assert f.last["resampled"] == bool(g["resampled"][s]), s
                nres += f.last["resampled"]
        assert out.getvalue().count("RESAMPLING") == nres >= 1              # fast_slam_2.py:63
        ps = f.particles
        assert len(ps) == 24
        np.testing.assert_array_equal([len(p.landmarks) for p in ps], g["counts"][-1])
        assert max_rel(g["x"][-1], np.array([p.x for p in ps])) < 1e-9
        assert max_rel(g["w"][-1], np.array([p.weight for p in ps])) < 1e-9
        p0 = ps[0]
        assert max_rel(g["lm"][-1][0, 0, 2:6].reshape(2, 2), np.asarray(p0.landmarks[0].cov)) < 1e-8
        assert isinstance(p0.to_dict()["x"], float)                          # serializer.py:39 needs JSON-able values
        f.store.close()
    finally:
        config.NUM_PARTICLES, config.LANDMARK_CAPACITY, config.RNG = 20, 256, "device"


def test_device_rng_mode_runs_and_keeps_invariants():
    from fast_slam_2 import FastSLAM2, Measurement, config
    config.NUM_PARTICLES, config.LANDMARK_CAPACITY, config.RNG, config.SEED = 5000, 32, "device", 3
    try:
        f = FastSLAM2()
        with contextlib.redirect_stdout(io.StringIO()):
            for s, (rot, tr, meas) in enumerate(sc.drive_stream(2, 30)):
                x, y, yaw = f.iterate(rot, tr, [Measurement(d, a) for d, a in meas])
                assert np.isfinite([x, y, yaw]).all() and -np.pi <= yaw <= np.pi
        w = f.particles.weights()
        assert (w >= 0).all() and np.isfinite(w).all()
        assert abs(w.sum() - 1) < 1e-9 or f.last["resampled"]                # not re-normalised after a resample (D6)
        assert f.particles.poses().shape == (5000, 3)
        f.store.close()
    finally:
        config.NUM_PARTICLES, config.LANDMARK_CAPACITY, config.RNG, config.SEED = 20, 256, "device", 0


def test_reference_main_loop_on_a_synthetic_room(tmp_path):
    """scripts/slam_loop.py = jde_robots_main.py:19-62 without the simulator: every stage of the loop runs through this
    package (front-end from laser ranges, filter step, map clustering, JSON snapshot)"""
    import importlib.util
    import json
    import os
    from fast_slam_2 import config
    spec = importlib.util.spec_from_file_location("slam_loop", os.path.join(os.path.dirname(os.path.dirname(__file__)), "scripts", "slam_loop.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        out = mod.main(steps=25, particles=512, out_dir=str(tmp_path))
    finally:
        config.NUM_PARTICLES, config.LANDMARK_CAPACITY, config.RNG = 20, 256, "device"
    assert out["known_landmarks"] >= 1 and out["mean_map_size"] >= 2
    snap = json.loads((tmp_path / "fast_slam.json").read_text())
    assert len(snap["particles"]) == 512 and len(snap["landmarks"]) == out["known_landmarks"]
    assert set(snap) == {"estimated_robot_pos", "actual_robot_pos", "particles", "landmarks", "results"}
