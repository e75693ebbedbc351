"""CPU: the replay HAL itself (tests/replay_hal.py) and the simulator glue of the drop-in package against it."""
import sys

import numpy as np
import pytest

from tests.replay_hal import START, make_hal, record_stream


def test_recorded_stream_follows_the_reference_control_law():
    s = record_stream(400, seed=3)
    assert len(s) == 400 and all(len(f["values"]) == 180 for f in s)
    stamps = np.array([f["stamp"] for f in s])
    assert np.allclose(np.diff(stamps), 0.1)
    states = np.array([f["bumper_state"] for f in s])
    assert states[0] == 0 and states.any(), "the robot must reach a wall and turn at least once in 400 frames"
    poses = np.array([f["pose"] for f in s])
    assert poses[0, 0] < -0.5 and poses[0, 1] > 0.5                       # evaluation_utils.py:36
    # inside the room at all times (room frame = HAL frame - START)
    assert (np.abs(poses[:, 0] - START[0]) < 4.0).all() and (np.abs(poses[:, 1] - START[1]) < 3.0).all()
    moved = np.hypot(*np.diff(poses[:, :2], axis=0).T)
    assert np.allclose(moved[states[:-1] == 0], 0.3 * 0.6 * 0.1) and np.allclose(moved[states[:-1] == 1], 0.0)


def test_replay_ends_the_loop_with_stopiteration(monkeypatch):
    hal = make_hal(record_stream(3))
    monkeypatch.setitem(sys.modules, "HAL", hal)
    from fast_slam_2 import EvaluationUtils, Robot
    EvaluationUtils.initialized = False
    EvaluationUtils.try_to_initialize()
    assert EvaluationUtils.initialized
    robot = Robot()
    for k in range(3):
        v, w = robot.move(0.3, 0.5)
        assert (v, w) == (0.3, 0) and hal.state.i == k + 1
        if k < 2:
            pts = robot.scan_environment()
            assert pts.shape[1] == 2 and len(pts) > 100
            rot, tr = robot.get_transformation(v, w)
            assert rot == 0 and abs(tr - 0.3 * 0.1 * 0.6) < 1e-9          # robot.py:147
    with pytest.raises(StopIteration):
        robot.scan_environment()
