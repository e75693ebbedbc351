"""A stand-in for JdeRobot's ``HAL`` module that replays a recorded stream (row N3 of SURVEY.md 8f).

The stream is recorded up front by driving a kinematic robot through the synthetic 8 m x 6 m room with the control
law of the reference's ``Robot.move`` (models/robot.py:61-90: straight ahead unless a bumper is pressed, then turn
away) -- the law the replayed loop will apply again, so commands and frames stay consistent.  A frame holds what the
simulator would report at one instant: laser message (180 ranges, min/max range, time stamp), ground-truth pose,
bumper.  ``setW`` -- the last HAL call of ``Robot.move`` -- advances to the next frame; when the stream is used up
the next read raises ``StopIteration``, which is how a test ends the reference's ``while True`` loop.
"""
from __future__ import annotations

import types

import numpy as np

from fast_slam_b200.synthetic import room_ranges

START = (-1.0, 2.0, 0.0)          # HAL frame: x < -0.5 and y > 0.5 lets EvaluationUtils initialise (evaluation_utils.py:36)


def record_stream(frames: int, dt: float = 0.1, v_cmd: float = 0.3, w_cmd: float = 0.5, seed: int = 0):
    """frames x dict(stamp, values[180], pose(x, y, yaw) in the HAL frame, bumper_state, bumper)."""
    angles = np.radians(np.arange(180) - 90)                     # robot.py:52
    x = y = yaw = 0.0                                            # room frame: starts in the middle of the room
    out = []
    turning = 0
    for k in range(frames):
        values = room_ranges(angles, (x, y, yaw), seed=seed * 100003 + k)
        ahead = float(np.min(values[80:101]))
        # a pressed bumper stays pressed until the way ahead is clear again
        if ahead < 0.45:
            turning = 1
        elif turning and ahead > 1.2:
            turning = 0
        out.append(dict(stamp=round(10.0 + k * dt, 6), values=values, pose=(START[0] + x, START[1] + y, START[2] + yaw),
                        bumper_state=turning, bumper=1))
        if turning:                                              # Robot.move: v = 0, w = -ang_velocity (centre bumper)
            yaw = (yaw - w_cmd * dt + np.pi) % (2 * np.pi) - np.pi
        else:                                                    # the simulator delivers 60 % of the commanded speed (robot.py:147)
            x += v_cmd * 0.6 * dt * np.cos(yaw)
            y += v_cmd * 0.6 * dt * np.sin(yaw)
    return out


def make_hal(stream, min_range: float = 0.1, max_range: float = 10.0):
    """A module object named HAL over ``stream``; ``hal.state`` exposes the cursor and the commands received."""
    hal = types.ModuleType("HAL")
    st = types.SimpleNamespace(i=0, v=[], w=[], laser_reads=0)
    hal.state = st

    def frame():
        if st.i >= len(stream):
            raise StopIteration("replay stream exhausted after %d frames" % len(stream))
        return stream[st.i]

    def get_laser():
        f = frame()
        st.laser_reads += 1
        return types.SimpleNamespace(values=[float(v) for v in f["values"]], minRange=min_range, maxRange=max_range, timeStamp=f["stamp"])

    def get_pose():
        f = frame()
        return types.SimpleNamespace(x=f["pose"][0], y=f["pose"][1], yaw=f["pose"][2])

    def get_bumper():
        f = frame()
        return types.SimpleNamespace(state=f["bumper_state"], bumper=f["bumper"])

    def set_v(v):
        st.v.append(v)

    def set_w(w):
        st.w.append(w)
        st.i += 1                                                # Robot.move is done: the world moves on

    hal.getLaserData, hal.getPose3d, hal.getBumperData, hal.setV, hal.setW = get_laser, get_pose, get_bumper, set_v, set_w
    return hal
