"""Simulator glue of the reference: ``Robot`` (fast_slam_2/models/robot.py:12-151), ``EvaluationUtils``
(utils/evaluation_utils.py:10-139) and ``EvaluationResults`` (models/evaluation_results.py).  They talk to JdeRobot's
``HAL`` module, which is looked up when a method needs it (the package imports without it) -- with ``HAL`` present,
the reference's jde_robots_main.py runs against this package unmodified.  Nothing here is on the accelerated path;
``Robot.laser_message()`` hands the raw ranges to the device front-end (LandmarkUtils.get_measurements_from_laser),
``Robot.scan_environment()`` keeps the reference's points-on-the-host contract."""
from __future__ import annotations

import math
from datetime import datetime

import numpy as np

from .models import DirectedPoint


def _hal():
    try:
        import HAL
    except ImportError as e:                                         # pragma: no cover - depends on the simulator
        raise ImportError("this call talks to the JdeRobot simulator: the HAL module is not importable") from e
    return HAL


class EvaluationResults:
    """evaluation_results.py:1-48"""
    FIELDS = ("timestamp", "average_deviation", "x_deviation", "y_deviation", "angular_deviation", "distance")

    def __init__(self, timestamp, average_deviation, x_deviation, y_deviation, angular_deviation, distance):
        self.timestamp, self.average_deviation = timestamp, average_deviation
        self.x_deviation, self.y_deviation = x_deviation, y_deviation
        self.angular_deviation, self.distance = angular_deviation, distance

    def to_dict(self):
        return {k: getattr(self, k) for k in self.FIELDS}


class Robot(DirectedPoint):
    BEAMS = 180                                                      # robot.py:43

    def __init__(self, x=0.0, y=0.0, yaw=0.0):
        super().__init__(x, y, yaw)
        self._prev_timestamp = _hal().getLaserData().timeStamp       # robot.py:28
        self._prev_points = self.scan_environment()                  # robot.py:29 (kept for get_transformation_icp)

    @staticmethod
    def laser_message():
        """(values [180], minRange, maxRange) of the current laser message, for the fused device front-end"""
        msg = _hal().getLaserData()
        return np.asarray(msg.values[:Robot.BEAMS], dtype=np.float64), float(msg.minRange), float(msg.maxRange)

    @staticmethod
    def scan_environment():
        """robot.py:32-58: beams inside [minRange, maxRange] as points (dist cos a, dist sin a), a = radians(i - 90)"""
        values, lo, hi = Robot.laser_message()
        pts = [[d * math.cos(a), d * math.sin(a)]
               for d, a in zip(values.tolist(), np.radians(np.arange(Robot.BEAMS) - 90).tolist()) if not (d < lo or d > hi)]
        return np.array(pts)

    @staticmethod
    def move(lin_velocity, ang_velocity):
        """robot.py:61-90: drive straight unless a bumper is pressed, then turn away from it"""
        hal = _hal()
        bumper = hal.getBumperData()
        if bumper.state == 1:
            v, w = 0, (ang_velocity if bumper.bumper == 0 else -ang_velocity)
        else:
            v, w = lin_velocity, 0
        hal.setV(v)
        hal.setW(w)
        return v, w

    def get_transformation(self, v, w):
        """robot.py:122-151: (rotation, translation) from the commands and the laser time stamps; the simulator
        delivers 60 % of the commanded speed"""
        now = _hal().getLaserData().timeStamp
        EvaluationUtils.set_actual_pos()
        dt, self._prev_timestamp = now - self._prev_timestamp, now
        return (0, v * dt * 0.6) if v != 0 else (w * dt, 0)

    def get_transformation_icp(self, target_points, v):
        """robot.py:92-120: the same from two consecutive scans (unused by the reference's loop)"""
        from .frontend import ICP
        EvaluationUtils.set_actual_pos()
        rot, tr = ICP.get_transformation(self._prev_points, target_points)
        self._prev_points = target_points
        if v != 0:
            return 0, float(np.linalg.norm(tr))
        return -float(np.arctan2(rot[1, 0], rot[0, 0])), 0


class EvaluationUtils:
    initialized = False
    _offset = (0.0, 0.0, 0.0)
    _actual = None

    @staticmethod
    def try_to_initialize():
        """evaluation_utils.py:22-41: the simulator needs a few iterations before it reports the start pose"""
        p = _hal().getPose3d()
        if p.x < -0.5 and p.y > 0.5:
            EvaluationUtils._offset = (p.x, p.y, p.yaw)
            EvaluationUtils.initialized = True

    @staticmethod
    def set_actual_pos():
        """evaluation_utils.py:44-53: ground truth pose in the filter's frame"""
        hal = _hal()
        ox, oy, oyaw = EvaluationUtils._offset
        EvaluationUtils._actual = DirectedPoint(hal.getPose3d().x - ox, hal.getPose3d().y - oy, hal.getPose3d().yaw - oyaw)

    @staticmethod
    def evaluate_estimation(estimated_pos):
        """evaluation_utils.py:56-105: deviations in per cent (100 % = 1 m, or pi rad) and the distance"""
        act = EvaluationUtils._actual
        dx, dy = act.x - estimated_pos.x, act.y - estimated_pos.y
        x_dev, y_dev = abs(dx) * 100, abs(dy) * 100
        ang = (abs(act.yaw - estimated_pos.yaw) + np.pi) % (2 * np.pi) - np.pi
        ang_dev = abs(ang) / np.pi * 100
        res = EvaluationResults(datetime.now().strftime("%m/%d/%Y %I:%M:%S %p"), round((x_dev + y_dev + ang_dev) / 3, 2),
                                round(x_dev, 2), round(y_dev, 2), round(ang_dev, 2), round(np.sqrt(dx ** 2 + dy ** 2), 4))
        print(f"\nTimestamp: {res.timestamp}")
        print(f"Average deviation: {res.average_deviation}%")
        print(f"X deviation: {res.x_deviation}%")
        print(f"Y deviation: {res.y_deviation}%")
        print(f"Angular deviation: {res.angular_deviation}%")
        print(f"Distance between actual and estimated position: {res.distance}m")
        return res, act
