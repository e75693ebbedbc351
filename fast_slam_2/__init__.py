"""Drop-in for the reference's ``fast_slam_2`` package, hot path only: ``from fast_slam_2 import FastSLAM2,
Measurement`` (jde_robots_main.py:4-9) resolves to the B200 implementation in ``fast_slam_b200``.
``fast_slam_2.config`` IS ``fast_slam_b200.config`` (same module object), so the reference's keys
(config.py:7-21) configure the filter.  All 15 exports of the reference (fast_slam_2/__init__.py:5-22) resolve; the two
that are glue around the simulator (Robot, EvaluationUtils) look the HAL module up when called, so importing without
it works."""
import sys as _sys

from fast_slam_b200 import config
from fast_slam_b200.filter import FastSLAM2
from fast_slam_b200.frontend import ICP, GeometryUtils, HoughTransformation, LandmarkUtils, LineFilter
from fast_slam_b200.models import DirectedPoint, Landmark, Measurement, Particle, Point
from fast_slam_b200.robot import EvaluationUtils, Robot
from fast_slam_b200.serializer import Serializer

_sys.modules[__name__ + ".config"] = config

__all__ = ["FastSLAM2", "DirectedPoint", "Landmark", "Measurement", "Particle", "Point", "config", "GeometryUtils",
           "LandmarkUtils", "LineFilter", "HoughTransformation", "Serializer", "ICP", "Robot", "EvaluationUtils"]
