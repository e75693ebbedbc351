"""Drop-in for the reference's ``fast_slam_2`` package, hot path only: ``from fast_slam_2 import FastSLAM2,
Measurement`` (jde_robots_main.py:4-9) resolves to the B200 implementation in ``fast_slam_b200``.
``fast_slam_2.config`` IS ``fast_slam_b200.config`` (same module object), so the reference's keys
(config.py:7-21) configure the filter.  Of the reference's 15 exports (fast_slam_2/__init__.py:5-22) the two that
are glue around the simulator module HAL (Robot, EvaluationUtils -- DESIGN.md section 7) are not provided; importing without the
simulator module HAL works."""
import sys as _sys

from fast_slam_b200 import config
from fast_slam_b200.filter import FastSLAM2
from fast_slam_b200.frontend import ICP, GeometryUtils, HoughTransformation, LandmarkUtils, LineFilter
from fast_slam_b200.models import DirectedPoint, Landmark, Measurement, Particle, Point
from fast_slam_b200.serializer import Serializer

_sys.modules[__name__ + ".config"] = config

_OUT_OF_SCOPE = {"Robot", "EvaluationUtils"}


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise ImportError("fast_slam_2.%s is simulator glue outside the accelerated path (DESIGN.md section 7); "
                          "use the reference's own module for it" % name)
    raise AttributeError(name)


__all__ = ["FastSLAM2", "DirectedPoint", "Landmark", "Measurement", "Particle", "Point", "config", "GeometryUtils",
           "LandmarkUtils", "LineFilter", "HoughTransformation", "Serializer", "ICP"]
