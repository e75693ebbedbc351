"""fast_slam_b200 -- B200-native FastSLAM filter step behind the API of cy-rae/fast-slam's fast_slam_2.

Only the hot path (FastSLAM2.iterate and what it calls) is implemented; see DESIGN.md.
"""
from ._lib import Fs2Error, LIB_PATH  # noqa: F401
from .store import DeviceFilter  # noqa: F401
