"""Configuration of the filter.  The first six names and their defaults are the reference's
(fast_slam_2/config.py:7-21); unlike the reference, which binds them at import, they are read when a
FastSLAM2 is constructed, so `config.NUM_PARTICLES = 1 << 20` before `FastSLAM2()` works."""
import numpy as np

NUM_PARTICLES = 20                                               # config.py:7
TRANSLATION_NOISE = 0.0055                                       # config.py:11
ROTATION_NOISE = 0.001                                           # config.py:12
MEASUREMENT_NOISE = np.array([[0.001, 0.0], [0.0, 0.001]])       # config.py:15
MAXIMUM_LANDMARK_DISTANCE = 8                                    # config.py:18
NUM_THREAD = 20                                                  # config.py:21 -- kept for compatibility, unused:
NUM_THREADS = NUM_THREAD                                         # the particle loop is a CUDA grid (README.md:63 spelling)

# ---- additions of the B200 implementation -------------------------------------------------------
LANDMARK_CAPACITY = 256      # map slots per particle (the reference's lists are unbounded, quirk Q16)
DEVICE = None                # CUDA device ordinal; None = torch's current device
SEED = 0                     # key of the device random generator
# "device": motion noise from the counter-based device generator, resampling start from a host hash.
# "reference": both drawn from the global np.random in the reference's order (fast_slam_2.py:79/81,183),
#              so that np.random.seed(s) reproduces the reference's trajectory draw for draw (quirk Q15).
RNG = "device"
