"""The loop of the reference's jde_robots_main.py:19-62 on this package, with a synthetic room instead of the
simulator: laser message -> LandmarkUtils.get_measurements_from_laser (device front-end) -> FastSLAM2.iterate ->
LandmarkUtils.update_known_landmarks (device map clustering) -> Serializer.serialize.  The drive follows the
reference's Robot.move / get_transformation (robot.py:61-151): straight at v = 0.3 m/s (x 0.6, dt = 0.1 s), a turn of
0.05 rad every tenth step.

    python scripts/slam_loop.py [steps] [particles] [out_dir]
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Results:                                  # stands in for EvaluationUtils.evaluate_estimation's results
    def __init__(self, est, actual):
        self.d = {"timestamp": time.strftime("%m/%d/%Y %I:%M:%S %p"), "distance": round(float(np.hypot(est[0] - actual[0], est[1] - actual[1])), 4)}

    def to_dict(self):
        return self.d


def main(steps=60, particles=4096, out_dir=None, quiet=True):
    from fast_slam_2 import DirectedPoint, FastSLAM2, LandmarkUtils, Serializer, config
    from fast_slam_b200.synthetic import room_ranges
    config.NUM_PARTICLES, config.LANDMARK_CAPACITY, config.RNG = particles, 64, "device"
    saved_paths = (Serializer.shared_path, Serializer.file_path)
    if out_dir:
        Serializer.shared_path = out_dir
        Serializer.file_path = os.path.join(out_dir, Serializer.file_name)
    fast_slam = FastSLAM2()
    angles = np.radians(np.arange(180) - 90)     # robot.py:52
    x = y = yaw = 0.0                            # the simulated robot
    ex = ey = eyaw = 0.0                         # dead reckoning, as the reference does for its first 150 iterations
    t_front = t_iter = t_known = 0.0
    known_ms = []
    sink = io.StringIO() if quiet else sys.stdout
    with contextlib.redirect_stdout(sink):
        for s in range(steps):
            if s % 10 == 9:
                rotation, translation = 0.05, 0.0
                yaw += rotation
            else:
                rotation, translation = 0.0, 0.3 * 0.1 * 0.6
                x += translation * np.cos(yaw); y += translation * np.sin(yaw)
            values = room_ranges(angles, (x, y, yaw), seed=s)
            t0 = time.perf_counter()
            measurements = LandmarkUtils.get_measurements_from_laser(values, 0.1, 10.0)
            t1 = time.perf_counter()
            est = fast_slam.iterate(rotation, translation, measurements)
            t2 = time.perf_counter()
            LandmarkUtils.update_known_landmarks(fast_slam.particles)
            t3 = time.perf_counter()
            t_front += t1 - t0; t_iter += t2 - t1; t_known += t3 - t2
            known_ms.append(1e3 * (t3 - t2))
            eyaw = (eyaw + rotation + np.pi) % (2 * np.pi) - np.pi
            ex += translation * np.cos(eyaw); ey += translation * np.sin(eyaw)
        if out_dir:
            Serializer.serialize(DirectedPoint(ex, ey, eyaw), DirectedPoint(x, y, yaw), fast_slam.particles,
                                 LandmarkUtils.known_landmarks, _Results(est, (x, y)))
    out = {"steps": steps, "particles": particles, "frontend_ms_per_scan": 1e3 * t_front / steps,
           "iterate_ms_per_step": 1e3 * t_iter / steps, "update_known_landmarks_ms_per_call": 1e3 * t_known / steps,
           "update_known_landmarks_ms_median": float(np.median(known_ms)),      # the mean carries the workspace (re)allocations
           "known_landmarks": len(LandmarkUtils.known_landmarks), "estimate": [float(v) for v in est],
           "robot": [x, y, yaw], "mean_map_size": float(np.mean(fast_slam.store.count.cpu().numpy()))}
    fast_slam.store.close()
    Serializer.shared_path, Serializer.file_path = saved_paths
    return out


if __name__ == "__main__":
    a = sys.argv[1:]
    print(json.dumps(main(int(a[0]) if a else 60, int(a[1]) if len(a) > 1 else 4096, a[2] if len(a) > 2 else None)))
