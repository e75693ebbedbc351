"""Times the UNMODIFIED reference (cy-rae/fast-slam, /root/reference, HAL stubbed) on BASELINE.json config 1:
20 particles, NUM_THREAD = 20, room drive with 360-beam scans -> get_measurements_to_landmarks -> iterate ->
update_known_landmarks (the loop of jde_robots_main.py:28-59 without the simulator).  Runs only where the reference
tree exists (the build container); the output is committed under profiles/ as the reference's own numbers next to the
C restatement that bench.py times on the GPU box.

    python scripts/time_reference_cfg1.py [steps] > profiles/r01_reference_python_cfg1.json
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
from oracle import ref_harness as rh          # noqa: E402
from fast_slam_b200.synthetic import room_scan  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    ref = rh.load_reference()
    np.random.seed(0)
    flt = ref.FastSLAM2()                       # unmodified config: NUM_PARTICLES = 20, NUM_THREAD = 20
    P = len(flt.particles)
    x = y = yaw = 0.0
    t_front = t_iter = t_known = 0.0
    updates = 0
    with contextlib.redirect_stdout(io.StringIO()):
        for s in range(steps):
            if s % 10 == 9:
                rot, tr = 0.05, 0.0
                yaw += rot
            else:
                rot, tr = 0.0, 0.018
                x += tr * np.cos(yaw); y += tr * np.sin(yaw)
            pts = room_scan(360, 2 * np.pi, (x, y, yaw), seed=s)
            t0 = time.perf_counter()
            meas = ref.LandmarkUtils.get_measurements_to_landmarks(pts)
            t1 = time.perf_counter()
            flt.iterate(rot, tr, meas)
            t2 = time.perf_counter()
            ref.LandmarkUtils.update_known_landmarks(flt.particles)
            t3 = time.perf_counter()
            t_front += t1 - t0; t_iter += t2 - t1; t_known += t3 - t2
            updates += P * len(meas)
    print(json.dumps({
        "impl": "reference (python, unmodified)", "config": "cfg1: 20 particles, 360-beam room scans, %d steps" % steps,
        "host_cores": os.cpu_count(), "particles": P, "observations_total": updates // P,
        "iterate_ms_per_step": 1e3 * t_iter / steps, "frontend_ms_per_scan": 1e3 * t_front / steps,
        "update_known_landmarks_ms_per_call": 1e3 * t_known / steps,
        "particle_observation_updates_per_s": updates / t_iter,
        "landmarks_per_particle_at_end": float(np.mean([len(p.landmarks) for p in flt.particles])),
    }))


if __name__ == "__main__":
    main()
