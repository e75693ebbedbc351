"""The JSON snapshot the reference's viewer reads (fast_slam_2/utils/serializer.py:36-49): same file, same keys, same
indentation.  Host-side glue -- nothing here touches the device; with ``FastSLAM2.particles`` the poses come from one
bulk read of the store instead of one Python object per particle."""
from __future__ import annotations

import json
import os


class Serializer:
    shared_path = "workspace/shared"                    # serializer.py:15-17
    file_name = "fast_slam.json"
    file_path = os.path.join(shared_path, file_name)
    # None = every particle, like the reference.  A viewer cannot draw 10^6 poses: set e.g. 5000 to write every
    # ceil(P / 5000)-th particle instead (the JSON keys do not change).
    max_particles = None

    @staticmethod
    def payload(estimated_robot_pos, actual_robot_pos, particles, landmarks, results) -> dict:
        poses = getattr(particles, "poses", None)
        if callable(poses):                             # ParticleSet: [P][3] array, no per-particle objects
            plist = [{"x": x, "y": y, "yaw": yaw} for x, y, yaw in poses(Serializer.max_particles).tolist()]
        else:
            plist = [p.to_dict() for p in particles]
        return {
            "estimated_robot_pos": estimated_robot_pos.to_dict(),
            "actual_robot_pos": actual_robot_pos.to_dict(),
            "particles": plist,
            "landmarks": [lm.to_dict() for lm in landmarks],
            "results": results.to_dict(),
        }

    @staticmethod
    def serialize(estimated_robot_pos, actual_robot_pos, particles, landmarks, results):
        data = Serializer.payload(estimated_robot_pos, actual_robot_pos, particles, landmarks, results)
        os.makedirs(Serializer.shared_path, exist_ok=True)
        with open(Serializer.file_path, "w") as fh:
            json.dump(data, fh, indent=4)
