"""CPU: the Python mirror of the reference's API surface (models, config keys) and the host-side helpers."""
import json

import numpy as np

from fast_slam_b200 import config
from fast_slam_b200.filter import _hash_uniform
from fast_slam_b200.models import DirectedPoint, Landmark, Measurement, Particle, ParticleSet, Point
from fast_slam_b200.synthetic import grid_world, synthetic_obs, synthetic_odometry, synthetic_state


def test_config_keys_and_defaults_match_the_reference():
    # fast_slam_2/config.py:7-21
    assert config.NUM_PARTICLES == 20 and config.TRANSLATION_NOISE == 0.0055 and config.ROTATION_NOISE == 0.001
    assert np.array_equal(config.MEASUREMENT_NOISE, np.array([[0.001, 0.0], [0.0, 0.001]]))
    assert config.MAXIMUM_LANDMARK_DISTANCE == 8 and config.NUM_THREAD == 20 and config.NUM_THREADS == 20


def test_models_behave_like_the_reference_classes():
    p = Point(1.0, 2.0)
    assert p.to_dict() == {"x": 1.0, "y": 2.0} and np.array_equal(p.as_vector(), [1.0, 2.0])          # point.py:18-33
    d = DirectedPoint(1.0, 2.0, 0.5)
    assert d.to_dict() == {"x": 1.0, "y": 2.0, "yaw": 0.5}                                              # directed_point.py:19-28
    l = Landmark(3.0, 4.0)
    assert np.array_equal(l.cov, [[0.1, 0.0], [0.0, 0.1]]) and l.to_dict() == {"x": 3.0, "y": 4.0}      # landmark.py:13
    l2 = Landmark(0.0, 0.0)
    l.cov[0, 0] = 9.0
    assert l2.cov[0, 0] == 0.1                      # unlike the reference's shared default array, no aliasing
    m = Measurement(2.0, 0.25)
    assert np.array_equal(m.as_vector(), [2.0, 0.25]) and (m.distance, m.yaw) == (2.0, 0.25)           # measurement.py:9-23
    q = Particle(0.0, 0.0, 0.0)
    assert q.weight == 1.0 / config.NUM_PARTICLES and q.landmarks == []                                  # particle.py:19-20


class _FakeStore:
    """What ParticleSet needs of a DeviceFilter, counting what is read."""

    def __init__(self):
        self.lm = np.zeros((3, 4, 6)); self.lm[1, 0] = (1, 2, .1, 0, 0, .1); self.lm[1, 1] = (3, 4, .2, .01, .01, .3)
        self.pose_reads, self.map_reads, self.map_particles = 0, 0, 0

    def download(self, maps=True):
        assert not maps, "a view must never pull every map"
        self.pose_reads += 1
        return dict(x=np.array([0., 1., 2.]), y=np.array([0., -1., -2.]), yaw=np.array([0., .1, .2]),
                    w=np.array([.2, .5, .3]), counts=np.array([0, 2, 1], np.int32), status=np.zeros(3, np.int32))

    def download_particles(self, sel):
        sel = np.asarray(sel)
        self.map_reads += 1
        self.map_particles += len(sel)
        return dict(counts=np.array([0, 2, 1], np.int32)[sel], lm=self.lm[sel])


def test_particle_set_view_is_a_lazy_json_serialisable_sequence():
    st = _FakeStore()
    epoch = [0]
    ps = ParticleSet(st, epoch=lambda: epoch[0])
    assert st.pose_reads == 0
    assert len(ps) == 3 and st.pose_reads == 1
    p1 = ps[1]
    assert (p1.x, p1.y, p1.yaw, p1.weight) == (1.0, -1.0, 0.1, 0.5) and isinstance(p1.x, float)
    assert len(p1.landmarks) == 2 and st.map_reads == 0                      # the map length comes with the poses
    assert p1.landmarks[1].x == 3.0 and p1.landmarks[1].cov[1, 1] == 0.3 and st.map_particles == 1
    assert p1.landmarks[-1].y == 4.0 and st.map_particles == 1               # fetched once per particle
    json.dumps([p.to_dict() for p in ps])                                     # serializer.py:39: poses only
    assert st.map_particles == 1 and st.pose_reads == 1
    assert [(l.x, l.y) for p in ps for l in p.landmarks] == [(1.0, 2.0), (3.0, 4.0), (0.0, 0.0)]      # landmark_utils.py:126-128
    assert ps.landmark_points().shape == (3, 2) and st.pose_reads == 1
    assert ps.poses().shape == (3, 3) and ps.poses(max_particles=2).shape == (2, 3)
    # a view that has to fetch something after the filter has moved on says so instead of mixing two steps
    late = ps[2]
    epoch[0] = 1
    assert late.x == 2.0
    try:
        late.landmarks[0]
        raise AssertionError("stale view served")
    except RuntimeError:
        pass


def test_serializer_reads_poses_only_and_can_decimate(tmp_path, monkeypatch):
    """jde_robots_main.py:59 calls Serializer.serialize(.., fast_slam.particles, ..) every iteration: with a
    ParticleSet that is one pose read, never the maps; Serializer.max_particles thins the list, same schema."""
    from fast_slam_2 import DirectedPoint, Landmark, Serializer

    class Results:
        def to_dict(self):
            return {"distance": 0.0}

    st = _FakeStore()
    ps = ParticleSet(st)
    monkeypatch.setattr(Serializer, "shared_path", str(tmp_path))
    monkeypatch.setattr(Serializer, "file_path", str(tmp_path / Serializer.file_name))
    Serializer.serialize(DirectedPoint(0, 0, 0), DirectedPoint(0, 0, 0), ps, [Landmark(1.0, 1.0)], Results())
    data = json.loads((tmp_path / Serializer.file_name).read_text())
    assert data["particles"] == [{"x": 0.0, "y": 0.0, "yaw": 0.0}, {"x": 1.0, "y": -1.0, "yaw": 0.1}, {"x": 2.0, "y": -2.0, "yaw": 0.2}]
    assert st.pose_reads == 1 and st.map_reads == 0
    monkeypatch.setattr(Serializer, "max_particles", 2)
    Serializer.serialize(DirectedPoint(0, 0, 0), DirectedPoint(0, 0, 0), ps, [], Results())
    data = json.loads((tmp_path / Serializer.file_name).read_text())
    assert data["particles"] == [{"x": 0.0, "y": 0.0, "yaw": 0.0}, {"x": 2.0, "y": -2.0, "yaw": 0.2}]


def test_hash_uniform_is_uniform_and_deterministic():
    u = np.array([_hash_uniform(7, s) for s in range(20000)])
    assert (u >= 0).all() and (u < 1).all() and abs(u.mean() - 0.5) < 0.01 and abs(u.std() - 12 ** -0.5) < 0.01
    assert _hash_uniform(7, 3) == _hash_uniform(7, 3) != _hash_uniform(8, 3)


def test_synthetic_workload_shapes():
    w = grid_world(256)
    assert w.shape == (256, 2) and abs(w.mean()) < 1e-12 and np.isclose(w[1, 0] - w[0, 0], 1.5)
    s = synthetic_state(1, 50, 64, 80)
    assert s["lm"].shape == (50, 80, 6) and (s["count"] == 64).all() and np.isclose(s["w"].sum(), 1.0)
    assert (s["lm"][:, :64, 2] > 0.002).all() and (s["lm"][:, 64:] == 0).all()
    o = synthetic_obs(1, 0, w, 32, novel=4)
    assert o.shape == (32, 2) and (o[:, 0] > 0).all()
    # observations are of distinct landmarks, novel ones sit in cell centres >= 1 m from every landmark
    xy = np.stack([o[:, 0] * np.cos(o[:, 1]), o[:, 0] * np.sin(o[:, 1])], 1)
    d = np.linalg.norm(xy[:, None] - w[None], axis=2).min(1)
    assert (d[:28] < 0.45).all() and (d[28:] > 0.8).all()
    assert synthetic_odometry(3) == (0.0, 0.0) and synthetic_odometry(9) == (0.001, 0.0) and synthetic_odometry(19) == (-0.001, 0.0)


def test_serializer_writes_the_reference_schema(tmp_path, monkeypatch):
    """serializer.py:36-49: keys, nesting and indent of the viewer's JSON file"""
    import json
    from fast_slam_2 import DirectedPoint, Landmark, Particle, Serializer

    class Results:
        def to_dict(self):
            return {"timestamp": "t", "average_deviation": 0.5, "x_deviation": 0.1, "y_deviation": 0.2,
                    "angular_deviation": 0.3, "distance": 0.4}

    monkeypatch.setattr(Serializer, "shared_path", str(tmp_path))
    monkeypatch.setattr(Serializer, "file_path", str(tmp_path / Serializer.file_name))
    parts = [Particle(1.0, 2.0, 0.5), Particle(-1.0, 0.0, 3.0)]
    Serializer.serialize(DirectedPoint(0.1, 0.2, 0.3), DirectedPoint(0.0, 0.0, 0.0), parts, [Landmark(4.0, 5.0)], Results())
    text = (tmp_path / Serializer.file_name).read_text()
    data = json.loads(text)
    assert list(data) == ["estimated_robot_pos", "actual_robot_pos", "particles", "landmarks", "results"]
    assert data["estimated_robot_pos"] == {"x": 0.1, "y": 0.2, "yaw": 0.3}
    assert data["particles"] == [{"x": 1.0, "y": 2.0, "yaw": 0.5}, {"x": -1.0, "y": 0.0, "yaw": 3.0}]
    assert data["landmarks"] == [{"x": 4.0, "y": 5.0}] and data["results"]["distance"] == 0.4
    assert text.startswith('{\n    "estimated_robot_pos": {\n        "x": 0.1')        # indent = 4


def _stub_hal(monkeypatch, values, lo=0.3, hi=6.0, stamp=0.0, pose=(0.0, 0.0, 0.0), bumper=(0, 0)):
    import sys
    import types
    hal = types.ModuleType("HAL")
    state = {"v": None, "w": None, "stamp": stamp, "pose": pose, "bumper": bumper}
    hal.getLaserData = lambda: types.SimpleNamespace(values=list(values), minRange=lo, maxRange=hi, timeStamp=state["stamp"])
    hal.getPose3d = lambda: types.SimpleNamespace(x=state["pose"][0], y=state["pose"][1], yaw=state["pose"][2])
    hal.getBumperData = lambda: types.SimpleNamespace(state=state["bumper"][0], bumper=state["bumper"][1])
    hal.setV = lambda v: state.__setitem__("v", v)
    hal.setW = lambda w: state.__setitem__("w", w)
    monkeypatch.setitem(sys.modules, "HAL", hal)
    return state


def test_robot_and_evaluation_utils_against_a_stub_simulator(monkeypatch, capsys):
    """models/robot.py:12-151 and utils/evaluation_utils.py:10-139 through a stub HAL: the scan points are the ones
    the reference's Robot.scan_environment produced for the recorded laser messages (frontend_polar_kats.npz)"""
    import numpy as np
    from fast_slam_2 import DirectedPoint, EvaluationUtils, Robot
    from tests.util import load_golden
    g = load_golden("frontend_polar_kats.npz")
    st = _stub_hal(monkeypatch, g["values"][0], float(g["min_range"]), float(g["max_range"]), stamp=10.0, pose=(-1.0, 2.0, 0.5))
    EvaluationUtils.initialized = False
    EvaluationUtils.try_to_initialize()
    assert EvaluationUtils.initialized                                   # x < -0.5 and y > 0.5
    robot = Robot()
    np.testing.assert_array_equal(robot.scan_environment(), g["pts"][0, :g["npts"][0]])
    vals, lo, hi = Robot.laser_message()
    assert len(vals) == 180 and (lo, hi) == (float(g["min_range"]), float(g["max_range"]))
    assert robot.move(0.3, 0.5) == (0.3, 0) and (st["v"], st["w"]) == (0.3, 0)
    st["bumper"] = (1, 0)
    assert robot.move(0.3, 0.5) == (0, 0.5)                              # right bumper: turn left
    st["bumper"] = (1, 2)
    assert robot.move(0.3, 0.5) == (0, -0.5)
    st["stamp"] = 10.5
    assert robot.get_transformation(0.3, 0) == (0, 0.3 * 0.5 * 0.6)      # translation = v * dt * 0.6
    st["stamp"] = 10.75
    assert robot.get_transformation(0, 0.5) == (0.5 * 0.25, 0)
    st["pose"] = (-0.5, 2.25, 0.6)                                       # moved by (0.5, 0.25, 0.1) since the start
    EvaluationUtils.set_actual_pos()
    res, actual = EvaluationUtils.evaluate_estimation(DirectedPoint(0.4, 0.25, 0.0))
    assert abs(actual.x - 0.5) < 1e-12 and abs(actual.y - 0.25) < 1e-12 and abs(actual.yaw - 0.1) < 1e-12
    assert res.x_deviation == 10.0 and res.y_deviation == 0.0 and res.distance == 0.1
    assert res.angular_deviation == round(0.1 / np.pi * 100, 2)
    assert set(res.to_dict()) == {"timestamp", "average_deviation", "x_deviation", "y_deviation", "angular_deviation", "distance"}
    assert "Average deviation" in capsys.readouterr().out


def test_package_exports_match_the_reference():
    import fast_slam_2
    names = {"FastSLAM2", "HoughTransformation", "ICP", "LineFilter", "DirectedPoint", "Landmark", "Measurement", "Particle",
             "Point", "Robot", "EvaluationUtils", "GeometryUtils", "LandmarkUtils", "Serializer"}      # __init__.py:5-22
    assert names <= set(dir(fast_slam_2)) and hasattr(fast_slam_2, "config")
