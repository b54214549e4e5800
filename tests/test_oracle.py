"""The oracle (oracle/arboris_oracle.py, numpy restatement of the reference step)
against (1) the reference's own known-answer vectors and (2) trajectories
recorded from the REAL reference (tests/golden, made by oracle/make_goldens.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle.arboris_oracle import OracleWorld, twist_exp, zaligned, adjoint
from arboris_b200 import scenarios
from arboris_b200.flatten import flatten


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max()/max(np.abs(b).max(), 1e-300)


def test_update_dynamic_known_answers():
    """reference tests/test_update_dynamic.py:10-19, 117-120, 195-198"""
    w = scenarios.simplearm_world()
    m = flatten(w)
    o = OracleWorld(m.to_dict())
    o.gpos[:] = (0.5, 1.0, 2.0/3.0)
    o.gvel[:] = (2.5, -1.0, -0.5)
    o.update_dynamic()
    np.testing.assert_allclose(o.mass, [[0.55132061, 0.1538999, 0.0080032],
                                        [0.1538999, 0.09002086, 0.00896043],
                                        [0.0080032, 0.00896043, 0.00267333]], atol=5e-8)
    np.testing.assert_allclose(o.pose[3], [[-0.56122931, -0.82766035, 0., -0.63871076],
                                           [0.82766035, -0.56122931, 0., 0.46708616],
                                           [0., 0., 1., 0.], [0., 0., 0., 1.]], atol=5e-8)
    np.testing.assert_allclose(o.twist[2], [0., 0., 1.5, -0.67537788, 1.05183873, 0.], atol=5e-8)
    np.testing.assert_allclose(o.jac[3], [[0, 0, 0], [0, 0, 0], [1, 1, 1],
                                          [-0.26649313, -0.3143549, 0.],
                                          [0.7450519, 0.24734792, 0.], [0, 0, 0]], atol=5e-8)
    assert np.abs(o.viscosity).max() == 0.


def test_human36_mass_diagonal():
    """reference tests/test_human36.rst:98-114 (checked against HuMAnS)"""
    o = OracleWorld(flatten(scenarios.human36_free_world()).to_dict())
    o.update_dynamic()
    expect = {5: 73.000000000000014, 41: 0.10208399155688053, 40: 0.020356790291165189,
              39: 0.10430013572386694, 16: 0.0093741757009949949, 17: 0.001397215796713388,
              10: 0.0093741757009949949, 11: 0.001397215796713388}
    for i, v in expect.items():
        assert abs(o.mass[i, i] - v) <= 1e-13*max(1., abs(v))


def test_human36_body_masses_h5():
    """reference tests/test_human36.py:93-115 + tests/human36.h5 (converted)"""
    ref = np.load(os.path.join(GOLDEN, "reference_h5.npz"))
    w = scenarios.human36_free_world()
    for b in list(w.iterbodies())[1:]:
        np.testing.assert_allclose(b.mass, ref["human36/masses/" + b.name], atol=1e-14)


def test_simplearm_h5_trajectory():
    """reference tests/simplearm_flat.h5 (recipe tests/test_visu_collada.py:11-27):
    99 steps at dt=1e-2, absolute body poses before each integrate."""
    ref = np.load(os.path.join(GOLDEN, "reference_h5.npz"))
    o = OracleWorld(flatten(scenarios.simplearm_world()).to_dict())
    t = np.arange(0, 1, 0.01)
    cur = t[0]
    for s, nxt in enumerate(t[1:]):
        dt = nxt - cur
        o.update_dynamic()
        o.update_controllers(dt)
        o.update_constraints(dt)
        for b, name in ((1, "Arm"), (2, "Forearm"), (3, "Hand")):
            assert np.abs(o.pose[b] - ref["simplearm_flat/transforms/" + name][s]).max() < 1e-12
        o.integrate(dt)
        cur += dt


def test_impedance_admittance_doctest():
    """reference core.py:746-761 (simplearm + PD controller on the elbow, kp = 2)"""
    from arboris_b200 import World
    from arboris_b200.robots.simplearm import add_simplearm
    from arboris_b200.controllers import ProportionalDerivativeController
    w = World()
    add_simplearm(w)
    joints = w.getjoints()
    w.register(ProportionalDerivativeController(joints[1:2], 2.))
    w.init()
    o = OracleWorld(flatten(w).to_dict())
    o.update_dynamic()
    o.update_controllers(0.001)
    np.testing.assert_allclose(o.impedance, [[686.98833333, 223.44666667, 20.67333333],
                                             [223.44666667, 93.44866667, 10.67333333],
                                             [20.67333333, 10.67333333, 2.67333333]], atol=5e-8)
    np.testing.assert_allclose(o.admittance, [[0.00732382, -0.0203006, 0.0244142],
                                              [-0.0203006, 0.07594182, -0.14621124],
                                              [0.0244142, -0.14621124, 0.76901683]], atol=5e-8)


def test_ball_and_socket_known_answer():
    """reference tests/test_constraints.py:39-60: force [0, 9.81, 0], pose stays identity"""
    o = OracleWorld(flatten(scenarios.ball_socket_world()).to_dict())
    dt = 0.001
    o.update_dynamic()
    o.update_controllers(dt)
    o.update_constraints(dt)
    np.testing.assert_allclose(o.cforce, [0., 9.81, 0.], atol=1e-7)
    o.integrate(dt)
    o.update_dynamic()
    np.testing.assert_allclose(o.pose[1], np.eye(4), atol=1e-7)


def test_math_doctests():
    """twistvector.py:41-47, homogeneousmatrix.py:211-215, :288-305"""
    np.testing.assert_allclose(
        twist_exp(np.array([1., 2., 3., 10., 11., 12.])),
        [[-0.69492056, 0.71352099, 0.08929286, 2.90756949],
         [-0.19200697, -0.30378504, 0.93319235, 11.86705709],
         [0.69297817, 0.6313497, 0.34810748, 13.78610544], [0, 0, 0, 1]], atol=5e-8)
    H, idx = zaligned((1., 0., 0.))
    np.testing.assert_allclose(H, [[0, 0, 1, 0], [0, -1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1]], atol=0)
    assert list(idx) == [1, 2, 0]
    H = np.array([[0.70738827, 0., -0.70682518, 3.], [0.61194086, 0.50045969, 0.61242835, 4.],
                  [0.35373751, -0.86575984, 0.35401931, 5.], [0., 0., 0., 1.]])
    np.testing.assert_allclose(adjoint(H)[3], [-1.64475426, -5.96533781, -1.64606451,
                                               0.70738827, 0., -0.70682518], atol=5e-8)


def _teacher_forced(name, stepper_factory, tol):
    """Run a stepper from the reference's state at every step and compare."""
    model, tr = load_golden(name)
    dt = float(tr["dt"])
    W, T = tr["gpos"].shape[:2]
    fs = list(tr["full_steps"])
    worst, flips = {}, 0
    for wi in range(W):
        o = stepper_factory(model)
        gpos, gvel, cf = tr["gpos_in"][wi], tr["gvel_in"][wi], np.zeros(model.nrows)
        for s in range(T):
            o.gpos[:], o.gvel[:], o.cforce[:] = gpos, gvel, cf
            o.update_dynamic()
            o.update_controllers(dt)
            if s in fs:
                i = fs.index(s)
                for k in ("mass", "nleffects", "impedance", "admittance"):
                    worst[k] = max(worst.get(k, 0), rel(getattr(o, k), tr[k][wi, i]))
            o.update_constraints(dt)
            if model.nc:
                a = np.array(o.active, dtype=np.int8)
                flips += int((a != tr["active"][wi, s]).sum())
                flips += int((np.array(o.branch)*a != tr["branch"][wi, s]).sum())
                worst["cforce"] = max(worst.get("cforce", 0), rel(o.cforce, tr["cforce"][wi, s]))
            o.integrate(dt)
            worst["gvel"] = max(worst.get("gvel", 0), rel(o.gvel, tr["gvel"][wi, s]))
            worst["gpos"] = max(worst.get("gpos", 0), rel(o.gpos, tr["gpos"][wi, s]))
            gpos, gvel, cf = tr["gpos"][wi, s], tr["gvel"][wi, s], tr["cforce"][wi, s]
    assert flips == 0, "active-set / branch flips vs the real reference: %d" % flips
    for k, v in worst.items():
        assert v <= tol, (k, v)


@pytest.mark.parametrize("name", ["simplearm", "human36_free", "ball_socket",
                                  "simplearm_limits", "snake_loop", "human36_contact", "balls", "zoo"])
def test_oracle_matches_real_reference(name):
    """The restatement reproduces the real reference bit-for-bit (same numpy calls)."""
    _teacher_forced(name, lambda m: OracleWorld(m.to_dict()), 1e-13)


def test_oracle_free_running_simplearm():
    """configs[0]: 1000 steps at dt=1e-3, free running, final state vs real reference."""
    model, tr = load_golden("simplearm")
    o = OracleWorld(model.to_dict())
    o.gpos[:], o.gvel[:] = tr["gpos_in"][0], tr["gvel_in"][0]
    for s in range(1000):
        o.step(1e-3)
    assert np.abs(o.gpos - tr["gpos"][0, -1]).max() < 1e-12


def _extra(model_name, file_name):
    from arboris_b200.flatten import FlatModel
    model = FlatModel.load(os.path.join(GOLDEN, "model_%s.npz" % model_name))
    with np.load(os.path.join(GOLDEN, file_name)) as z:
        return model, {k: z[k] for k in z.files}


def test_oracle_vs_contact64_sample():
    """The 64-world contact fixture (real reference, free running): the oracle from the reference's
    state at every step, on a sample of the worlds (the whole fixture runs on the device)."""
    model, tr = _extra("human36_contact", "traj_human36_contact64.npz")
    dt = float(tr["dt"])
    flips, worst = 0, 0.
    for wi in (0, 17, 38, 63):
        o = OracleWorld(model.to_dict())
        gpos, gvel, cf = tr["gpos_in"][wi], tr["gvel_in"][wi], np.zeros(model.nrows)
        for s in range(tr["gpos"].shape[1]):
            o.gpos[:], o.gvel[:], o.cforce[:] = gpos, gvel, cf
            o.step(dt)
            a = np.array(o.active, dtype=np.int8)
            flips += int((a != tr["active"][wi, s]).sum()) + int((np.array(o.branch)*a != tr["branch"][wi, s]).sum())
            worst = max(worst, rel(o.gvel, tr["gvel"][wi, s]), rel(o.gpos, tr["gpos"][wi, s]),
                        rel(o.cforce, tr["cforce"][wi, s]))
            gpos, gvel, cf = tr["gpos"][wi, s], tr["gvel"][wi, s], tr["cforce"][wi, s]
    assert flips == 0 and worst <= 1e-13, (flips, worst)


@pytest.mark.parametrize("name,nworlds", [("human36_free", 1), ("snake_loop", 2)])
def test_oracle_free_running_1000_steps(name, nworlds):
    """1000 free-running steps of the oracle against checkpoints of the real reference."""
    model, tr = _extra(name, "free_%s.npz" % name)
    for wi in range(nworlds):
        o = OracleWorld(model.to_dict())
        o.gpos[:], o.gvel[:] = tr["gpos_in"][wi], tr["gvel_in"][wi]
        done = 0
        for i, k in enumerate(tr["kept_steps"]):
            for _ in range(int(k) + 1 - done):
                o.step(float(tr["dt"]))
            done = int(k) + 1
            assert np.abs(o.gpos - tr["gpos"][wi, i]).max() < 1e-9, (name, wi, int(k))


def test_model_snapshot_follows_parameter_changes():
    """the reference reads constraint / controller parameters live every step; the
    single-world facade rebuilds its device model when one of them changes (and only then)."""
    w = scenarios.zoo_world()
    m0 = flatten(w)
    for j in w.iterjoints():                    # moving the world is not a parameter change
        if hasattr(j.gpos, "shape") and j.gpos.ndim == 1:
            j.gpos[:] = j.gpos + 0.01
    assert m0.same_parameters(flatten(w))
    pd = [c for c in w._controllers if type(c).__name__ == "ProportionalDerivativeController"][0]
    pd.gpos_des[0] += 0.5
    assert not m0.same_parameters(flatten(w))
    m1 = flatten(w)
    w._constraints[0].disable()
    assert not m1.same_parameters(flatten(w))
