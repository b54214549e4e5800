"""The reference's own behavioural test cases (tests/test_pdcontroller.py,
tests/test_constraints.py, tests/test_human36_falling.py), written against the drop-in API
and run on the GPU: same models, same timelines, same assertions."""
import numpy as np
import pytest
from numpy import arange, diag, sqrt, eye

pytestmark = pytest.mark.gpu


def _batched(world, n=3):
    from arboris_b200.batch import BatchedWorld
    return BatchedWorld(world, n, device="cuda:0")


def test_pdcontroller():
    """tests/test_pdcontroller.py:8-27: the simplearm reaches gpos_des within one decimal place."""
    from arboris_b200 import controllers
    from arboris_b200.core import World, simulate
    from arboris_b200.robots.simplearm import add_simplearm
    world = World()
    add_simplearm(world)
    joints = world.getjoints()
    gpos_des = (3.14/4, 3.14/4, 3.14/4)
    kp = 7*diag((1., 1., 1.))
    world.register(controllers.ProportionalDerivativeController(
        joints, gpos_des=gpos_des, kp=kp, kd=kp/sqrt(2)))
    bw = _batched(world)
    simulate(bw, arange(0, 3, 1e-3))
    gpos = bw.gpos.cpu().numpy()
    for w in range(3):
        assert np.abs(gpos[:, w] - np.array(gpos_des)).max() < 0.05      # assertListsAlmostEqual(..., 1)


@pytest.mark.parametrize("start,check", [(3.14/2 - 0.1, lambda q: 3.14/2 >= q),
                                         (-3.14/2 + 0.1, lambda q: -3.14/2 <= q)])
def test_joint_limits(start, check):
    """tests/test_constraints.py:11-32"""
    from arboris_b200.constraints import JointLimits
    from arboris_b200.controllers import WeightController
    from arboris_b200.core import World, simulate
    from arboris_b200.robots.simplearm import add_simplearm
    world = World()
    add_simplearm(world)
    world.register(WeightController())
    shoulder = world.getjoints()['Shoulder']
    world.register(JointLimits(shoulder, -3.14/2, 3.14/2))
    shoulder.gpos[0] = start
    bw = _batched(world)
    simulate(bw, arange(0., 0.1, 1e-3))
    q = bw.gpos.cpu().numpy()[0]
    assert all(check(x) for x in q)
    assert abs(q[0] - start) > 1e-3          # it moved


def test_ball_and_socket():
    """tests/test_constraints.py:39-60: a unit mass hung from the ground; constraint force
    [0, 9.81, 0] after one update_constraints, and the body stays put."""
    from arboris_b200.constraints import BallAndSocketConstraint
    from arboris_b200.controllers import WeightController
    from arboris_b200.core import World, Body, simulate
    from arboris_b200.joints import FreeJoint
    world = World()
    b0 = Body(mass=eye(6))
    world.add_link(world.ground, FreeJoint(), b0)
    world.register(WeightController())
    world.register(BallAndSocketConstraint(frames=(world.ground, b0)))
    world.init()
    bw = _batched(world)
    dt = 0.001
    bw.update_dynamic()
    bw.update_controllers(dt)
    bw.update_constraints(dt)
    f = bw.cforce.cpu().numpy()[:3]
    assert np.abs(f - np.array([[0.], [9.81], [0.]])).max() < 1e-9
    bw.integrate(dt)
    simulate(bw, arange(0., 0.05, dt))
    assert np.abs(bw.gpos.cpu().numpy()[[3, 7, 11]]).max() < 1e-9      # translation of the 4x4 pose


def test_human36_falling():
    """tests/test_human36_falling.py:7-46: the humanoid lifted by 3 cm falls on the ground
    plane (dt = 5 ms, 0.2 s); the contact points are still above the ground."""
    from arboris_b200 import homogeneousmatrix
    from arboris_b200.constraints import get_all_contacts
    from arboris_b200.controllers import WeightController
    from arboris_b200.core import World, simulate
    from arboris_b200.robots.human36 import add_human36
    from arboris_b200.robots.simpleshapes import add_groundplane
    world = World()
    add_groundplane(world)
    add_human36(world)
    root = world.ground.childrenjoints[0]
    root.gpos = np.dot(homogeneousmatrix.transl(0, 0.03, 0), root.gpos)
    world.register(WeightController())
    contacts = get_all_contacts(world, friction_coeff=.6)
    assert len(contacts) == 8
    for c in contacts:
        world.register(c)
    world.init()
    bw = _batched(world)
    simulate(bw, arange(0., 20e-2, 5e-3))
    m = bw.model
    bw.begin_step(5e-3)          # body poses of the final state
    for k in range(8):
        body = int(m.cons_int[k][1])
        bp1 = np.asarray(m.cons_dbl[k][16:32]).reshape(4, 4)
        pose = bw.body("pose", body).cpu().numpy()
        y = np.einsum("wij,jk->wik", pose, bp1)[:, 1, 3]
        assert (y >= -1e-3).all(), (k, y)       # the reference asserts f.pose[1,3] >= 0 on its contact frames
    assert int(bw.constraints("active").sum()) >= 8


def test_parameters_edited_between_steps_are_honoured():
    """the reference reads controller / constraint parameters live at every step
    (controllers.py:141-159, core.py:913).  Retarget a PD set-point and disable a constraint between
    two steps of the single-world facade: each step must equal the oracle stepping a model flattened
    AFTER the edit (and differ from the result without the edit)."""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from arboris_b200 import scenarios
    from arboris_b200.flatten import flatten
    from oracle.arboris_oracle import OracleWorld
    w = scenarios.zoo_world()
    dt = 1e-3

    def oracle_step():
        m = flatten(w)
        o = OracleWorld(m.to_dict())
        o.step(dt)
        return o.gvel.copy()

    def facade_step():
        w.update_dynamic()
        w.update_controllers(dt)
        w.update_constraints(dt)
        w.integrate(dt)
        return w.gvel.copy()

    ref = oracle_step()
    got = facade_step()
    assert np.abs(got - ref).max() <= 1e-10*max(1., np.abs(ref).max())
    pd = [c for c in w._controllers if type(c).__name__ == "ProportionalDerivativeController"][0]
    pd.gpos_des[:] = pd.gpos_des + 0.4                      # retarget between steps
    w._constraints[0].disable()
    ref2 = oracle_step()
    got2 = facade_step()
    assert np.abs(got2 - ref2).max() <= 1e-10*max(1., np.abs(ref2).max())
    pd.gpos_des[:] = pd.gpos_des - 0.4                      # and the edit mattered
    assert np.abs(oracle_step() - ref2).max() > 1e-6
