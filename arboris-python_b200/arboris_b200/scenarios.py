"""The benchmark / parity scenarios of BASELINE.json (``configs[0..4]``) and their
seeded synthetic initial states.

Builders take the module namespace to build with (``api``): this package by
default, or the real reference when ``oracle/make_goldens.py`` generates the
fixtures -- the same function therefore defines the scenario for both sides.
Initial states are numpy on the host; the GPU path uploads them.
"""
import os

import numpy as np

from . import homogeneousmatrix as Hg

SEED0 = 20260000
SNAKE_LENGTHS = [1., .9, .8, .7, .6, .5, .4, .3, .2]   # matlab/test_snake.py:47-58
SNAKE_QREF = 0.25                                     # arc the loop is closed on
KNEE_LIMITS = (-2.4, 0.05)


def snake_tip(q):
    """Tip position of the planar snake (base at the origin) for hinge angles q."""
    th = np.cumsum(q)
    L = np.asarray(SNAKE_LENGTHS)
    return float(-(L*np.sin(th)).sum()), float((L*np.cos(th)).sum())


class _Api(object):
    """Names a scenario needs, resolved from this package or the reference."""

    def __init__(self, reference=False):
        if reference:
            import arboris.core as core
            import arboris.constraints as constraints
            import arboris.controllers as controllers
            import arboris.shapes as shapes
            import arboris.joints as joints
            import arboris.massmatrix as massmatrix
            from arboris.robots import human36, simplearm, snake, simpleshapes
        else:
            from . import core, constraints, controllers, shapes, joints, massmatrix
            from .robots import human36, simplearm, snake, simpleshapes
        self.core, self.constraints, self.controllers = core, constraints, controllers
        self.shapes, self.joints, self.massmatrix = shapes, joints, massmatrix
        self.human36, self.simplearm, self.snake = human36, simplearm, snake
        self.simpleshapes = simpleshapes


def simplearm_world(reference=False):
    """configs[0]: 3-dof planar arm under gravity, shoulder at 3.14/4
    (the recipe of tests/test_visu_collada.py:11-27)."""
    a = _Api(reference)
    w = a.core.World()
    w.register(a.controllers.WeightController())
    a.simplearm.add_simplearm(w)
    w.getjoints()['Shoulder'].gpos[0] = 3.14/4
    w.init()
    return w


def human36_free_world(reference=False):
    """configs[1]: free-floating humanoid under gravity, no constraints."""
    a = _Api(reference)
    w = a.core.World()
    a.human36.add_human36(w)
    w.register(a.controllers.WeightController())
    w.init()
    return w


def human36_contact_world(reference=False):
    """configs[2] and [4]: humanoid over a ground plane, 8 SoftFingerContact
    (mu = .6, tests/test_human36_falling.py:13-38) then JointLimits on the two
    1-dof knees."""
    a = _Api(reference)
    w = a.core.World()
    a.simpleshapes.add_groundplane(w)
    a.human36.add_human36(w)
    w.register(a.controllers.WeightController())
    for c in a.constraints.get_all_contacts(w, friction_coeff=.6):
        w.register(c)
    joints = list(w.iterjoints())
    for k in (2, 5):   # ShankR, ShankL knees: dofs 9 and 15
        assert joints[k].ndof == 1
        w.register(a.constraints.JointLimits(joints[k], *KNEE_LIMITS))
    w.init()
    return w


def snake_loop_world(reference=False):
    """configs[3]: free 9-link snake whose base and tip are pinned to the
    ground by two BallAndSocketConstraint (a kinematic loop)."""
    a = _Api(reference)
    w = a.core.World()
    n = len(SNAKE_LENGTHS)
    a.snake.add_snake(w, n, lengths=list(SNAKE_LENGTHS), masses=list(SNAKE_LENGTHS),
                      gpos=[0.]*n, gvel=[0.]*n, is_fixed=False)
    w.register(a.controllers.WeightController())
    joints = list(w.iterjoints())
    base = joints[0]._frame1.body
    tip = w._subframes[-1]          # the frame add_snake registers last
    # the loop is closed on the arc q_i = SNAKE_QREF (a straight chain between two
    # pins would be a singular configuration)
    tx, ty = snake_tip([SNAKE_QREF]*n)
    tip_pose0 = Hg.transl(tx, ty, 0.)
    w.register(a.constraints.BallAndSocketConstraint(frames=(w.ground, base)))
    w.register(a.constraints.BallAndSocketConstraint(
        frames=(a.core.SubFrame(w.ground, tip_pose0), tip)))
    w.init()
    return w


def ball_socket_world(reference=False):
    """tests/test_constraints.py:39-47: a unit-mass free body hung from the
    ground by a ball and socket, under gravity."""
    a = _Api(reference)
    w = a.core.World()
    if reference:
        from arboris.joints import FreeJoint
    else:
        from .joints import FreeJoint
    b0 = a.core.Body(mass=np.eye(6))
    w.add_link(w.ground, FreeJoint(), b0)
    w.register(a.controllers.WeightController())
    w.register(a.constraints.BallAndSocketConstraint(frames=(w.ground, b0)))
    w.init()
    return w


def simplearm_limits_world(reference=False, shoulder=3.14/2 - 0.1):
    """tests/test_constraints.py:11-32: JointLimits on the shoulder."""
    a = _Api(reference)
    w = a.core.World()
    a.simplearm.add_simplearm(w)
    w.register(a.controllers.WeightController())
    sh = w.getjoints()['Shoulder']
    w.register(a.constraints.JointLimits(sh, -3.14/2, 3.14/2))
    sh.gpos[0] = shoulder
    w.init()
    return w


BALLS = (("BallA", .10, 1.0, (0., .13, 0.)), ("BallB", .12, 1.5, (.05, .38, .02)),
         ("BallC", .08, 0.5, (.60, .31, .05)))          # name, radius, mass, start position
BALLS_POINT = ("Needle", 0.2, (.61, .50, .04))                     # name, mass, start position


def balls_world(reference=False):
    """the collision pairs of collisions.py beyond plane/point.  Three
    free balls and a free body carrying a Point fall on a ground plane and on a box fixed to
    the ground: plane/sphere, plane/point, box/sphere, sphere/sphere, sphere/point contacts,
    registered explicitly (``get_all_contacts`` would trip over the reference's undefined
    ``box_point_collision``)."""
    a = _Api(reference)
    w = a.core.World()
    a.simpleshapes.add_groundplane(w)
    plane = list(w.itershapes())[0]
    box = a.shapes.Box(a.core.SubFrame(w.ground, Hg.transl(.6, .1, 0.), 'BoxFrame'), (.2, .1, .2), 'Box')
    w.register(box)
    balls = []
    for name, radius, mass, pos in BALLS:
        body = a.core.Body(name=name, mass=a.massmatrix.sphere(radius, mass))
        j = a.joints.FreeJoint(name=name + 'Joint')
        w.add_link(w.ground, j, body)
        j.gpos[:] = Hg.transl(*pos)
        sh = a.shapes.Sphere(body, radius, name + 'Shape')
        w.register(sh)
        balls.append(sh)
    name, mass, pos = BALLS_POINT
    body = a.core.Body(name=name, mass=a.massmatrix.sphere(.05, mass))
    j = a.joints.FreeJoint(name=name + 'Joint')
    w.add_link(w.ground, j, body)
    j.gpos[:] = Hg.transl(*pos)
    tip = a.shapes.Point(a.core.SubFrame(body, Hg.transl(0., -.05, 0.), 'Tip'), 'TipShape')
    w.register(tip)
    w.register(a.controllers.WeightController())
    A, B, C = balls
    for s0, s1 in ((plane, A), (plane, B), (plane, C), (plane, tip), (box, C), (box, A),
                   (A, B), (B, C), (C, tip), (tip, B)):
        w.register(a.constraints.SoftFingerContact((s0, s1), .5))
    w.init()
    return w


def zoo_world(reference=False):
    """what the three benchmark robots do not exercise.  Two
    simplearms side by side (name prefixes), and a chain that goes through every other stock
    joint type -- TxTyTz, RzRx, RyRx, RzRy, RzRyRx, Ry, Rx -- with joint frames that are
    SubFrames on both sides (H_pr, H_cn != I), bodies with viscosity, one massless body, a
    two ProportionalDerivativeControllers with diagonal gains on five of the dofs and JointLimits
    on a hinge."""
    a = _Api(reference)
    w = a.core.World()
    a.simplearm.add_simplearm(w, name='Left')
    a.simplearm.add_simplearm(w, name='Right', lengths=(.4, .3, .1), masses=(.7, .5, .1))
    J, mm = a.joints, a.massmatrix
    specs = (('Slider', J.TxTyTzJoint, mm.box((.1, .2, .15), 2.)),
             ('RzRx', J.RzRxJoint, mm.ellipsoid((.1, .05, .2), 1.2)),
             ('RyRx', J.RyRxJoint, mm.cylinder(.3, .05, .8)),
             ('RzRy', J.RzRyJoint, None),
             ('Ball', J.RzRyRxJoint, mm.sphere(.1, 1.1)),
             ('Ry', J.RyJoint, mm.box((.05, .1, .05), .4)),
             ('Rx', J.RxJoint, mm.sphere(.05, .3)))
    parent = w.ground
    joints = {}
    for i, (name, cls, mass) in enumerate(specs):
        visc = np.diag([.02, .03, .01, .2, .1, .3])*(i % 3 == 0)
        body = a.core.Body(name=name + 'Body', mass=mass, viscosity=visc if visc.any() else None)
        H0 = np.dot(Hg.transl(.1 + .05*i, .2, -.03*i), Hg.rotzyx(.2*i, -.1, .3))
        H1 = np.dot(Hg.transl(0., -.05*i, .02), Hg.rotzyx(.1, .15*i, -.2))
        f0 = a.core.SubFrame(parent, H0, name + 'Base')
        f1 = a.core.SubFrame(body, H1, name + 'Tip') if i % 2 == 0 else body
        j = cls(name=name)
        w.add_link(f0, j, f1)
        joints[name] = j
        parent = body
    w.register(a.controllers.WeightController())
    w.init()      # the reference's PD controller reads joint.dof in its constructor
    # (the reference's PD controller stacks joint.gpos with array(): joints of equal ndof only)
    w.register(a.controllers.ProportionalDerivativeController(
        [joints['Slider']], kp=np.diag([30., 20., 25.]), kd=np.diag([3., 2., 2.5]),
        gpos_des=[.1, -.05, .02], gvel_des=[0., .1, 0.]))
    w.register(a.controllers.ProportionalDerivativeController(
        [joints['Ry'], joints['Rx']], kp=np.diag([4., 20.]), kd=np.diag([.4, .5]),
        gpos_des=[.3, .8], gvel_des=[-.2, 0.]))     # drives the Rx hinge into its limit
    w.register(a.constraints.JointLimits(joints['Rx'], -.4, .4))
    w.init()
    return w


BUILDERS = {
    "simplearm": simplearm_world,
    "human36_free": human36_free_world,
    "human36_contact": human36_contact_world,
    "snake_loop": snake_loop_world,
    "ball_socket": ball_socket_world,
    "simplearm_limits": simplearm_limits_world,
    "balls": balls_world,
    "zoo": zoo_world,
}


# ---------------------------------------------------------------------------
# seeded initial states, per world index (numpy, host)
# ---------------------------------------------------------------------------
def _free_and_linear(model):
    free = [j for j in range(len(model.joint_type)) if int(model.joint_type[j]) == 0]
    lin = np.ones(model.ngpos, bool)
    for j in free:
        g = int(model.joint_gpos[j])
        lin[g:g + 16] = False
    return free, lin


def _rotzyx_stack(r):
    """(W, 4, 4) stack of ``Hg.rotzyx(*r[i])`` (same closed form, evaluated on arrays)."""
    sz, cz = np.sin(r[:, 0]), np.cos(r[:, 0])
    sy, cy = np.sin(r[:, 1]), np.cos(r[:, 1])
    sx, cx = np.sin(r[:, 2]), np.cos(r[:, 2])
    H = np.zeros((r.shape[0], 4, 4))
    H[:, 0, 0], H[:, 0, 1], H[:, 0, 2] = cz*cy, cz*sy*sx - sz*cx, cz*sy*cx + sz*sx
    H[:, 1, 0], H[:, 1, 1], H[:, 1, 2] = sz*cy, sz*sy*sx + cz*cx, sz*sy*cx - cz*sx
    H[:, 2, 0], H[:, 2, 1], H[:, 2, 2] = -sy, cy*sx, cy*cx
    H[:, 3, 3] = 1.
    return H


def _matmul44(A, B):
    """Stacked 4x4 products with a fixed summation order (independent of the stack size, so a
    world's state does not depend on how many worlds are generated with it)."""
    out = np.zeros(np.broadcast_shapes(A.shape, B.shape))
    for k in range(4):
        out += A[..., :, k:k + 1]*B[..., k:k + 1, :]
    return out


# human36_contact: SURVEY.md section 8(d) config 3 -- root lifted by U(0, 0.05) m (left-multiplied
# as in tests/test_human36_falling.py:16-18), tilted by U(-.05, .05)^3 rad, joint angles U(-.1, .1),
# zero velocity.  (Round 1 used lift U(0.02, 0.07), tilt +-0.02.)
CONTACT_LIFT = (0., 0.05)
CONTACT_TILT = 0.05
if os.environ.get("ARB_B200_CONTACT_DIST") == "r1":      # A/B runs against round 1's distribution
    CONTACT_LIFT, CONTACT_TILT = (0.02, 0.07), 0.02


def _human36_states(model, scenario, w0, w1):
    """Vectorised seeded states of the two humanoid scenarios: the per-world random draws
    (``default_rng(SEED0 + w)``) stay per world, the matrix algebra is batched."""
    W = w1 - w0
    free, lin = _free_and_linear(model)
    nlin = int(lin.sum())
    gpos = np.repeat(np.array(model.gpos0, dtype=float)[:, None], W, 1)
    gvel = np.repeat(np.array(model.gvel0, dtype=float)[:, None], W, 1)
    ang = np.empty((W, nlin))
    t = np.zeros((W, 3))
    r = np.empty((W, 3))
    vel = np.empty((W, model.ndof)) if scenario == "human36_free" else None
    for i in range(W):
        rng = np.random.default_rng(SEED0 + w0 + i)
        if scenario == "human36_free":
            ang[i] = rng.uniform(-0.5, 0.5, nlin)
            t[i] = rng.uniform([-.1, 0., -.1], [.1, .2, .1])
            r[i] = rng.uniform(-.3, .3, 3)
            vel[i] = rng.uniform(-1., 1., model.ndof)
        else:
            ang[i] = rng.uniform(-0.1, 0.1, nlin)
            t[i, 1] = rng.uniform(*CONTACT_LIFT)
            r[i] = rng.uniform(-CONTACT_TILT, CONTACT_TILT, 3)
    gpos[lin] = ang.T
    g = int(model.joint_gpos[free[0]])
    H0 = np.array(model.gpos0, dtype=float)[g:g + 16].reshape(4, 4)
    T = np.zeros((W, 4, 4))
    T[:, range(4), range(4)] = 1.
    T[:, 0:3, 3] = t
    R = _rotzyx_stack(r)
    if scenario == "human36_free":
        H = _matmul44(_matmul44(T, R), H0[None])
        gvel[:] = vel.T
    else:
        H = _matmul44(T, _matmul44(H0[None], R))
    gpos[g:g + 16] = H.reshape(W, 16).T
    for c in range(len(model.cons_type)):       # limited joints start strictly inside their limits
        if int(model.cons_type[c]) == 0:
            k = int(model.cons_int[c][2])
            mn, mx, prox = model.cons_dbl[c][0:3]
            gpos[k] = np.minimum(np.maximum(gpos[k], mn + 2*prox), mx - 2*prox)
    return gpos, gvel


def initial_state(model, scenario, w):
    """(gpos, gvel) of world ``w`` for ``scenario``; deterministic in ``w``."""
    if scenario in ("human36_free", "human36_contact"):
        gp, gv = _human36_states(model, scenario, int(w), int(w) + 1)
        return gp[:, 0], gv[:, 0]
    rng = np.random.default_rng(SEED0 + int(w))
    gpos = np.array(model.gpos0, dtype=float)
    gvel = np.array(model.gvel0, dtype=float)
    free, lin = _free_and_linear(model)
    if scenario == "snake_loop":
        # start (almost) on the closed loop: a large loop error would be
        # corrected within one dt by BallAndSocketConstraint.solve and diverge
        gpos[lin] = SNAKE_QREF + rng.uniform(-1e-3, 1e-3, int(lin.sum()))
        gvel[:] = rng.uniform(-.2, .2, model.ndof)
    elif scenario == "balls":
        for j in free:
            g = int(model.joint_gpos[j])
            H0 = gpos[g:g + 16].reshape(4, 4)
            t = rng.uniform(-.01, .01, 3)
            r = rng.uniform(-.3, .3, 3)
            gpos[g:g + 16] = np.dot(Hg.transl(*t), np.dot(H0, Hg.rotzyx(*r))).reshape(-1)
        gvel[:] = rng.uniform(-.05, .05, model.ndof)
    elif scenario == "zoo":
        gpos[lin] = rng.uniform(-.3, .3, int(lin.sum()))
        gvel[:] = rng.uniform(-.5, .5, model.ndof)
    elif scenario in ("simplearm", "ball_socket", "simplearm_limits"):
        if w > 0:
            gpos[lin] += rng.uniform(-0.2, 0.2, int(lin.sum()))
            gvel[:] = rng.uniform(-.5, .5, model.ndof)
    else:
        raise KeyError(scenario)
    if scenario != "simplearm_limits":
        # keep limited joints strictly inside their limits (outside the
        # activation band) so that no world starts in violation
        for c in range(len(model.cons_type)):
            if int(model.cons_type[c]) == 0:
                g = int(model.cons_int[c][2])
                mn, mx, prox = model.cons_dbl[c][0:3]
                gpos[g] = min(max(gpos[g], mn + 2*prox), mx - 2*prox)
    return gpos, gvel


def initial_states(model, scenario, w0, w1):
    """Stacked states of worlds ``w0..w1-1``: gpos (ngpos, W), gvel (ndof, W)."""
    if scenario in ("human36_free", "human36_contact"):
        return _human36_states(model, scenario, int(w0), int(w1))
    W = w1 - w0
    gpos = np.empty((model.ngpos, W))
    gvel = np.empty((model.ndof, W))
    for i in range(W):
        gpos[:, i], gvel[:, i] = initial_state(model, scenario, w0 + i)
    return gpos, gvel
