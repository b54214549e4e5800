"""Constraints solved by the per-world Gauss-Seidel projection
(reference ``arboris/constraints.py``).

The classes describe the constraint (frames, limits, friction) and expose what
the device computed for the last step (``_force``, ``is_active()``).  The
arithmetic -- collision, activation, Jacobians, the three-way soft-finger
``solve`` -- is in ``csrc/arb_constraints.cuh``.
"""
from numpy import array, zeros

from .core import Constraint, MovingSubFrame, Shape, World, LinearConfigurationSpaceJoint
from .shapes import Plane, Point, Sphere, Box

point_contact_proximity = 0.02   # constraints.py:12
joint_limits_proximity = 0.01    # constraints.py:13


class JointLimits(Constraint):
    """``min <= q <= max`` on a joint (constraints.py:15-90).  As in the
    reference only 1-dof joints are meaningful (``ndol`` is 1)."""
    ndol = 1

    def __init__(self, joint, min, max, proximity=None, name=None):
        if not isinstance(joint, LinearConfigurationSpaceJoint):
            raise ValueError()
        Constraint.__init__(self, name)
        n = joint.ndof
        self._joint = joint
        self._min = array(min, dtype=float).reshape((n,))
        self._max = array(max, dtype=float).reshape((n,))
        self._proximity = zeros(n)
        self._proximity[:] = joint_limits_proximity if proximity is None else proximity
        self._force = zeros(n)


class BallAndSocketConstraint(Constraint):
    """Keeps the origins of two frames together (constraints.py:92-237).
    ``_force`` persists across steps (warm start), as in the reference."""
    ndol = 3

    def __init__(self, frames, name=None):
        Constraint.__init__(self, name)
        self._force = zeros(3)
        self._pos0 = None
        self._frames = frames

    def is_active(self):
        return True


def choose_solver(shape0, shape1):
    """Order the pair and name the solver as the reference's
    ``collisions.choose_solver`` (collisions.py:14-65) does.  The solvers themselves
    are device code (``contact_collide`` in csrc/arb_constraints.cuh).  Box/Point is
    refused: the reference names a ``box_point_collision`` that it never defines."""
    assert isinstance(shape0, Shape)
    assert isinstance(shape1, Shape)
    table = {(Sphere, Sphere): (False, 'sphere_sphere_collision'),
             (Sphere, Point): (False, 'sphere_point_collision'),
             (Sphere, Plane): (True, 'plane_sphere_collision'),
             (Sphere, Box): (True, 'box_sphere_collision'),
             (Point, Sphere): (True, 'sphere_point_collision'),
             (Point, Plane): (True, 'plane_point_collision'),
             (Plane, Sphere): (False, 'plane_sphere_collision'),
             (Plane, Point): (False, 'plane_point_collision'),
             (Box, Sphere): (False, 'box_sphere_collision')}
    for (c0, c1), (swap, solver) in table.items():
        if type(shape0) is c0 and type(shape1) is c1:
            return ((shape1, shape0) if swap else (shape0, shape1)), solver
    raise NotImplementedError()


class PointContact(Constraint):
    """Parent of the point contacts (constraints.py:240-297)."""

    def __init__(self, shapes, collision_solver, proximity, name):
        assert isinstance(shapes[0], Shape)
        assert isinstance(shapes[1], Shape)
        Constraint.__init__(self, name)
        if collision_solver is None:
            shapes, collision_solver = choose_solver(shapes[0], shapes[1])
        self._shapes = shapes
        self._sdist = None
        self._frames = (MovingSubFrame(shapes[0].frame.body),
                        MovingSubFrame(shapes[1].frame.body))
        self._contact_frames = self._frames  # World.register adds them (core.py:550-554)
        self._collision_solver = collision_solver
        self._proximity = proximity


class SoftFingerContact(PointContact):
    """Point contact with elliptic Coulomb friction incl. torsion
    (constraints.py:300-836); rows are ``[w_z, v_x, v_y, v_z]``."""
    ndol = 4

    def __init__(self, shapes, friction_coeff, collision_solver=None,
                 proximity=0.02, name=None):
        self._mu = friction_coeff
        PointContact.__init__(self, shapes, collision_solver, proximity, name)
        self._force = zeros(4)
        self._eps = array((1., 1., 1.))


def get_all_contacts(world, contact_class=None, **args):
    """All shape pairs i<j on different bodies that have a collision solver,
    in shape registration order (constraints.py:839-878)."""
    assert isinstance(world, World)
    if contact_class is None:
        contact_class = SoftFingerContact
    else:
        assert issubclass(contact_class, PointContact)
    contacts = []
    shapes = tuple(world.itershapes())
    for i, s0 in enumerate(shapes):
        for s1 in shapes[i + 1:]:
            if s0.frame.body is s1.frame.body:
                continue
            try:
                contacts.append(contact_class((s0, s1), **args))
            except NotImplementedError:
                pass
    return contacts
