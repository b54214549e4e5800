"""Host-side mirror of the reference's ``arboris/core.py`` object model.

Same names and argument meanings as the reference (``World`` core.py:342,
``Body`` :1055, ``SubFrame``/``MovingSubFrame`` :1039/:1045, the ``Joint`` :158,
``Constraint`` :269, ``Controller`` :327, ``Observer`` :1318 base classes,
``NamedObjectsList`` :56, ``JointsList`` :243, ``simulate`` :1334) so robot
factories and scripts written against the reference build the same tree here.

What differs is *where the arithmetic happens*: these classes only describe the
model.  ``World.update_dynamic / update_controllers / update_constraints /
integrate`` flatten the tree once (``arboris_b200.flatten``) and run the CUDA
kernels through the C ABI (``arboris_b200.batch.BatchedWorld``) -- for one
world when called on a ``World``, for thousands in lockstep when called on a
``BatchedWorld``.  There is no numpy implementation of the step in this package
and no CPU fallback: without the CUDA library the step methods raise.
"""
import numpy as np
from numpy import array, zeros, eye

from . import homogeneousmatrix as Hg


class NamedObject(object):
    def __init__(self, name=None):
        self.name = name

    def __repr__(self):
        if self.name is None:
            return object.__repr__(self)
        return '<{0}.{1} object named "{2}" at "{3}")>'.format(
            type(self).__module__, type(self).__name__, self.name, hex(id(self)))


class DuplicateNameError(Exception):
    pass


class NamedObjectsList(list):
    """A list whose items can also be fetched by ``name`` (core.py:56-123)."""

    def __init__(self, iterable=None):
        list.__init__(self, () if iterable is None else iterable)

    def find(self, name):
        return [o for o in self
                if isinstance(o, NamedObject) and o.name == name]

    def __getitem__(self, index):
        if isinstance(index, str):
            for o in self:
                if isinstance(o, NamedObject) and o.name == index:
                    return o
            raise KeyError('No object named "{0}".'.format(index))
        return list.__getitem__(self, index)

    def as_dict(self):
        out = {}
        for o in self:
            if isinstance(o, NamedObject) and o.name is not None:
                if o.name in out:
                    raise DuplicateNameError()
                out[o.name] = o
        return out


class Frame(object):
    """Anything with a ``pose``, ``jacobian``, ``djacobian``, ``twist``, ``body``
    and ``bpose`` (core.py:129-156)."""


class Joint(NamedObject):
    """Base class of the ideal joints (core.py:158-220).

    A joint here is a *description*: its type (the subclass), its generalized
    position ``gpos`` and velocity ``gvel``.  The closed-form ``pose`` /
    ``jacobian`` / ``djacobian`` of the 9 stock joint types live on the device
    (``csrc/arb_joints.cuh``); user subclasses are refused by the flattener.
    """
    ndof = None

    def __init__(self, name=None):
        NamedObject.__init__(self, name)
        self._frame0 = None
        self._frame1 = None
        self._dof = None  # set by World.init()

    @property
    def dof(self):
        if self._dof is None:
            raise ValueError
        return self._dof

    @property
    def frames(self):
        return (self._frame0, self._frame1)


class LinearConfigurationSpaceJoint(Joint):
    """Joints whose configuration space is R^ndof (core.py:223-240)."""

    def __init__(self, gpos=None, gvel=None, name=None):
        n = self.ndof
        # always float64: integer gpos is a latent bug of the reference
        # (energy_drift.h5)
        self.gpos = zeros(n) if gpos is None else array(gpos, dtype=float).reshape(n)
        self.gvel = zeros(n) if gvel is None else array(gvel, dtype=float).reshape(n)
        Joint.__init__(self, name)


class JointsList(NamedObjectsList):
    def __init__(self, iterable):
        NamedObjectsList.__init__(self, iterable)
        dofs = []
        for o in self:
            if isinstance(o, Joint):
                dofs.extend(range(o.dof.start, o.dof.stop))
        if dofs == list(range(dofs[0], dofs[0] + len(dofs))) if dofs else True:
            self._dof = slice(dofs[0], dofs[-1] + 1) if dofs else slice(0, 0)
        else:
            self._dof = dofs

    @property
    def dof(self):
        return self._dof


class Constraint(NamedObject):
    """Base class of the Gauss-Seidel constraints (core.py:269-315)."""

    def __init__(self, name=None):
        NamedObject.__init__(self, name)
        self._is_enabled = True
        self._is_active = None  # refreshed from the device after update_constraints
        self._jacobian = None
        self._dol = None

    def is_enabled(self):
        return self._is_enabled

    def enable(self):
        self._is_enabled = True

    def disable(self):
        self._is_enabled = False

    def init(self, world):
        pass

    def is_active(self):
        return self._is_active

    @property
    def jacobian(self):
        return self._jacobian

    @property
    def gforce(self):
        return np.dot(self.jacobian.T, self._force)


class Shape(NamedObject):
    def __init__(self, frame, name=None):
        assert isinstance(frame, Frame)
        self.frame = frame
        NamedObject.__init__(self, name)


class Controller(NamedObject):
    def __init__(self, name=None):
        NamedObject.__init__(self, name)

    def init(self, world):
        pass


class Observer(object):
    """``init(world, timeline)``, ``update(dt)``, ``finish()`` (core.py:1318-1331)."""

    def init(self, world, timeline):
        pass

    def update(self, dt):
        pass

    def finish(self):
        pass


class _SubFrame(NamedObject, Frame):
    """A frame rigidly fixed to a body (core.py:983-1036)."""

    def __init__(self, body, bpose=None, name=None):
        if bpose is None:
            bpose = eye(4)
        NamedObject.__init__(self, name)
        assert Hg.ishomogeneousmatrix(bpose)
        self._bpose = array(bpose, dtype=float)
        if not isinstance(body, Body):
            raise ValueError("The ``body`` argument must be an instance of the ``Boby`` class")
        self._body = body

    @property
    def pose(self):
        return np.dot(self._body.pose, self._bpose)

    @property
    def twist(self):
        return np.dot(Hg.iadjoint(self._bpose), self._body.twist)

    @property
    def jacobian(self):
        return np.dot(Hg.iadjoint(self._bpose), self._body.jacobian)

    @property
    def djacobian(self):
        return np.dot(Hg.iadjoint(self._bpose), self._body.djacobian)

    @property
    def body(self):
        return self._body

    @property
    def bpose(self):
        return self._bpose.copy()


class SubFrame(_SubFrame):
    pass


class MovingSubFrame(_SubFrame):
    @_SubFrame.bpose.setter
    def bpose(self, bpose):
        assert Hg.ishomogeneousmatrix(bpose)
        self._bpose[:] = bpose


class Body(NamedObject, Frame):
    """A rigid body: a 6x6 mass and viscosity matrix plus its place in the tree
    (core.py:1055-1133).  ``pose/jacobian/djacobian/twist/nleffects`` hold what
    the last ``World.update_dynamic`` read back from the device."""

    def __init__(self, name=None, mass=None, viscosity=None):
        NamedObject.__init__(self, name)
        self.parentjoint = None
        self.childrenjoints = []
        self.mass = zeros((6, 6)) if mass is None else array(mass, dtype=float)
        self.viscosity = zeros((6, 6)) if viscosity is None else array(viscosity, dtype=float)
        self._pose = None
        self._jacobian = None
        self._djacobian = None
        self._twist = None
        self._nleffects = None

    def iter_descendant_bodies(self):
        for j in self.childrenjoints:
            b = j._frame1.body
            yield b
            for bb in b.iter_descendant_bodies():
                yield bb

    def iter_ancestor_bodies(self):
        if self.parentjoint is not None:
            parent = self.parentjoint._frame0.body
            yield parent
            for a in parent.iter_ancestor_bodies():
                yield a

    def iter_descendant_joints(self):
        for j in self.childrenjoints:
            yield j
            for jj in j._frame1.body.iter_descendant_joints():
                yield jj

    def iter_ancestor_joints(self):
        if self.parentjoint is not None:
            yield self.parentjoint
            for a in self.parentjoint._frame0.body.iter_ancestor_joints():
                yield a

    pose = property(lambda self: self._pose)
    jacobian = property(lambda self: self._jacobian)
    djacobian = property(lambda self: self._djacobian)
    twist = property(lambda self: self._twist)
    nleffects = property(lambda self: self._nleffects)

    @property
    def bpose(self):
        return eye(4)

    @property
    def body(self):
        return self


class World(NamedObject):
    """The model tree plus the four step methods (core.py:342-980).

    Building (``add_link``, ``replace_joint``, ``register``, ``init``) is host
    Python.  Stepping runs on the GPU: the first step call flattens the tree
    and creates a one-world ``BatchedWorld``; host-visible state (``joint.gpos``,
    ``joint.gvel``, ``body.pose`` ...) is synchronised around each call so code
    written for the reference (``tests/test_update_dynamic.py``,
    ``tests/test_constraints.py``) reads the same.  For many worlds use
    ``arboris_b200.BatchedWorld(world, nworlds)`` directly.
    """

    def __init__(self, name=None):
        NamedObject.__init__(self, name)
        self.ground = Body('ground')
        self._current_time = 0.
        self._up = array((0., 1., 0.))
        self._controllers = []
        self._constraints = []
        self._subframes = []
        self._shapes = []
        self._ndof = 0
        self._gvel = array([])
        self._mass = array([])
        self._gforce = array([])
        self._viscosity = array([])
        self._nleffects = array([])
        self._impedance = array([])
        self._admittance = array([])
        self._batch = None  # one-world device batch, built lazily

    # ---- iteration helpers (core.py:365-436) -------------------------------
    def iterbodies(self):
        yield self.ground
        for b in self.ground.iter_descendant_bodies():
            yield b

    def getbodies(self):
        return NamedObjectsList(self.iterbodies())

    def iterconstraints(self):
        return iter(self._constraints)

    def itersubframes(self):
        return iter(self._subframes)

    def itermovingsubframes(self):
        return (f for f in self._subframes if isinstance(f, MovingSubFrame))

    def iterframes(self):
        for b in self.iterbodies():
            yield b
        for f in self._subframes:
            yield f

    def getframes(self):
        frames = self.getbodies()
        frames.extend(self._subframes)
        return frames

    def itershapes(self):
        return iter(self._shapes)

    def getshapes(self):
        return NamedObjectsList(self._shapes)

    def iterjoints(self):
        return self.ground.iter_descendant_joints()

    def getjoints(self):
        return JointsList(self.iterjoints())

    # ---- building (core.py:438-560) -----------------------------------------
    def add_link(self, frame0, joint, frame1, *args):
        assert isinstance(frame0, Frame)
        assert isinstance(frame1, Frame)
        assert isinstance(joint, Joint)
        assert joint._frame0 is None
        assert joint._frame1 is None
        assert len(args) % 3 == 0
        joint._frame0 = frame0
        joint._frame1 = frame1
        if frame1.body.parentjoint is not None:
            raise ValueError(
                'frame1\'s body already has a parent joint, which means you\'re '
                'probably trying to create a kinematic loop. Try using a '
                'constraint instead.')
        frame1.body.parentjoint = joint
        frame0.body.childrenjoints.append(joint)
        self.register(frame0)
        self.register(frame1)
        self._batch = None
        if args:
            self.add_link(*args)

    def replace_joint(self, old_joint, *args):
        assert isinstance(old_joint, Joint)
        assert old_joint in old_joint._frame0.body.childrenjoints
        assert old_joint is old_joint._frame1.body.parentjoint
        if len(args) == 1:
            return self.replace_joint(old_joint, old_joint._frame0, args[0],
                                      old_joint._frame1)
        if len(args) == 0 or len(args) % 3 != 0:
            raise RuntimeError()
        body0 = args[0].body
        body1 = args[-1].body
        assert old_joint._frame0.body is body0
        assert old_joint._frame1.body is body1
        body1.parentjoint = None
        old_joint._frame0 = None
        old_joint._frame1 = None
        self.add_link(*args)
        # the new chain was appended; put it where the old joint was
        i = body0.childrenjoints.index(old_joint)
        body0.childrenjoints[i] = body0.childrenjoints.pop()
        self.init()

    def register(self, obj):
        if isinstance(obj, Body):
            pass
        elif isinstance(obj, Joint):
            raise ValueError('Joints should not be registered. Use add_link() instead.')
        elif isinstance(obj, _SubFrame):
            if obj not in self._subframes:
                self._subframes.append(obj)
        elif isinstance(obj, Shape):
            if obj not in self._shapes:
                self._shapes.append(obj)
            self.register(obj.frame)
        elif isinstance(obj, Constraint):
            if obj not in self._constraints:
                self._constraints.append(obj)
                for f in getattr(obj, '_contact_frames', ()):
                    self.register(f)
        elif isinstance(obj, Controller):
            if obj not in self._controllers:
                self._controllers.append(obj)
        else:
            raise ValueError(
                'I do not know how to register objects of type {0}'.format(type(obj)))
        self._batch = None

    def parse(self, target):
        """Depth-first walk calling ``target.register/add_link`` (core.py:562-606)."""
        registered = set()

        def reg_frame(frame):
            if id(frame) in registered:
                return
            registered.add(id(frame))
            target.register(frame)
            if isinstance(frame, Body):
                for f in self._subframes:
                    if id(f) not in registered and frame is f.body:
                        reg_frame(f)
            for s in self._shapes:
                if frame is s.frame:
                    target.register(s)

        def walk(children):
            for j in children:
                f0, f1 = j.frames
                target.add_link(f0, j, f1)
                reg_frame(f1.body)
                walk(f1.body.childrenjoints)

        target.init_parse(self.ground, self.up, self.current_time)
        reg_frame(self.ground)
        walk(self.ground.childrenjoints)
        for c in self._constraints:
            target.register(c)
        for c in self._controllers:
            target.register(c)

    def init(self):
        """Number the dofs in depth-first joint order and size the model
        matrices (core.py:608-635)."""
        n = 0
        for j in self.iterjoints():
            j._dof = slice(n, n + j.ndof)
            n += j.ndof
        self._ndof = n
        self._mass = zeros((n, n))
        self._nleffects = zeros((n, n))
        self._viscosity = zeros((n, n))
        self._gforce = zeros(n)
        self._gvel = zeros(n)
        for j in self.iterjoints():
            self._gvel[j.dof] = j.gvel[:]
            j.gvel = self._gvel[j.dof]  # a view, as in the reference
        for c in self._constraints:
            c.init(self)
        for a in self._controllers:
            a.init(self)
        self._batch = None

    current_time = property(lambda self: self._current_time)
    up = property(lambda self: self._up)
    mass = property(lambda self: self._mass)
    viscosity = property(lambda self: self._viscosity)
    nleffects = property(lambda self: self._nleffects)
    ndof = property(lambda self: self._ndof)
    gvel = property(lambda self: self._gvel.copy())
    gforce = property(lambda self: self._gforce.copy())

    # ---- the step, on the GPU -----------------------------------------------
    def _device(self, refresh=False):
        """The single-world device batch.  The reference reads constraint and controller
        parameters live at every step (``is_enabled()`` core.py:913, ``gpos_des`` / ``kp``
        controllers.py:141-159, ``_min`` / ``_max`` constraints.py:73-90, ``_mu``, body masses):
        with ``refresh`` (the start of a step) the world is flattened again and the device model is
        rebuilt when any parameter differs from the one the batch was made from, so scripts that
        retarget a PD set-point or toggle a contact between steps behave as they do there."""
        if self._batch is not None and refresh:
            from .flatten import flatten
            if not self._batch.model.same_parameters(flatten(self)):
                self._batch.close()
                self._batch = None
        if self._batch is None:
            from .batch import BatchedWorld
            self._batch = BatchedWorld(self, nworlds=1)
        return self._batch

    def update_geometric(self):
        self.update_dynamic()

    def update_dynamic(self):
        """core.py:682-734, executed by ``arb_update_dynamic``."""
        b = self._device(refresh=True)
        b.push_host_state(self)
        b.update_dynamic()
        b.pull_dynamic(self)

    def update_controllers(self, dt):
        """core.py:736-818, executed by ``arb_update_controllers``."""
        assert dt > 0
        b = self._device()
        b.update_controllers(dt)
        b.pull_controllers(self)

    def update_constraints(self, dt):
        """core.py:820-937, executed by ``arb_update_constraints``."""
        assert dt > 0
        b = self._device()
        b.update_constraints(dt)
        b.pull_constraints(self)

    def integrate(self, dt):
        """core.py:939-980, executed by ``arb_integrate``."""
        assert dt > 0
        b = self._device()
        b.integrate(dt)
        b.pull_state(self)
        self._current_time += dt


def simulate(world, timeline, observers=()):
    """Run a full simulation (core.py:1334-1365): for each interval of
    ``timeline``: update_dynamic, update_controllers, update_constraints,
    observers, integrate.  ``world`` may be a ``World`` (one world) or a
    ``BatchedWorld`` (many worlds in lockstep)."""
    world._current_time = timeline[0]
    world.init()
    for obs in observers:
        obs.init(world, timeline)
    # A batched world keeps the fused CUDA step when every observer only reads what the fused
    # step leaves behind (state, body poses / twists, constraint forces and active sets):
    # begin_step = the three update_* phases, end_step = integrate.  Without observers whole
    # runs of steps go to the device in one call.
    fused = hasattr(world, "begin_step") and all(getattr(o, "fused_ok", False) for o in observers)
    if fused and not observers:
        import numpy as _np
        dts = _np.diff(_np.asarray(timeline, dtype=float))
        if len(dts):
            world.step(dts)
        return
    for next_time in timeline[1:]:
        dt = next_time - world._current_time
        if fused:
            world.begin_step(dt)
        else:
            world.update_dynamic()
            world.update_controllers(dt)
            world.update_constraints(dt)
        for obs in observers:
            obs.update(dt)
        if fused:
            world.end_step(dt)
        else:
            world.integrate(dt)
    for obs in observers:
        obs.finish()
