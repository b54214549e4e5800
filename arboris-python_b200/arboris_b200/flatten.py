"""Flatten a ``World`` object tree into the immutable, array-only model the
device consumes (``arb_model_desc`` in ``include/arboris_b200.h``).

The walk is duck-typed on class *names* so that it accepts both this package's
``World`` and an unmodified reference ``arboris.core.World`` (that is the
"existing robots drop in" contract: ``tests/test_flatten.py`` flattens the real
reference's ``human36``/``simplearm``/``snake`` and this package's and requires
identical arrays).  Ordering follows the reference exactly:

* joints, hence dofs and moving bodies, in depth-first order of
  ``Body.childrenjoints`` (core.py:416-419, 608-615, 1094-1099); body ``j+1`` is
  the child of joint ``j``, body 0 is the ground;
* constraints in registration order of ``World._constraints`` (core.py:913);
* controllers in registration order (core.py:814).

Unknown ``Joint`` / ``Constraint`` / ``Controller`` / shape-pair classes raise:
there is no host fallback for arithmetic the device does not implement.
"""
import numpy as np

# enums shared with include/arboris_b200.h
JOINT_TYPES = ("FreeJoint", "RzRyRxJoint", "RzRyJoint", "RzRxJoint", "RyRxJoint",
               "RzJoint", "RyJoint", "RxJoint", "TxTyTzJoint")
JOINT_NDOF = (6, 3, 2, 2, 2, 1, 1, 1, 3)
JOINT_NGPOS = (16, 3, 2, 2, 2, 1, 1, 1, 3)
CONS_JOINT_LIMITS, CONS_BALL_SOCKET, CONS_SOFT_FINGER = 0, 1, 2
CONS_NDOL = (1, 3, 4)
CTRL_WEIGHT, CTRL_PD = 0, 1
PAIR_PLANE_SPHERE, PAIR_SPHERE_SPHERE, PAIR_BOX_SPHERE = 0, 1, 2   # arb_contact_pair
CONS_NINT = 4
CONS_NDBL = 48

_FIELDS = ("ndof", "ngpos", "joint_type", "joint_parent", "joint_dof",
           "joint_gpos", "joint_Hpr", "joint_Hcn", "body_mass", "body_visc",
           "cons_type", "cons_int", "cons_dbl", "cons_row", "nrows",
           "ctrl_type", "ctrl_int", "ctrl_dbl", "ctrl_blob", "up",
           "gpos0", "gvel0", "cforce0")


def _class_names(obj):
    return [c.__name__ for c in type(obj).__mro__]


def _kind(obj, candidates):
    names = _class_names(obj)
    for i, cand in enumerate(candidates):
        if cand in names:
            return i
    return -1


class FlatModel(object):
    """Array-only model description (see module docstring for conventions)."""

    def __init__(self, **kw):
        for k in _FIELDS:
            setattr(self, k, kw[k])
        self.body_names = list(kw.get("body_names", []))
        self.joint_names = list(kw.get("joint_names", []))

    nj = property(lambda self: len(self.joint_type))
    nc = property(lambda self: len(self.cons_type))

    def to_dict(self):
        d = {k: np.asarray(getattr(self, k)) for k in _FIELDS}
        d["body_names"] = np.array([n or "" for n in self.body_names])
        d["joint_names"] = np.array([n or "" for n in self.joint_names])
        return d

    @classmethod
    def from_dict(cls, d):
        kw = {}
        for k in _FIELDS:
            v = np.asarray(d[k])
            kw[k] = int(v) if v.ndim == 0 and k in ("ndof", "ngpos", "nrows") else v
        kw["body_names"] = [str(s) for s in d["body_names"]] if "body_names" in d else []
        kw["joint_names"] = [str(s) for s in d["joint_names"]] if "joint_names" in d else []
        return cls(**kw)

    def save(self, path):
        np.savez(path, **self.to_dict())

    # fields that hold state (flatten() copies the world's current gpos / gvel / forces there), not
    # parameters: a model does not change because the world moved
    _STATE_FIELDS = ("gpos0", "gvel0", "cforce0")

    def same_parameters(self, other):
        """True when ``other`` describes the same model: topology, constants, constraint and
        controller parameters, enable flags (everything but the initial state)."""
        for k in _FIELDS:
            if k in self._STATE_FIELDS:
                continue
            a, b = np.asarray(getattr(self, k)), np.asarray(getattr(other, k))
            if a.shape != b.shape or not np.array_equal(a, b):
                return False
        return True

    @classmethod
    def load(cls, path):
        with np.load(path, allow_pickle=False) as z:
            return cls.from_dict({k: z[k] for k in z.files})

    # ---- derived topology tables used by host code and the C side ----------
    def ancestors_dofs(self):
        """For each moving body (index 1..nj) the dofs of the joints on its
        root path, root first: the non-zero Jacobian columns."""
        out = [[]]
        for j in range(self.nj):
            p = int(self.joint_parent[j])
            own = list(range(int(self.joint_dof[j]),
                             int(self.joint_dof[j]) + JOINT_NDOF[int(self.joint_type[j])]))
            out.append(out[p] + own)
        return out


def _frame_of(frame, body_index):
    """(body index, 4x4 bpose) of a Body or SubFrame-like object."""
    body = frame.body
    return body_index[id(body)], np.array(frame.bpose, dtype=float).reshape(4, 4)


def flatten(world):
    """Return the ``FlatModel`` of ``world`` (this package's or the reference's)."""
    joints = list(world.ground.iter_descendant_joints())
    nj = len(joints)
    body_index = {id(world.ground): 0}
    for k, j in enumerate(joints):
        body_index[id(j._frame1.body)] = k + 1

    joint_type = np.zeros(nj, np.int32)
    joint_parent = np.zeros(nj, np.int32)
    joint_dof = np.zeros(nj, np.int32)
    joint_gpos = np.zeros(nj, np.int32)
    Hpr = np.zeros((nj, 4, 4))
    Hcn = np.zeros((nj, 4, 4))
    mass = np.zeros((nj, 6, 6))
    visc = np.zeros((nj, 6, 6))
    gpos0, gvel0 = [], []
    ndof = ngpos = 0
    joint_index = {}
    for k, j in enumerate(joints):
        t = _kind(j, JOINT_TYPES)
        if t < 0:
            raise NotImplementedError(
                "joint class %s has no device implementation" % type(j).__name__)
        joint_index[id(j)] = k
        joint_type[k] = t
        joint_parent[k], Hpr[k] = _frame_of(j._frame0, body_index)
        child, Hcn[k] = _frame_of(j._frame1, body_index)
        assert child == k + 1
        joint_dof[k] = ndof
        joint_gpos[k] = ngpos
        ndof += JOINT_NDOF[t]
        ngpos += JOINT_NGPOS[t]
        body = j._frame1.body
        mass[k] = np.asarray(body.mass, dtype=float)
        visc[k] = np.asarray(body.viscosity, dtype=float)
        gpos0.extend(np.asarray(j.gpos, dtype=float).reshape(-1))
        gvel0.extend(np.asarray(j.gvel, dtype=float).reshape(-1))

    # ---- constraints ---------------------------------------------------------
    cons = list(world._constraints)
    nc = len(cons)
    cons_type = np.zeros(nc, np.int32)
    cons_int = np.zeros((nc, CONS_NINT), np.int32)
    cons_dbl = np.zeros((nc, CONS_NDBL))
    cons_row = np.zeros(nc, np.int32)
    cforce0 = []
    nrows = 0
    for k, c in enumerate(cons):
        t = _kind(c, ("JointLimits", "BallAndSocketConstraint", "SoftFingerContact"))
        if t < 0:
            raise NotImplementedError(
                "constraint class %s has no device implementation" % type(c).__name__)
        cons_type[k] = t
        cons_row[k] = nrows
        nrows += CONS_NDOL[t]
        enabled = 1 if c.is_enabled() else 0
        if t == CONS_JOINT_LIMITS:
            jk = joint_index[id(c._joint)]
            if JOINT_NDOF[joint_type[jk]] != 1:
                # the reference's JointLimits (constraints.py:54-71) only works
                # for 1-dof joints: ndol is 1 and is_active() compares arrays
                raise NotImplementedError("JointLimits needs a 1-dof joint")
            cons_int[k] = (jk, joint_dof[jk], joint_gpos[jk], enabled)
            cons_dbl[k, 0:3] = (float(c._min[0]), float(c._max[0]), float(c._proximity[0]))
            cforce0.extend(np.asarray(c._force, dtype=float).reshape(-1)[:1])
        elif t == CONS_BALL_SOCKET:
            b0, H0 = _frame_of(c._frames[0], body_index)
            b1, H1 = _frame_of(c._frames[1], body_index)
            cons_int[k] = (b0, b1, 0, enabled)
            cons_dbl[k, 0:16] = H0.reshape(-1)
            cons_dbl[k, 16:32] = H1.reshape(-1)
            cforce0.extend(np.asarray(c._force, dtype=float).reshape(-1)[:3])
        else:
            s0, s1 = c._shapes      # already ordered by choose_solver (collisions.py:14-65)
            k0 = _kind(s0, ("Plane", "Sphere", "Box", "Point"))
            k1 = _kind(s1, ("Sphere", "Point"))
            pair = {0: PAIR_PLANE_SPHERE, 1: PAIR_SPHERE_SPHERE, 2: PAIR_BOX_SPHERE}.get(k0, -1)
            if pair < 0 or k1 < 0:
                raise NotImplementedError(
                    "contact pair %s/%s has no device collision solver"
                    % (type(s0).__name__, type(s1).__name__))
            b0, H0 = _frame_of(s0.frame, body_index)
            b1, H1 = _frame_of(s1.frame, body_index)
            cons_int[k] = (b0, b1, pair, enabled)
            cons_dbl[k, 0:16] = H0.reshape(-1)
            cons_dbl[k, 16:32] = H1.reshape(-1)
            if pair == PAIR_PLANE_SPHERE:
                cons_dbl[k, 32:36] = np.asarray(s0.coeffs, dtype=float)
            elif pair == PAIR_BOX_SPHERE:
                cons_dbl[k, 32:35] = np.asarray(s0.half_extents, dtype=float)
            else:
                cons_dbl[k, 41] = float(s0.radius)
            cons_dbl[k, 42] = float(getattr(s1, "radius", 0.))
            cons_dbl[k, 36] = float(c._mu)
            cons_dbl[k, 37:40] = np.asarray(c._eps, dtype=float)
            cons_dbl[k, 40] = float(c._proximity)
            cforce0.extend(np.asarray(c._force, dtype=float).reshape(-1)[:4])

    # ---- controllers -----------------------------------------------------------
    ctrls = list(world._controllers)
    na = len(ctrls)
    ctrl_type = np.zeros(na, np.int32)
    ctrl_int = np.zeros((na, 4), np.int32)
    ctrl_dbl = np.zeros((na, 4))
    blob = []
    for k, a in enumerate(ctrls):
        t = _kind(a, ("WeightController", "ProportionalDerivativeController"))
        if t < 0:
            raise NotImplementedError(
                "controller class %s has no device implementation" % type(a).__name__)
        ctrl_type[k] = t
        if t == CTRL_WEIGHT:
            ctrl_dbl[k, 0] = float(a.gravity)
        else:
            dofs, gposs = [], []
            for j in a.joints:
                jk = joint_index[id(j)]
                nd = JOINT_NDOF[joint_type[jk]]
                dofs.extend(range(joint_dof[jk], joint_dof[jk] + nd))
                gposs.extend(range(joint_gpos[jk], joint_gpos[jk] + nd))
            m = len(dofs)
            # blob layout: [dof map (m), gpos map (m), kp (m*m), kd (m*m), q_des (m), dq_des (m)]
            ctrl_int[k] = (m, len(blob), 0, 0)
            blob.extend(float(x) for x in dofs)
            blob.extend(float(x) for x in gposs)
            blob.extend(np.asarray(a.kp, dtype=float).reshape(-1))
            blob.extend(np.asarray(a.kd, dtype=float).reshape(-1))
            blob.extend(np.asarray(a.gpos_des, dtype=float).reshape(-1))
            blob.extend(np.asarray(a.gvel_des, dtype=float).reshape(-1))

    bodies = [world.ground] + [j._frame1.body for j in joints]
    return FlatModel(
        ndof=ndof, ngpos=ngpos, joint_type=joint_type, joint_parent=joint_parent,
        joint_dof=joint_dof, joint_gpos=joint_gpos, joint_Hpr=Hpr, joint_Hcn=Hcn,
        body_mass=mass, body_visc=visc, cons_type=cons_type, cons_int=cons_int,
        cons_dbl=cons_dbl, cons_row=cons_row, nrows=nrows, ctrl_type=ctrl_type,
        ctrl_int=ctrl_int, ctrl_dbl=ctrl_dbl, ctrl_blob=np.array(blob, dtype=float),
        up=np.asarray(world.up, dtype=float).copy(),
        gpos0=np.array(gpos0, dtype=float), gvel0=np.array(gvel0, dtype=float),
        cforce0=np.array(cforce0, dtype=float),
        body_names=[getattr(b, "name", None) for b in bodies],
        joint_names=[getattr(j, "name", None) for j in joints])
