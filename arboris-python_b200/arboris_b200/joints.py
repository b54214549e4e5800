"""The nine stock joint types (reference ``arboris/joints.py``).

Each class only carries its state (``gpos``, ``gvel``) and its type; the closed
forms for ``pose``, ``jacobian`` and ``djacobian`` (joints.py:10-384) are
evaluated on the device by ``csrc/arb_joints.cuh``.  ``TYPE_CODE`` is the enum
shared with ``include/arboris_b200.h`` (``arb_joint_type``).
"""
from numpy import array, eye, zeros

from .core import Joint, LinearConfigurationSpaceJoint


class FreeJoint(Joint):
    """6-dof joint; ``gpos`` is the 4x4 pose, ``gvel`` the body twist (joints.py:10-57)."""
    ndof = 6
    TYPE_CODE = 0

    def __init__(self, gpos=None, gvel=None, name=None):
        self.gpos = eye(4) if gpos is None else array(gpos, dtype=float).reshape((4, 4))
        self.gvel = zeros(6) if gvel is None else array(gvel, dtype=float).reshape(6)
        Joint.__init__(self, name)


def _linear(name, ndof, code, doc):
    return type(name, (LinearConfigurationSpaceJoint,),
                {"ndof": ndof, "TYPE_CODE": code, "__doc__": doc})


RzRyRxJoint = _linear("RzRyRxJoint", 3, 1, "Ball joint as three hinges, H = Rz.Ry.Rx (joints.py:59-104).")
RzRyJoint = _linear("RzRyJoint", 2, 2, "Two hinges, H = Rz.Ry (joints.py:107-146).")
RzRxJoint = _linear("RzRxJoint", 2, 3, "Two hinges, H = Rz.Rx (joints.py:149-185).")
RyRxJoint = _linear("RyRxJoint", 2, 4, "Two hinges, H = Ry.Rx (joints.py:188-224).")
RzJoint = _linear("RzJoint", 1, 5, "Hinge about z (joints.py:227-303).")
RyJoint = _linear("RyJoint", 1, 6, "Hinge about y (joints.py:305-326).")
RxJoint = _linear("RxJoint", 1, 7, "Hinge about x (joints.py:328-349).")
TxTyTzJoint = _linear("TxTyTzJoint", 3, 8, "Three prismatic axes, H = transl(q) (joints.py:352-384).")

JOINT_TYPE_CODES = {c.__name__: c.TYPE_CODE for c in (
    FreeJoint, RzRyRxJoint, RzRyJoint, RzRxJoint, RyRxJoint, RzJoint, RyJoint,
    RxJoint, TxTyTzJoint)}
JOINT_NDOF = {0: 6, 1: 3, 2: 2, 3: 2, 4: 2, 5: 1, 6: 1, 7: 1, 8: 3}
