"""Controllers (reference ``arboris/controllers.py``).

``WeightController`` (controllers.py:10-60): gravity as a generalized force,
zero impedance -- evaluated on the device (``arb_update_controllers``).
``ProportionalDerivativeController`` (controllers.py:63-159) is described here
so models carrying one flatten; its device evaluation is SURVEY.md section 8(f) row 2.
"""
from numpy import array, zeros

from .core import Controller, LinearConfigurationSpaceJoint


class WeightController(Controller):
    def __init__(self, gravity=-9.81, name=None):
        self.gravity = float(gravity)
        Controller.__init__(self, name=name)


class ProportionalDerivativeController(Controller):
    def __init__(self, joints, kp=None, kd=None, gpos_des=None, gvel_des=None,
                 name=None):
        Controller.__init__(self, name=name)
        self.joints = list(joints)
        n = 0
        for j in self.joints:
            if not isinstance(j, LinearConfigurationSpaceJoint):
                raise ValueError('Joints must be LinearConfigurationSpaceJoint instances')
            n += j.ndof
        self._cndof = n
        self.kp = zeros((n, n)) if kp is None else array(kp, dtype=float).reshape((n, n))
        self.kd = zeros((n, n)) if kd is None else array(kd, dtype=float).reshape((n, n))
        self.gpos_des = zeros(n) if gpos_des is None else array(gpos_des, dtype=float).reshape(n)
        self.gvel_des = zeros(n) if gvel_des is None else array(gvel_des, dtype=float).reshape(n)
