"""arboris_b200 -- B200-native batched simulation step behind the Arboris API.

Host Python describes the model with the reference's own class names
(``World``, ``Body``, joints, constraints, controllers, robots); the step
(update_dynamic -> update_controllers -> update_constraints -> integrate) runs
as hand-written fp64 CUDA kernels for sm_100a through the C ABI declared in
``include/arboris_b200.h``.  There is no CPU implementation of the step here.
"""
from .core import (World, Body, Joint, JointsList, NamedObjectsList, Frame,  # noqa: F401
                   SubFrame, MovingSubFrame, simulate, Constraint, Controller,
                   Observer, Shape, LinearConfigurationSpaceJoint)
from .flatten import flatten, FlatModel  # noqa: F401

__all__ = ['core', 'joints', 'shapes', 'constraints', 'controllers',
           'homogeneousmatrix', 'massmatrix', 'robots', 'flatten', 'batch']


def __getattr__(name):
    if name == "BatchedWorld":
        from .batch import BatchedWorld
        return BatchedWorld
    raise AttributeError(name)
