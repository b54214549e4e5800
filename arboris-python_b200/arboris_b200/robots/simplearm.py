"""Planar 3R arm, same model as the reference ``arboris/robots/simplearm.py:20-78``:
three box links along +y hinged about z (Shoulder, Elbow, Wrist)."""
from ..core import World, Body, SubFrame
from .. import homogeneousmatrix as Hg
from .. import massmatrix
from .. import shapes as _shapes
from ..joints import RzJoint


def add_simplearm(world, name='', lengths=(0.5, 0.4, 0.2),
                  masses=(1.0, 0.8, 0.2), with_shapes=False):
    assert isinstance(world, World)
    anchor = world.ground
    specs = (('Arm', 'Shoulder', 'ElbowBaseFrame'),
             ('Forearm', 'Elbow', 'WristBaseFrame'),
             ('Hand', 'Wrist', 'EndEffector'))
    for (body_name, joint_name, tip_name), length, mass in zip(specs, lengths, masses):
        half_extents = (length/20., length/2., length/20.)
        # box inertia at the centre of mass, expressed at the link base
        m_base = massmatrix.transport(massmatrix.box(half_extents, mass),
                                      Hg.transl(0., -length/2., 0.))
        body = Body(name + body_name, m_base)
        if with_shapes:
            world.register(_shapes.Box(SubFrame(body, Hg.transl(0., length/2., 0.)),
                                       half_extents))
        world.add_link(anchor, RzJoint(name=name + joint_name), body)
        anchor = SubFrame(body, Hg.transl(0, length, 0), name + tip_name)
    world.register(anchor)
    world.init()
