"""Anthropometric 36-dof humanoid on a free base (42 dof), same model as the
reference ``arboris/robots/human36.py`` (``anat_lengths_from_height`` :57-163,
``height_from_anat_lengths`` :165-185, ``add_human36`` :187-399; data from the
HuMAnS toolbox): 17 bodies, 17 joints, 28 anatomical tag frames and 8 ``Point``
shapes under the feet.  The model is table-driven here; names, ordering (hence
dof numbering) and every constant follow the reference so the flattened models
are identical (checked by ``tests/test_flatten.py``).
"""
import numpy as np

from ..core import World, Body, SubFrame, NamedObjectsList
from .. import homogeneousmatrix as Hg
from ..joints import (FreeJoint, RzRyRxJoint, RzRyJoint, RzRxJoint, RyRxJoint,
                      RzJoint)
from ..shapes import Point

_INCH = 0.0254

# length name -> fraction of the total height (side suffix added below)
_SIDED = (('yfoot', 0.0222), ('ytibia', 0.2493), ('yfemur', 0.2425),
          ('ysternoclav', 0.0980), ('xsternoclav', 0.1052), ('yshoulder', 0.0104),
          ('xshoulder', 0.0526), ('yhumerus', 0.1618), ('yforearm', 0.1544),
          ('yhand', 0.1091), ('xfoot', 0.1482), ('xheel', 0.0248))
_CENTRAL = (('yvT10', 0.2075), ('xvT10', 0.0526), ('zhip', 0.1002),
            ('yvC7', 0.139), ('yhead', 0.1395))


def anat_lengths_from_height(height):
    """Anatomical lengths (m) scaled from the body height, keyed as in HuMAnS."""
    L = {}
    for key, frac in _CENTRAL:
        L[key] = float(frac * height)
    zsternoclav = 0.5 * _INCH
    for side in 'RL':
        for key, frac in _SIDED:
            L[key + side] = float(frac * height)
        L['zsternoclav' + side] = float(zsternoclav)
        L['zshoulder' + side] = float(0.1295 * height - zsternoclav)
    return L


def height_from_anat_lengths(lengths):
    legs = [lengths['yfoot' + s] + lengths['ytibia' + s] + lengths['yfemur' + s]
            for s in 'RL']
    if legs[0] != legs[1]:
        raise ValueError("The legs have different lengths")
    return legs[1] + lengths['yvT10'] + lengths['yvC7'] + lengths['yhead']


def _segments(L):
    """(name, mass fraction, centre of mass, radii of gyration) per body."""
    def limb(side):
        s = side
        return [
            ('Thigh' + s, 0.1416, (0, -0.4095*L['yfemur' + s], 0),
             np.array([0.329, 0.149, 0.329])*L['yfemur' + s]),
            ('Shank' + s, 0.0433, (0, -0.4459*L['ytibia' + s], 0),
             np.array([0.255, 0.103, 0.249])*L['ytibia' + s]),
            ('Foot' + s, 0.0137,
             (0.4415*L['xfoot' + s] - L['xheel' + s], -L['yfoot' + s]/2., 0.),
             np.array([0.124, 0.257, 0.245])*L['xfoot' + s]),
        ]

    def arm(side, hand_gyr):
        s = side
        return [
            ('Scapula' + s, 0., (0., 0., 0.), np.zeros(3)),
            ('Arm' + s, 0.0271, (0., -0.5772*L['yhumerus' + s], 0.),
             np.array([0.285, 0.158, 0.269])*L['yhumerus' + s]),
            ('Forearm' + s, 0.0162, (0., -0.4574*L['yforearm' + s], 0.),
             np.array([0.276, 0.121, 0.265])*L['yforearm' + s]),
            ('Hand' + s, 0.0061, (0, -0.3691*L['yhand' + s], 0),
             np.array(hand_gyr)*L['yhand' + s]),
        ]

    segs = [('LPT', 0.275, (0, 0.5108*L['yvT10'], 0),
             np.array([0.2722, 0.2628, 0.226])*L['yvT10'])]
    segs += limb('R') + limb('L')
    segs.append(('UPT', 0.1596,
                 ((L['xsternoclavR'] + L['xsternoclavL'])/4.,
                  0.7001*(L['ysternoclavR'] + L['ysternoclavL'])/2., 0.),
                 np.array([0.716, 0.659, 0.454])*L['ysternoclavR']))
    segs += arm('R', (0.235, 0.184, 0.288)) + arm('L', (0.288, 0.184, 0.235))
    segs.append(('Head', 0.0694, (0, 0.4998*L['yhead'], 0),
                 np.array([0.303, 0.261, 0.315])*L['yhead']))
    return segs


def _links(L):
    """(parent body or None for ground, anchor translation, joint class, child)."""
    def leg(s, sign):
        return [('LPT', (0, 0, sign*L['zhip']/2.), RzRyRxJoint, 'Thigh' + s),
                ('Thigh' + s, (0, -L['yfemur' + s], 0), RzJoint, 'Shank' + s),
                ('Shank' + s, (0, -L['ytibia' + s], 0), RzRxJoint, 'Foot' + s)]

    def arm(s, sign):
        return [('UPT', (L['xsternoclav' + s], L['ysternoclav' + s],
                         sign*L['zsternoclav' + s]), RyRxJoint, 'Scapula' + s),
                ('Scapula' + s, (-L['xshoulder' + s], L['yshoulder' + s],
                                 sign*L['zshoulder' + s]), RzRyRxJoint, 'Arm' + s),
                ('Arm' + s, (0, -L['yhumerus' + s], 0), RzRyJoint, 'Forearm' + s),
                ('Forearm' + s, (0, -L['yforearm' + s], 0), RzRxJoint, 'Hand' + s)]

    links = [(None, (0, L['yfootL'] + L['ytibiaL'] + L['yfemurL'], 0), FreeJoint, 'LPT')]
    links += leg('R', 1.) + leg('L', -1.)
    links.append(('LPT', (-L['xvT10'], L['yvT10'], 0), RzRyRxJoint, 'UPT'))
    links += arm('R', 1.) + arm('L', -1.)
    links.append(('UPT', (L['xvT10'], L['yvC7'], 0), RzRyRxJoint, 'Head'))
    return links


def _tags(L, h):
    """(tag name, body, position in the body frame), HuMAnS landmark names."""
    def foot(side_word, s, zsign, toe5, toe1):
        return [
            (side_word + ' foot toe tip', 'Foot' + s,
             [L['xfoot' + s] - L['xheel' + s] + 1e-4*h, -L['yfoot' + s], 0.]),
            (side_word + ' foot heel', 'Foot' + s, [-L['xheel' + s], -L['yfoot' + s], 0.]),
            (side_word + ' foot ' + toe5, 'Foot' + s,
             [0.0662*h, -L['yfoot' + s], zsign*0.0305*h]),
            (side_word + ' foot ' + toe1, 'Foot' + s,
             [0.0662*h, -L['yfoot' + s], -zsign*0.0305*h]),
            (side_word + ' foot lateral malleolus', 'Shank' + s,
             [0., -L['ytibia' + s], zsign*0.0249*h]),
        ]

    tags = foot('Right', 'R', 1., 'phalange 5', 'Phalange 1')
    tags += [('Femoral lateral epicondyle', 'ThighR', [0., -L['yfemurR'], 0.0290*h]),
             ('Right great trochanter', 'ThighR', [0., 0., 0.0941*h - L['zhip']/2.]),
             ('Right iliac crest', 'LPT', [0.0271*h, 0.0366*h, 0.0697*h])]
    tags += foot('Left', 'L', -1., 'phalange 5', 'phalange 1')
    tags += [('Left femoral lateral epicondyle', 'ThighL', [0, -L['yfemurL'], -0.0290*h]),
             ('Left great trochanter', 'ThighL', [0, 0, -0.0941*h + L['zhip']/2.]),
             ('Left iliac crest', 'LPT', [0.0271*h, 0.0366*h, -0.0697*h]),
             ('Substernale (Xyphoid)', 'UPT', [0.1219*h, 0, 0]),
             ('Suprasternale', 'UPT', [(L['xsternoclavL'] + L['xsternoclavL'])/2.,
                                       (L['ysternoclavL'] + L['ysternoclavL'])/2., 0])]
    for word, s, sg in (('Right', 'R', 1.), ('Left', 'L', -1.)):
        tags += [(word + ' acromion', 'Scapula' + s,
                  [-L['xshoulder' + s], 0.0198*h + L['yshoulder' + s],
                   sg*L['zshoulder' + s]]),
                 (word + ' humeral lateral epicondyle (radiale)', 'Arm' + s,
                  [0., -L['yhumerus' + s], sg*0.0211*h]),
                 (word + ' stylion', 'Forearm' + s, [0., -0.1533*h, sg*0.0331*h]),
                 (word + ' 3rd dactylion', 'Hand' + s, [0., -L['yhand' + s], 0.])]
    tags += [('Cervicale', 'UPT', [-0.0392*0. + L['xvT10'], L['yvC7'], 0.]),
             ('Vertex', 'Head', [0., L['yhead'], 0.])]
    return tags


_FOOT_POINTS = ('Right foot toe tip', 'Right foot heel', 'Right foot phalange 5',
                'Right foot Phalange 1', 'Left foot toe tip', 'Left foot heel',
                'Left foot phalange 5', 'Left foot phalange 1')


def add_human36(world, height=1.741, mass=73, anat_lengths=None, name=''):
    """Add the humanoid to ``world`` (prefixing every object name with ``name``)."""
    assert isinstance(world, World)
    L = anat_lengths_from_height(height) if anat_lengths is None else anat_lengths
    h = height_from_anat_lengths(L)

    bodies = NamedObjectsList()
    for seg_name, frac, com, gyration in _segments(L):
        seg_mass = frac * mass
        inertia_com = seg_mass * np.diag(np.hstack((np.asarray(gyration)**2, (1, 1, 1))))
        H_fg = np.eye(4)
        H_fg[0:3, 3] = com
        Ad = Hg.adjoint(Hg.inv(H_fg))
        bodies.append(Body(name=name + seg_name,
                           mass=np.dot(Ad.T, np.dot(inertia_com, Ad))))

    for parent, anchor, joint_class, child in _links(L):
        parent_body = world.ground if parent is None else bodies[name + parent]
        world.add_link(SubFrame(parent_body, Hg.transl(*anchor)), joint_class(),
                       bodies[name + child])

    tags = NamedObjectsList()
    for tag_name, body, position in _tags(L, h):
        tag = SubFrame(bodies[name + body], Hg.transl(*position), name + tag_name)
        tags.append(tag)
        world.register(tag)

    for key in _FOOT_POINTS:
        world.register(Point(tags[name + key], name=name + key))
    world.init()
