"""Planar snake: a chain of cylinders hinged about z, optionally on a free
base (reference ``arboris/robots/snake.py:17-61``)."""
from ..core import World, Body, SubFrame
from ..massmatrix import transport, cylinder, box
from ..homogeneousmatrix import transl
from ..joints import FreeJoint, RzJoint


def add_snake(w, nbody, lengths=None, masses=None, gpos=None, gvel=None,
              is_fixed=True):
    assert isinstance(w, World)
    lengths = [.5]*nbody if lengths is None else lengths
    masses = [2.]*nbody if masses is None else masses
    gpos = [0.]*nbody if gpos is None else gpos
    gvel = [0.]*nbody if gvel is None else gvel
    for seq in (lengths, masses, gpos, gvel):
        assert nbody == len(seq)
    frame = w.ground
    if not is_fixed:
        half = lengths[0]/2.
        base = Body(mass=box([half, half, half], masses[0]))
        w.add_link(w.ground, FreeJoint(), base)
        frame = base
    for length, mass, q, dq in zip(lengths, masses, gpos, gvel):
        link = Body(mass=transport(cylinder(length, length/10., mass),
                                   transl(0., -length/2., 0.)))
        w.add_link(frame, RzJoint(gpos=float(q), gvel=float(dq)), link)
        frame = SubFrame(link, transl(0., length, 0.))
    w.register(frame)
    w.init()
