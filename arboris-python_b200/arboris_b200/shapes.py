"""Collision shapes: plain data holders (reference ``arboris/shapes.py:9-68``)."""
import numpy

from .core import Shape


class Plane(Shape):
    """Plane ``a x + b y + c z + d = 0`` in the coordinates of ``frame``; the
    normal is normalised at construction (shapes.py:27-32)."""

    def __init__(self, frame, coeffs=(0., 1., 0., 0.), name=None):
        Shape.__init__(self, frame, name)
        coeffs = numpy.array(coeffs, dtype=float)
        self.coeffs = coeffs/numpy.linalg.norm(coeffs[0:3])


class Point(Shape):
    def __init__(self, frame, name=None):
        Shape.__init__(self, frame, name)


class Box(Shape):
    def __init__(self, frame, half_extents=(1., 1., 1.), name=None):
        Shape.__init__(self, frame, name)
        self.half_extents = half_extents


class Cylinder(Shape):
    def __init__(self, frame, length=1., radius=1., name=None):
        assert radius >= 0.
        Shape.__init__(self, frame, name)
        self.radius = radius
        self.length = length


class Sphere(Shape):
    def __init__(self, frame, radius=1., name=None):
        assert radius >= 0.
        Shape.__init__(self, frame, name)
        self.radius = radius
