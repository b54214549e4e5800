"""Host-side 6x6 spatial mass matrices for model building.

Same public names as the reference ``arboris/massmatrix.py`` (``transport`` :27,
``box`` :103, ``ellipsoid`` :124, ``cylinder`` :146, ``sphere`` :165).  Runs once
per model on the host; its outputs (body mass matrices) are constants of the
flattened model the CUDA kernels consume.  Twist/wrench order is [angular; linear].
"""
import numpy as np
from . import homogeneousmatrix as Hg


def ismassmatrix(M, semi=False):
    M = np.asarray(M)
    ok = (M.shape == (6, 6) and np.allclose(M, M.T)
          and np.allclose(M[3:, 3:], M[3, 3]*np.eye(3)))
    if not ok:
        return False
    ev = np.linalg.eigvalsh((M + M.T)/2)
    return bool((ev >= 0.).all() if semi else (ev > 0.).all())


def transport(M, H):
    """Express mass matrix ``M`` (frame a) in frame b, ``H = H_ab``."""
    assert ismassmatrix(M)
    Ad = Hg.adjoint(H)
    return Ad.T.dot(np.asarray(M).dot(Ad))


def _diag_inertia(ix, iy, iz, mass):
    return np.diag((ix, iy, iz, mass, mass, mass)).astype(float)


def box(half_extents, mass):
    x, y, z = half_extents
    k = mass/3.
    return _diag_inertia(k*(y**2 + z**2), k*(x**2 + z**2), k*(x**2 + y**2), mass)


def ellipsoid(radii, mass):
    x, y, z = radii
    k = mass/5.
    return _diag_inertia(k*(y**2 + z**2), k*(x**2 + z**2), k*(x**2 + y**2), mass)


def cylinder(length, radius, mass):
    """Homogeneous cylinder, symmetry axis along z."""
    itan = mass*(radius**2/4. + length**2/12.)
    return _diag_inertia(itan, itan, mass*radius**2/2., mass)


def sphere(radius, mass):
    i = 2.*mass*radius**2/5.
    return _diag_inertia(i, i, i, mass)
