"""Host-side SE(3) helpers used while *building* a model (constant frames).

Mirrors the public names of the reference module
``arboris/homogeneousmatrix.py`` (``transl`` :11, ``rotzyx`` :32 ... ``rotz`` :178,
``zaligned`` :201, ``ishomogeneousmatrix`` :234, ``pdot`` :242, ``vdot`` :248,
``inv`` :254, ``adjoint`` :277, ``iadjoint`` :321) so that robot factories written
against the reference run unchanged.  These run once per model on the host; the
per-step versions are ``__device__`` functions in ``csrc/arb_se3.cuh``.
"""
import numpy as np

tol = 1e-9  # reference homogeneousmatrix.py:9


def _h(R=None, p=None):
    H = np.eye(4)
    if R is not None:
        H[:3, :3] = R
    if p is not None:
        H[:3, 3] = p
    return H


def transl(t_x, t_y, t_z):
    return _h(p=(t_x, t_y, t_z))


def _rot1(axis, angle):
    c, s = np.cos(angle), np.sin(angle)
    i, j = [(1, 2), (2, 0), (0, 1)][axis]
    R = np.eye(3)
    R[i, i] = c
    R[j, j] = c
    R[i, j] = -s
    R[j, i] = s
    return R


def rotx(angle):
    return _h(_rot1(0, angle))


def roty(angle):
    return _h(_rot1(1, angle))


def rotz(angle):
    return _h(_rot1(2, angle))


def rotzyx(angle_z, angle_y, angle_x):
    sz, cz = np.sin(angle_z), np.cos(angle_z)
    sy, cy = np.sin(angle_y), np.cos(angle_y)
    sx, cx = np.sin(angle_x), np.cos(angle_x)
    # closed form of Rz.Ry.Rx with the same operation order as the reference
    R = np.array([[cz*cy, cz*sy*sx - sz*cx, cz*sy*cx + sz*sx],
                  [sz*cy, sz*sy*sx + cz*cx, sz*sy*cx - cz*sx],
                  [-sy, cy*sx, cy*cx]])
    return _h(R)


def rotzy(angle_z, angle_y):
    sz, cz = np.sin(angle_z), np.cos(angle_z)
    sy, cy = np.sin(angle_y), np.cos(angle_y)
    return _h(np.array([[cz*cy, -sz, cz*sy], [sz*cy, cz, sz*sy], [-sy, 0., cy]]))


def rotzx(angle_z, angle_x):
    sz, cz = np.sin(angle_z), np.cos(angle_z)
    sx, cx = np.sin(angle_x), np.cos(angle_x)
    return _h(np.array([[cz, -sz*cx, sz*sx], [sz, cz*cx, -cz*sx], [0., sx, cx]]))


def rotyx(angle_y, angle_x):
    sy, cy = np.sin(angle_y), np.cos(angle_y)
    sx, cx = np.sin(angle_x), np.cos(angle_x)
    return _h(np.array([[cy, sy*sx, sy*cx], [0., cx, -sx], [-sy, cy*sx, cy*cx]]))


def ishomogeneousmatrix(H, tol=tol):
    H = np.asarray(H)
    return bool(H.shape == (4, 4)
                and abs(np.linalg.det(H[:3, :3]) - 1.) <= tol
                and (H[3, :] == [0, 0, 0, 1]).all())


def inv(H):
    assert ishomogeneousmatrix(H)
    Rt = H[:3, :3].T
    return _h(Rt, -Rt.dot(H[:3, 3]))


def skew(v):
    return np.array([[0., -v[2], v[1]], [v[2], 0., -v[0]], [-v[1], v[0], 0.]])


def adjoint(H):
    assert ishomogeneousmatrix(H), H
    R = H[:3, :3]
    Ad = np.zeros((6, 6))
    Ad[:3, :3] = R
    Ad[3:, 3:] = R
    Ad[3:, :3] = skew(H[:3, 3]).dot(R)
    return Ad


def iadjoint(H):
    return adjoint(inv(H))


def pdot(H, point):
    assert ishomogeneousmatrix(H)
    return H[:3, :3].dot(point) + H[:3, 3]


def vdot(H, vec):
    assert ishomogeneousmatrix(H)
    return H[:3, :3].dot(vec)


def zaligned(vec):
    """Frame whose z axis is ``vec`` (reference ``homogeneousmatrix.py:201-232``).

    x is built from the stable ``argsort(|z|)`` index shuffle, y = z cross x.
    """
    vec = np.asarray(vec, dtype=float)
    assert abs(np.linalg.norm(vec) - 1) < 1e-9
    idx = np.argsort(np.absolute(vec), kind="stable")
    x = np.zeros(3)
    x[idx[1]] = vec[idx[2]]
    x[idx[2]] = -vec[idx[1]]
    x /= np.linalg.norm(x)
    H = np.eye(4)
    H[:3, 0] = x
    H[:3, 1] = np.cross(vec, x)
    H[:3, 2] = vec
    return H
