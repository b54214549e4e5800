"""CPU build (g++) of the very routines the CUDA kernels run per world
(arboris-python_b200/csrc/arb_world.cuh and friends), checked against the
real-reference trajectories and against numpy.linalg for the small dense kernels.
No GPU needed; the GPU tests (-m gpu) repeat the trajectory checks through the
C ABI on the device."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import load_golden

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hosttest"))
import harness  # noqa: E402


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max()/max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name,tol", [("simplearm", 1e-12), ("human36_free", 1e-10),
                                      ("ball_socket", 1e-12), ("simplearm_limits", 1e-12),
                                      ("snake_loop", 1e-10), ("human36_contact", 1e-10), ("balls", 1e-10), ("zoo", 1e-10)])
def test_world_routines_vs_real_reference(name, tol):
    """per step from identical states: M, N, Z, Y, forces, velocities <= tol relative;
    active sets and solver branches bit-exact."""
    model, tr = load_golden(name)
    n, dt = model.ndof, float(tr["dt"])
    W, T = tr["gpos"].shape[:2]
    hb = harness.HostBatch(model, W)
    gpos, gvel = tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy()
    cf = np.zeros((max(model.nrows, 1), W))
    fs = list(tr["full_steps"])
    flips, worst = 0, {}
    for s in range(T):
        hb.gpos[:], hb.gvel[:], hb.cforce[:] = gpos, gvel, cf
        hb.update_dynamic()
        hb.update_controllers(dt)
        if s in fs:
            i = fs.index(s)
            for k, nm in (("mass", "M"), ("nleffects", "N"), ("impedance", "Z"), ("admittance", "Y")):
                got = np.moveaxis(hb.arr(nm, n, n), -1, 0)
                worst[k] = max(worst.get(k, 0), rel(got, tr[k][:, i]))
        hb.update_constraints(dt)
        if model.nc:
            a = hb.iarr("cactive", model.nc).T
            br = hb.iarr("cbranch", model.nc).T
            flips += int((a != tr["active"][:, s]).sum()) + int((br*a != tr["branch"][:, s]).sum())
            if name != "ball_socket":
                worst["cforce"] = max(worst.get("cforce", 0),
                                      rel(hb.cforce.T[:, :model.nrows], tr["cforce"][:, s]))
        hb.integrate(dt)
        worst["gvel"] = max(worst.get("gvel", 0), rel(hb.gvel.T, tr["gvel"][:, s]))
        worst["gpos"] = max(worst.get("gpos", 0), rel(hb.gpos.T, tr["gpos"][:, s]))
        gpos, gvel = tr["gpos"][:, s].T.copy(), tr["gvel"][:, s].T.copy()
        if model.nrows:
            cf = tr["cforce"][:, s].T.copy()
    assert flips == 0
    assert not hb.iarr("status").any()
    for k, v in worst.items():
        assert v <= tol, (k, v)


def test_pinv_small_vs_numpy():
    L = harness.lib()
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 4):
        for trial in range(50):
            a = rng.normal(size=(n, n))
            full = True
            if trial % 5 == 0 and n > 1:      # rank deficient
                a[:, -1] = a[:, 0]*2.
                full = False
            if trial % 7 == 0:
                a *= 1e-6
            out = np.zeros((n, n))
            L.ht_pinv(n, a.ctypes.data_as(dp), out.ctypes.data_as(dp))
            ref = np.linalg.pinv(a)
            assert rel(out, ref) < 1e-11*(np.linalg.cond(a) if full else 1e3), (n, trial)
    z, out = np.zeros((4, 4)), np.ones((4, 4))
    L.ht_pinv(4, z.ctypes.data_as(dp), out.ctypes.data_as(dp))
    assert not out.any()


def test_eig6_vs_numpy():
    L = harness.lib()
    rng = np.random.default_rng(1)
    dp = C.POINTER(C.c_double)
    for trial in range(300):
        a = rng.normal(size=(6, 6))
        if trial % 3 == 0:
            a = a*np.logspace(-3, 3, 6)[None, :]
        if trial % 4 == 0:   # shape of the sliding-friction matrix (constraints.py:815-821)
            s = rng.normal(size=(3, 3))
            s = s + s.T
            a[:3, :3] = s + rng.normal()
            a[3:, 3:] = s
            a[:3, 3:] = -np.eye(3)*abs(rng.normal())
            a[3:, :3] = np.eye(3)*rng.normal()
        wr, wi = np.zeros(6), np.zeros(6)
        ok = L.ht_eig6(a.ctypes.data_as(dp), wr.ctypes.data_as(dp), wi.ctypes.data_as(dp))
        assert ok
        ref = np.sort_complex(np.linalg.eigvals(a))
        got = np.sort_complex(wr + 1j*wi)
        assert np.abs(got - ref).max() < 1e-8*max(1., np.abs(ref).max()), (trial, got, ref)
        assert (wi == 0).sum() == (ref.imag == 0).sum()


def test_solve4_and_exp():
    L = harness.lib()
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(2)
    L.ht_solve4.restype = C.c_int
    for _ in range(50):
        a, b, x = rng.normal(size=(4, 4)), rng.normal(size=4), np.zeros(4)
        assert L.ht_solve4(a.ctypes.data_as(dp), b.ctypes.data_as(dp), x.ctypes.data_as(dp))
        assert rel(x, np.linalg.solve(a, b)) < 1e-12*np.linalg.cond(a)
    from oracle.arboris_oracle import twist_exp
    for scale in (1., 1e-2, 1e-4, 0.):
        tw = rng.normal(size=6)*np.array([scale]*3 + [1.]*3)
        out = np.zeros(12)
        L.ht_exp(tw.ctypes.data_as(dp), out.ctypes.data_as(dp))
        H = twist_exp(tw)
        assert np.abs(out[:9].reshape(3, 3) - H[:3, :3]).max() < 1e-14
        assert np.abs(out[9:] - H[:3, 3]).max() < 1e-13*max(1., np.abs(H[:3, 3]).max())


@pytest.mark.parametrize("name,tol", [("simplearm", 1e-12), ("human36_free", 1e-10),
                                      ("ball_socket", 1e-12), ("simplearm_limits", 1e-12),
                                      ("snake_loop", 1e-10), ("human36_contact", 1e-10), ("balls", 1e-10), ("zoo", 1e-10)])
@pytest.mark.parametrize("coop", [1, 0])
def test_fused_algorithm_vs_real_reference(name, tol, coop):
    """The fused step's algorithm in scalar form (no-fill tree factorisation of Z instead of
    the explicit inverse, Gauss-Seidel in generator space) against the real reference."""
    model, tr = load_golden(name)
    dt = float(tr["dt"])
    W, T = tr["gpos"].shape[:2]
    hb = harness.HostBatch(model, W)
    hb.set_coop(coop)
    gpos, gvel = tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy()
    cf = np.zeros((max(model.nrows, 1), W))
    flips, worst = 0, {}
    for s in range(T):
        hb.gpos[:], hb.gvel[:], hb.cforce[:] = gpos, gvel, cf
        hb.fused_step(dt)
        if model.nc:
            a = hb.iarr("factive", model.nc).T
            br = hb.iarr("fbranch", model.nc).T
            flips += int((a != tr["active"][:, s]).sum()) + int((br*a != tr["branch"][:, s]).sum())
            if name != "ball_socket":
                worst["cforce"] = max(worst.get("cforce", 0),
                                      rel(hb.cforce.T[:, :model.nrows], tr["cforce"][:, s]))
        worst["gvel"] = max(worst.get("gvel", 0), rel(hb.gvel.T, tr["gvel"][:, s]))
        worst["gpos"] = max(worst.get("gpos", 0), rel(hb.gpos.T, tr["gpos"][:, s]))
        gpos, gvel = tr["gpos"][:, s].T.copy(), tr["gvel"][:, s].T.copy()
        if model.nrows:
            cf = tr["cforce"][:, s].T.copy()
    assert flips == 0
    assert not hb.iarr("status").any()
    for k, v in worst.items():
        assert v <= tol, (k, v)


def _sliding_matrix(A, alpha, mu):
    """B of SoftFingerContact.solve exactly as the reference builds it (constraints.py:807-821,
    eps = 1): note the scalar inner products."""
    Y_c, y_n, Y_t = A[0:3, 3], A[3, 3], A[0:3, 0:3]
    beta = alpha[0:3] - alpha[3]/y_n*Y_c
    a = mu/y_n*alpha[3]
    b = mu/y_n*Y_c
    B = np.zeros((6, 6))
    E = np.eye(3)
    Y_that = Y_t - np.dot(Y_c, Y_c.T)/y_n
    B[3:6, 3:6] = np.dot(E, Y_that)
    B[0:3, 0:3] = np.dot(E, Y_that + 2/a*np.dot(beta, b.T))
    B[0:3, 3:6] = -np.dot(E, np.dot(beta, beta.T)/(a**2))
    B[3:6, 0:3] = np.dot(E, np.dot(b, b.T)) - np.eye(3)
    return B


def test_structured_sliding_root_vs_numpy_eigvals():
    """The structured replacement of eigvals() in the sliding branch (constraints.py:825-830):
    same admissible root as LAPACK on admittance-like blocks over several scales."""
    L = harness.lib()
    dp = C.POINTER(C.c_double)
    L.ht_sliding_root.argtypes = [dp, dp, C.c_double, dp, C.POINTER(C.c_int)]
    rng = np.random.default_rng(7)
    nfound = nnone = 0
    for trial in range(2000):
        G = rng.normal(size=(4, 6))
        A = G.dot(G.T)*10.**rng.uniform(-4, 1)                 # SPD like J Y J^T
        A = A + rng.normal(size=(4, 4))*1e-2*np.abs(A).max()   # the N term makes it non-symmetric
        alpha = rng.normal(size=4)*10.**rng.uniform(-3, 1)
        mu = rng.uniform(0.1, 1.5)
        S = np.linalg.eigvals(_sliding_matrix(A, alpha, mu))
        S = S[np.logical_and(S.imag == 0, S.real <= 0)]
        s, found = C.c_double(0.), C.c_int(0)
        ok = L.ht_sliding_root(np.ascontiguousarray(A).ctypes.data_as(dp), alpha.ctypes.data_as(dp), mu,
                               C.byref(s), C.byref(found))
        assert ok
        if len(S) == 0:
            nnone += 1
            assert not found.value, (trial, s.value)
        else:
            nfound += 1
            ref = float(min(S).real)
            assert found.value, (trial, ref)
            assert abs(s.value - ref) <= 1e-9*abs(ref) + 1e-13*np.abs(A).max(), (trial, s.value, ref)
    assert nfound > 500 and nnone > 10


def test_contact_aligned_blocks_are_detected():
    """human36 on the ground plane: both feet are contact-aligned generator bodies (their 8
    plane/point contacts become translations in the Gauss-Seidel); bodies that carry a
    ball-and-socket constraint are not."""
    model, _ = load_golden("human36_contact")
    g, c = harness.HostBatch(model, 1).aligned()
    assert g.tolist() == [1, 1]
    assert c.tolist() == [1]*8 + [0, 0]
    for name in ("snake_loop", "ball_socket"):
        model, _ = load_golden(name)
        g, c = harness.HostBatch(model, 1).aligned()
        assert not g.any() and not c.any()


def _extra(model_name, file_name):
    import os
    from conftest import GOLDEN
    from arboris_b200.flatten import FlatModel
    model = FlatModel.load(os.path.join(GOLDEN, "model_%s.npz" % model_name))
    with np.load(os.path.join(GOLDEN, file_name)) as z:
        return model, {k: z[k] for k in z.files}


def test_fused_algorithm_vs_contact64():
    """SURVEY.md 8(d) config 3 parity subset: 64 falling humanoids, 120 steps of the real reference;
    the fused algorithm (host build of the device routines) from the reference's state at every step."""
    model, tr = _extra("human36_contact", "traj_human36_contact64.npz")
    dt = float(tr["dt"])
    W, T = tr["gpos"].shape[:2]
    hb = harness.HostBatch(model, W)
    hb.set_coop(0)
    gpos, gvel = tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy()
    cf = np.zeros((model.nrows, W))
    flips, worst = 0, {}
    for s in range(T):
        hb.gpos[:], hb.gvel[:], hb.cforce[:] = gpos, gvel, cf
        hb.fused_step(dt)
        a = hb.iarr("factive", model.nc).T
        br = hb.iarr("fbranch", model.nc).T
        flips += int((a != tr["active"][:, s]).sum()) + int((br*a != tr["branch"][:, s]).sum())
        for k, x in (("cforce", hb.cforce.T[:, :model.nrows]), ("gvel", hb.gvel.T), ("gpos", hb.gpos.T)):
            worst[k] = max(worst.get(k, 0), rel(x, tr[k][:, s]))
        gpos, gvel, cf = tr["gpos"][:, s].T.copy(), tr["gvel"][:, s].T.copy(), tr["cforce"][:, s].T.copy()
    assert flips == 0
    for k, v in worst.items():
        assert v <= 1e-10, (k, v)


@pytest.mark.parametrize("path", ["fused", "phases"])
def test_per_world_pd_parameters_host(path):
    """per-world kp, kd, gpos_des, gvel_des (arb_batch_bind_controller_params) against the real
    reference stepped once per world with those values (tests/golden/traj_zoo_pd.npz)."""
    model, tr = _extra("zoo", "traj_zoo_pd.npz")
    dt = float(tr["dt"])
    W, T = tr["gpos"].shape[:2]
    hb = harness.HostBatch(model, W)
    hb.set_coop(0)
    hb.bind_params(tr["pd_kp"], tr["pd_kd"], tr["pd_gpos_des"], tr["pd_gvel_des"])
    gpos, gvel = tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy()
    cf = np.zeros((max(model.nrows, 1), W))
    worst = {}
    for s in range(T):
        hb.gpos[:], hb.gvel[:], hb.cforce[:] = gpos, gvel, cf
        if path == "fused":
            hb.fused_step(dt)
        else:
            hb.update_dynamic(); hb.update_controllers(dt); hb.update_constraints(dt); hb.integrate(dt)
        worst["gvel"] = max(worst.get("gvel", 0), rel(hb.gvel.T, tr["gvel"][:, s]))
        worst["gpos"] = max(worst.get("gpos", 0), rel(hb.gpos.T, tr["gpos"][:, s]))
        gpos, gvel = tr["gpos"][:, s].T.copy(), tr["gvel"][:, s].T.copy()
        if model.nrows:
            cf = tr["cforce"][:, s].T.copy()
    for k, v in worst.items():
        assert v <= 1e-10, (k, v)


@pytest.mark.parametrize("name,tol", [("simplearm", 1e-12), ("human36_free", 1e-10), ("ball_socket", 1e-12),
                                      ("simplearm_limits", 1e-12), ("snake_loop", 1e-10),
                                      ("human36_contact", 1e-10), ("balls", 1e-10), ("zoo", 1e-10)])
def test_group_prepare_vs_real_reference(name, tol):
    """The prepare stage with 16 lanes per world and on-chip scratch (arb_group.cuh; the lanes of a
    group are emulated one after the other between the barriers) + the K-matrix finish stage,
    against the real reference's trajectories, from the reference's state at every step."""
    model, tr = load_golden(name)
    dt = float(tr["dt"])
    W, T = tr["gpos"].shape[:2]
    T = min(T, 150)
    hb = harness.HostBatch(model, W)
    gpos, gvel = tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy()
    cf = np.zeros((max(model.nrows, 1), W))
    flips, worst = 0, {}
    for s in range(T):
        hb.gpos[:], hb.gvel[:], hb.cforce[:] = gpos, gvel, cf
        hb.fused_step_group(dt)
        if model.nc:
            a = hb.iarr("factive", model.nc).T
            br = hb.iarr("fbranch", model.nc).T
            flips += int((a != tr["active"][:, s]).sum()) + int((br*a != tr["branch"][:, s]).sum())
            worst["cforce"] = max(worst.get("cforce", 0), rel(hb.cforce.T[:, :model.nrows], tr["cforce"][:, s]))
        worst["gvel"] = max(worst.get("gvel", 0), rel(hb.gvel.T, tr["gvel"][:, s]))
        worst["gpos"] = max(worst.get("gpos", 0), rel(hb.gpos.T, tr["gpos"][:, s]))
        gpos, gvel = tr["gpos"][:, s].T.copy(), tr["gvel"][:, s].T.copy()
        if model.nrows:
            cf = tr["cforce"][:, s].T.copy()
    assert flips == 0
    assert not hb.iarr("status").any()
    for k, v in worst.items():
        assert v <= tol, (k, v)


def _poly6_lib():
    L = harness.lib()
    dp = C.POINTER(C.c_double)
    L.ht_poly6_positive.argtypes = [dp]
    L.ht_poly6_fast.argtypes = [dp, dp]
    L.ht_poly6_slow.argtypes = [dp, dp]
    return L, dp


def _random_sextic(rng, nreal):
    """Monic sextic with `nreal` real roots (0, 2, 4 or 6) and conjugate pairs for the rest, scaled like
    the sliding-friction polynomials (coefficients O(1))."""
    roots = list(rng.uniform(-2., 2., size=nreal))
    for _ in range((6 - nreal)//2):
        re, im = rng.uniform(-2., 2.), rng.uniform(0.05, 2.)
        roots += [complex(re, im), complex(re, -im)]
    c = np.real(np.poly(roots))          # highest degree first, c[0] = 1
    return np.ascontiguousarray(c[::-1][:6]), np.array(roots)


def test_positivity_march_never_claims_a_polynomial_with_a_root():
    """poly6_positive_on_halfline (the proof that the sliding problem has no admissible root, the
    reference's s = -1e10 clamp) is a sufficient test: it must never hold for a sextic with a root t >= 0,
    and it should hold for the clearly positive ones."""
    L, dp = _poly6_lib()
    rng = np.random.default_rng(11)
    proven = total_pos = 0
    for trial in range(4000):
        p, roots = _random_sextic(rng, int(rng.choice([0, 2, 4, 6])))
        real = roots[np.abs(np.imag(roots)) == 0].real
        has_root = bool((real >= 0.).any())
        ok = L.ht_poly6_positive(p.ctypes.data_as(dp))
        if has_root:
            assert not ok, (trial, p, real)
        else:
            total_pos += 1
            proven += ok
    assert total_pos > 500 and proven > 0.9*total_pos, (proven, total_pos)


def test_fast_root_with_bracket_recovery_is_the_largest_real_root():
    """poly6_largest_root_fast (Laguerre from the Samuelson bound, an iterate left of a root refined in
    its bracket, certificate): whenever it answers, the answer is the largest real root -- checked
    against numpy.roots and against the rigorous isolation."""
    L, dp = _poly6_lib()
    rng = np.random.default_rng(12)
    answered = 0
    for trial in range(4000):
        p, roots = _random_sextic(rng, int(rng.choice([2, 4, 6, 6])))
        r = np.roots(np.concatenate(([1.], p[::-1])))
        real = np.sort(r[np.abs(r.imag) < 1e-9*np.maximum(1., np.abs(r))].real)
        t = C.c_double(0.)
        if L.ht_poly6_fast(p.ctypes.data_as(dp), C.byref(t)):
            answered += 1
            assert len(real) > 0 and abs(t.value - real[-1]) <= 1e-7*max(1., abs(real[-1])), (trial, t.value, real)
        ts = C.c_double(0.)
        nr = L.ht_poly6_slow(p.ctypes.data_as(dp), C.byref(ts))
        if nr and len(real) and real[-1] > 1e-6:
            assert abs(ts.value - real[-1]) <= 1e-7*max(1., abs(real[-1])), (trial, ts.value, real)
    assert answered > 2000
