"""Parity of the CUDA path (through the C ABI) with the reference: per step from
identical states against trajectories recorded from the REAL reference
(tests/golden) and against the oracle on seeded random states.

Tolerances are BASELINE.json's: M, N and generalized velocities within 1e-10
relative per step; trajectories within 1e-6 after 1000 steps; contact active
sets, solver branches and argsort indices bit-exact."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

REL_TOL = 1e-10


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max()/max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("GPU tests need a CUDA device")
    return torch


def _batch(model, W):
    from arboris_b200.batch import BatchedWorld
    return BatchedWorld(model, W, device="cuda:0")


@pytest.mark.parametrize("name", ["simplearm", "human36_free", "ball_socket",
                                  "simplearm_limits", "snake_loop", "human36_contact", "balls", "zoo"])
def test_phases_vs_real_reference(torch_cuda, name):
    model, tr = load_golden(name)
    n, dt = model.ndof, float(tr["dt"])
    W, T = tr["gpos"].shape[:2]
    T = min(T, 300)
    bw = _batch(model, W)
    gpos, gvel = tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy()
    cf = np.zeros((max(model.nrows, 1), W))
    fs = list(tr["full_steps"])
    flips, worst = 0, {}
    for s in range(T):
        bw.set_state(gpos, gvel, cf)
        bw.update_dynamic()
        bw.update_controllers(dt)
        # a10: World._gforce after the controllers (WeightController / PD generalized force)
        if np.abs(tr["gforce_ctrl"][:, s]).max() > 0:
            worst["gforce_ctrl"] = max(worst.get("gforce_ctrl", 0),
                                       rel(bw.gforce().cpu().numpy(), tr["gforce_ctrl"][:, s]))
        if s in fs:
            i = fs.index(s)
            for k in ("mass", "nleffects", "impedance", "admittance"):
                worst[k] = max(worst.get(k, 0), rel(bw.matrix(k).cpu().numpy(), tr[k][:, i]))
        bw.update_constraints(dt)
        if model.nc:
            a = bw.constraints("active").cpu().numpy()
            br = bw.constraints("branch").cpu().numpy()
            flips += int((a != tr["active"][:, s]).sum()) + int((br*a != tr["branch"][:, s]).sum())
            worst["cforce"] = max(worst.get("cforce", 0),
                                  rel(bw.cforce.cpu().numpy().T[:, :model.nrows], tr["cforce"][:, s]))
        bw.integrate(dt)
        g, v, _ = bw.get_state()
        worst["gvel"] = max(worst.get("gvel", 0), rel(v.T, tr["gvel"][:, s]))
        worst["gpos"] = max(worst.get("gpos", 0), rel(g.T, tr["gpos"][:, s]))
        gpos, gvel = tr["gpos"][:, s].T.copy(), tr["gvel"][:, s].T.copy()
        if model.nrows:
            cf = tr["cforce"][:, s].T.copy()
    assert flips == 0, "active set / branch flips: %d" % flips
    assert int(bw.status().max()) == 0
    for k, v in worst.items():
        assert v <= REL_TOL, (k, v)


@pytest.mark.parametrize("name", ["simplearm", "human36_free", "ball_socket",
                                  "simplearm_limits", "snake_loop", "human36_contact", "balls", "zoo"])
def test_fused_step_vs_real_reference(torch_cuda, name):
    """arb_step (the fused path) from the reference's state at every step."""
    model, tr = load_golden(name)
    dt = float(tr["dt"])
    W, T = tr["gpos"].shape[:2]
    T = min(T, 300)
    bw = _batch(model, W)
    gpos, gvel = tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy()
    cf = np.zeros((max(model.nrows, 1), W))
    worst, flips = {}, 0
    for s in range(T):
        bw.set_state(gpos, gvel, cf)
        bw.step(dt, 1)
        g, v, f = bw.get_state()
        if model.nc:      # active sets, solver branches and signed distances of the FUSED path
            a = bw.constraints("active").cpu().numpy()
            br = bw.constraints("branch").cpu().numpy()
            flips += int((a != tr["active"][:, s]).sum()) + int((br*a != tr["branch"][:, s]).sum())
            if "sdist" in tr and name == "human36_contact":
                sd = bw.constraints("sdist").cpu().numpy()[:, :8]
                worst["sdist"] = max(worst.get("sdist", 0), np.abs(sd - tr["sdist"][:, s][:, :8]).max()*1e-2)
        worst["gvel"] = max(worst.get("gvel", 0), rel(v.T, tr["gvel"][:, s]))
        worst["gpos"] = max(worst.get("gpos", 0), rel(g.T, tr["gpos"][:, s]))
        if model.nrows:
            worst["cforce"] = max(worst.get("cforce", 0), rel(f.T, tr["cforce"][:, s]))
        gpos, gvel = tr["gpos"][:, s].T.copy(), tr["gvel"][:, s].T.copy()
        if model.nrows:
            cf = tr["cforce"][:, s].T.copy()
    assert flips == 0, "active set / branch flips on the fused path: %d" % flips
    assert int(bw.status().max()) == 0
    for k, v in worst.items():
        assert v <= REL_TOL, (k, v)


def test_simplearm_1000_steps_free_running(torch_cuda):
    """configs[0]: 1000 steps at dt = 1e-3, free running; final q within 1e-6."""
    model, tr = load_golden("simplearm")
    bw = _batch(model, 1)
    bw.set_state(tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy())
    bw.step(1e-3, 1000)
    g, v, _ = bw.get_state()
    assert np.abs(g[:, 0] - tr["gpos"][0, -1]).max() < 1e-6
    assert np.abs(v[:, 0] - tr["gvel"][0, -1]).max() < 1e-6


def test_human36_free_100_steps_free_running(torch_cuda):
    """configs[1]: 100 free-running steps on the 42-dof model vs the real reference."""
    model, tr = load_golden("human36_free")
    W = tr["gpos"].shape[0]
    bw = _batch(model, W)
    bw.set_state(tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy())
    bw.step(1e-3, 100)
    g, v, _ = bw.get_state()
    assert np.abs(g.T - tr["gpos"][:, -1]).max() < 1e-6
    assert rel(v.T, tr["gvel"][:, -1]) < 1e-6


def test_seeded_worlds_vs_oracle(torch_cuda):
    """configs[1]/[2] distribution: 4096 seeded worlds on the GPU, a sample of them
    stepped by the oracle from the same states (step 0: M, N, Z, Y; 3 steps: state)."""
    from arboris_b200 import scenarios
    from arboris_b200.flatten import flatten
    from oracle.arboris_oracle import OracleWorld
    for scen, W, sample in (("human36_free", 4096, (0, 1, 77, 4095)),
                            ("human36_contact", 1000, (0, 5, 999))):
        model = flatten(scenarios.BUILDERS[scen]())
        gpos, gvel = scenarios.initial_states(model, scen, 0, W)
        bw = _batch(model, W)
        bw.set_state(gpos, gvel)
        dt = 1e-3
        bw.update_dynamic()
        bw.update_controllers(dt)
        M, N, Z, Y = (bw.matrix(k).cpu().numpy() for k in ("mass", "nleffects", "impedance", "admittance"))
        bw.step(dt, 3)
        g, v, f = bw.get_state()
        for w in sample:
            o = OracleWorld(model.to_dict())
            o.gpos[:], o.gvel[:] = gpos[:, w], gvel[:, w]
            o.update_dynamic()
            o.update_controllers(dt)
            assert rel(M[w], o.mass) < REL_TOL and rel(N[w], o.nleffects) < REL_TOL
            assert rel(Z[w], o.impedance) < REL_TOL and rel(Y[w], o.admittance) < REL_TOL
            for _ in range(3):
                o.step(dt)
            assert rel(v[:, w], o.gvel) < REL_TOL*10 and rel(g[:, w], o.gpos) < REL_TOL
        # (ARB_STATUS_EIG_NOROOT = the reference's own silent clamp s = -1e10, constraints.py:827-830)
        assert int((bw.status() & ~4).max()) == 0


def test_zaligned_indices_bit_exact(torch_cuda):
    """integer index work: argsort(|normal|) of the contact plane (homogeneousmatrix.py:225)"""
    from oracle.arboris_oracle import OracleWorld
    model, tr = load_golden("human36_contact")
    bw = _batch(model, 1)
    bw.set_state(tr["gpos_in"][:1].T.copy(), tr["gvel_in"][:1].T.copy())
    bw.update_dynamic(); bw.update_controllers(1e-3); bw.update_constraints(1e-3)
    o = OracleWorld(model.to_dict())
    o.gpos[:], o.gvel[:] = tr["gpos_in"][0], tr["gvel_in"][0]
    o.update_dynamic(); o.update_controllers(1e-3); o.update_constraints(1e-3)
    z = bw.constraints("zidx").cpu().numpy()[0]
    for c in range(8):
        assert list(z[c]) == list(o.zidx[c])
    # and from the fused step's own scratch
    bw.set_state(tr["gpos_in"][:1].T.copy(), tr["gvel_in"][:1].T.copy())
    bw.step(1e-3, 1)
    z = bw.constraints("zidx").cpu().numpy()[0]
    for c in range(8):
        assert list(z[c]) == list(o.zidx[c])


def test_ragged_batch_sizes_and_errors(torch_cuda):
    """W = 1, W not a multiple of the block/warp size; ABI misuse returns errors."""
    from arboris_b200 import _capi
    model, tr = load_golden("human36_free")
    ref = None
    for W in (1, 33, 130):
        bw = _batch(model, W)
        bw.set_state(np.repeat(tr["gpos_in"][:1].T, W, 1), np.repeat(tr["gvel_in"][:1].T, W, 1))
        bw.step(1e-3, 2)
        g, v, _ = bw.get_state()
        assert np.abs(v - v[:, :1]).max() == 0.      # identical worlds give identical results
        if ref is None:
            ref = v[:, 0].copy()
        assert np.abs(v[:, 0] - ref).max() == 0.
    with pytest.raises(_capi.ArbError):
        bw.matrix("mass", 0, W + 1)
    with pytest.raises(AssertionError):
        bw.update_controllers(0.)


def test_cooperative_gauss_seidel_is_bit_identical_to_per_lane(torch_cuda):
    """The block-cooperative Gauss-Seidel kernel (sliding solves pooled through shared memory)
    does the same arithmetic in the same order inside every world as the per-lane kernel:
    states after 120 steps of falling humanoids (contacts in all three branches) are equal
    bit for bit, for a batch that is not a multiple of the block size."""
    from arboris_b200 import scenarios
    from arboris_b200.flatten import flatten
    model = flatten(scenarios.BUILDERS["human36_contact"]())
    W = 1000
    gpos, gvel = scenarios.initial_states(model, "human36_contact", 0, W)
    out = []
    for coop in (1, 0):
        bw = _batch(model, W)
        bw.set_option("gs_coop", coop)
        bw.set_state(gpos, gvel)
        bw.step(1e-3, 120)
        out.append(bw.get_state())
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    assert np.abs(out[0][2]).max() > 0        # contact forces are present


def test_staged_gs_bit_identical(torch_cuda):
    """The Gauss-Seidel kernel that stages A_cc and pinv(A_cc) of every contact visit in shared
    memory (TMA bulk copies, option gs_stage) does the same arithmetic in the same order as the
    kernel that loads them from global memory: states, forces, branches after 120 steps of falling
    humanoids are equal bit for bit, for a batch that is not a multiple of the warp size, with and
    without world sorting, in the plain and in the general instantiation of the staged kernel."""
    from arboris_b200 import scenarios
    from arboris_b200.flatten import flatten
    model = flatten(scenarios.BUILDERS["human36_contact"]())
    W = 1003
    gpos, gvel = scenarios.initial_states(model, "human36_contact", 0, W)
    out = []
    for stage, sort, plain in ((1, 2, 1), (0, 2, 1), (1, 0, 1), (1, 2, 0)):
        bw = _batch(model, W)
        bw.set_option("gs_stage", stage)
        bw.set_option("sort_period", sort)
        bw.set_option("gs_plain", plain)     # (0: the general instantiation of the staged kernel)
        bw.set_state(gpos, gvel)
        bw.step(1e-3, 120)
        out.append(bw.get_state() + (bw.constraints("branch").cpu().numpy(), bw.status().cpu().numpy()))
    for o in out[1:]:
        for a, b in zip(out[0], o):
            assert np.array_equal(a, b)
    assert np.abs(out[0][2]).max() > 0        # contact forces are present
    assert (out[0][3] == 3).any()             # some contact slides at the end


# ---------------------------------------------------------------------------------------------
# round 2: SURVEY.md 8(d) config 3 parity subset (64 worlds), free-running trajectories,
# per-world controller parameters, body Jacobians on the GPU
# ---------------------------------------------------------------------------------------------
def _load_extra(model_name, file_name):
    import os
    from conftest import GOLDEN
    from arboris_b200.flatten import FlatModel
    model = FlatModel.load(os.path.join(GOLDEN, "model_%s.npz" % model_name))
    with np.load(os.path.join(GOLDEN, file_name)) as z:
        return model, {k: z[k] for k in z.files}


@pytest.mark.parametrize("path", ["fused", "phases"])
def test_contact64_teacher_forced(torch_cuda, path):
    """64 falling humanoids: every step from the REAL reference's state;
    active sets and branches bit-exact, velocities / positions / forces within 1e-10."""
    model, tr = _load_extra("human36_contact", "traj_human36_contact64.npz")
    dt = float(tr["dt"])
    W, T = tr["gpos"].shape[:2]
    bw = _batch(model, W)
    if path == "phases":
        bw.set_option("force_phases", 1)
        T = min(T, 60)
    gpos, gvel = tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy()
    cf = np.zeros((model.nrows, W))
    worst, flips, nact = {}, 0, 0
    for s in range(T):
        bw.set_state(gpos, gvel, cf)
        bw.step(dt, 1)
        g, v, f = bw.get_state()
        a = bw.constraints("active").cpu().numpy()
        br = bw.constraints("branch").cpu().numpy()
        flips += int((a != tr["active"][:, s]).sum()) + int((br*a != tr["branch"][:, s]).sum())
        nact += int(a.sum())
        worst["gvel"] = max(worst.get("gvel", 0), rel(v.T, tr["gvel"][:, s]))
        worst["gpos"] = max(worst.get("gpos", 0), rel(g.T, tr["gpos"][:, s]))
        worst["cforce"] = max(worst.get("cforce", 0), rel(f.T, tr["cforce"][:, s]))
        gpos, gvel, cf = tr["gpos"][:, s].T.copy(), tr["gvel"][:, s].T.copy(), tr["cforce"][:, s].T.copy()
    assert nact > 1000, "the fixture must exercise contacts"
    assert flips == 0, "active set / branch flips: %d" % flips
    for k, v in worst.items():
        assert v <= REL_TOL, (k, v)


def test_contact64_free_running(torch_cuda):
    """The same 64 worlds FREE RUNNING on the device for the whole fixture (120 steps, from first
    touch-down into sliding contact): positions within 1e-6 of the real reference's own
    trajectory, and active sets / branches identical at every step, for every world that has not
    hit a branch tie (contact dynamics amplify the 1e-16 per-step differences; a world whose
    friction-cone test is decided at that level flips -- those worlds are counted and bounded)."""
    model, tr = _load_extra("human36_contact", "traj_human36_contact64.npz")
    dt = float(tr["dt"])
    W, T = tr["gpos"].shape[:2]
    bw = _batch(model, W)
    bw.set_state(tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy(), np.zeros((model.nrows, W)))
    first_flip = np.full(W, T)
    err = np.zeros((T, W))
    for s in range(T):
        bw.step(dt, 1)
        g, v, _ = bw.get_state()
        a = bw.constraints("active").cpu().numpy()
        br = bw.constraints("branch").cpu().numpy()
        bad = (a != tr["active"][:, s]).any(1) | ((br*a) != tr["branch"][:, s]).any(1)
        first_flip = np.where(bad & (first_flip == T), s, first_flip)
        err[s] = np.abs(g.T - tr["gpos"][:, s]).max(1)
    clean = first_flip == T
    print("free-running contact: %d of %d worlds without any flip over %d steps; first flips at %s; "
          "max |dq| of the clean worlds %.3g" % (clean.sum(), W, T, sorted(first_flip[~clean].tolist()),
                                                  err[:, clean].max()))
    assert clean.sum() >= int(0.9*W)
    assert err[:, clean].max() < 1e-6
    # before its first flip every world follows the reference
    for w in np.nonzero(~clean)[0]:
        assert err[:first_flip[w], w].max(initial=0.) < 1e-6


@pytest.mark.parametrize("name", ["human36_free", "snake_loop"])
def test_free_running_1000_steps(torch_cuda, name):
    """north_star: trajectories within 1e-6 after 1000 steps (configs 2 and 4), against
    checkpoints of 1000 free-running steps of the real reference.  The closed-loop snake is
    compared over all 1000 steps.  The uncontrolled free-floating humanoid is compared for as long
    as the REFERENCE's own trajectory exists: its semi-implicit Euler step is unstable on this
    model (max |gvel| of the reference: 2.6 at step 99, 19.7 at 299, 168 at 399, 1.9e6 at 499 --
    tests/golden/free_human36_free.npz), so a world is checked at every checkpoint where the
    reference's velocities are still below 200 rad/s (steps <= 399 for all four worlds, later for
    the tamer ones) and the test requires that this covers at least 400 steps of every world."""
    model, tr = _load_extra(name, "free_%s.npz" % name)
    W = tr["gpos"].shape[0]
    bw = _batch(model, W)
    bw.set_state(tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy(),
                 np.zeros((max(model.nrows, 1), W)))
    done = 0
    alive = np.ones(W, bool)
    covered = np.zeros(W, int)
    for i, k in enumerate(tr["kept_steps"]):
        bw.step(float(tr["dt"]), int(k) + 1 - done)
        done = int(k) + 1
        g, v, f = bw.get_state()
        alive &= np.abs(tr["gvel"][:, i]).max(1) < 200.
        if not alive.any():
            continue
        covered[alive] = done
        assert np.abs(g.T - tr["gpos"][:, i])[alive].max() < 1e-6, (name, int(k))
        vs = max(1., np.abs(tr["gvel"][:, i][alive]).max())
        assert np.abs(v.T - tr["gvel"][:, i])[alive].max() < 1e-6*vs, (name, int(k))
    assert done == 1000
    assert covered.min() >= (1000 if name == "snake_loop" else 400), covered
    if name == "snake_loop":
        assert int(bw.status().max()) == 0


@pytest.mark.parametrize("path", ["fused", "phases"])
def test_per_world_pd_parameters(torch_cuda, path):
    """every world its own kp, kd, gpos_des, gvel_des
    (arb_batch_bind_controller_params) against the real reference run once per world with those
    values on its ProportionalDerivativeController objects (controllers.py:113-159)."""
    model, tr = _load_extra("zoo", "traj_zoo_pd.npz")
    dt = float(tr["dt"])
    W, T = tr["gpos"].shape[:2]
    bw = _batch(model, W)
    assert len(bw.pd_dofs()) == tr["pd_kp"].shape[0]
    bw.set_controller_params(kp=tr["pd_kp"], kd=tr["pd_kd"], gpos_des=tr["pd_gpos_des"],
                             gvel_des=tr["pd_gvel_des"])
    gpos, gvel = tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy()
    cf = np.zeros((max(model.nrows, 1), W))
    fs = list(tr["full_steps"])
    worst = {}
    for s in range(T):
        bw.set_state(gpos, gvel, cf)
        if path == "phases":
            bw.update_dynamic()
            bw.update_controllers(dt)
            worst["gforce_ctrl"] = max(worst.get("gforce_ctrl", 0),
                                       rel(bw.gforce().cpu().numpy(), tr["gforce_ctrl"][:, s]))
            if s in fs:
                for k in ("impedance", "admittance"):
                    worst[k] = max(worst.get(k, 0), rel(bw.matrix(k).cpu().numpy(), tr[k][:, fs.index(s)]))
            bw.update_constraints(dt)
            bw.integrate(dt)
        else:
            bw.step(dt, 1)
        g, v, _ = bw.get_state()
        worst["gvel"] = max(worst.get("gvel", 0), rel(v.T, tr["gvel"][:, s]))
        worst["gpos"] = max(worst.get("gpos", 0), rel(g.T, tr["gpos"][:, s]))
        gpos, gvel = tr["gpos"][:, s].T.copy(), tr["gvel"][:, s].T.copy()
        if model.nrows:
            cf = tr["cforce"][:, s].T.copy()
    for k, v in worst.items():
        assert v <= REL_TOL, (k, v)
    # and the worlds really differ: with the model's parameters the result is another one
    bw.set_controller_params(kp=False, kd=False, gpos_des=False, gvel_des=False)
    bw.set_state(tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy(), np.zeros((max(model.nrows, 1), W)))
    bw.step(dt, 1)
    assert rel(bw.get_state()[1].T, tr["gvel"][:, 0]) > 1e-6


def test_body_jacobians_on_device(torch_cuda):
    """a3: Body.jacobian / djacobian / twist / nleffects read back from the DEVICE (arb_get_body)
    against the oracle (pinned to the real reference by tests/test_oracle.py) on the 42-dof model."""
    from oracle.arboris_oracle import OracleWorld
    model, tr = load_golden("human36_free")
    W = tr["gpos_in"].shape[0]
    bw = _batch(model, W)
    bw.set_state(tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy())
    bw.update_dynamic()
    for w in range(W):
        o = OracleWorld(model.to_dict())
        o.gpos[:], o.gvel[:] = tr["gpos_in"][w], tr["gvel_in"][w]
        o.update_dynamic()
        for b in range(1, model.nj + 1):
            assert rel(bw.body("jacobian", b, w, w + 1)[0].cpu().numpy(), o.jac[b]) < REL_TOL
            assert rel(bw.body("djacobian", b, w, w + 1)[0].cpu().numpy(), o.djac[b]) < REL_TOL
            assert rel(bw.body("twist", b, w, w + 1)[0].cpu().numpy(), o.twist[b]) < REL_TOL
            assert rel(bw.body("pose", b, w, w + 1)[0].cpu().numpy(), o.pose[b]) < REL_TOL
            nle = np.abs(o.body_nle[b]).max()
            if nle > 0:
                assert rel(bw.body("nleffects", b, w, w + 1)[0].cpu().numpy(), o.body_nle[b]) < REL_TOL


def test_getters_refuse_stale_phase_scratch(torch_cuda):
    """After a fused step the assembled matrices are not formed: the getters must say so instead
    of returning what an earlier phase call left behind."""
    from arboris_b200 import _capi
    model, tr = load_golden("human36_free")
    bw = _batch(model, 2)
    bw.set_state(tr["gpos_in"][:2].T.copy(), tr["gvel_in"][:2].T.copy())
    bw.update_dynamic()
    bw.update_controllers(1e-3)
    M0 = bw.matrix("mass").cpu().numpy()
    bw.step(1e-3, 1)
    for call in (lambda: bw.matrix("mass"), lambda: bw.gforce(), lambda: bw.body("jacobian", 1)):
        with pytest.raises(_capi.ArbError):
            call()
    bw.body("pose", 1)                       # the fused step does leave poses and twists
    bw.update_dynamic()
    assert rel(bw.matrix("mass").cpu().numpy(), M0) > 0     # fresh, and the world has moved


# ---------------------------------------------------------------------------------------------
# the prepare stage with 16 lanes per world and on-chip scratch (csrc/arb_group.cuh)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["simplearm", "human36_free", "ball_socket", "simplearm_limits",
                                  "snake_loop", "human36_contact", "balls", "zoo", "contact64"])
def test_group_prepare_vs_real_reference(torch_cuda, name):
    """arb_step with the "prepare_group" option (group prepare stage + K-matrix finish stage) from
    the real reference's state at every step: same bars as the lane-per-world stages."""
    if name == "contact64":
        model, tr = _load_extra("human36_contact", "traj_human36_contact64.npz")
    else:
        model, tr = load_golden(name)
    dt = float(tr["dt"])
    W, T = tr["gpos"].shape[:2]
    T = min(T, 300)
    bw = _batch(model, W)
    bw.set_option("prepare_group", 1)
    gpos, gvel = tr["gpos_in"].T.copy(), tr["gvel_in"].T.copy()
    cf = np.zeros((max(model.nrows, 1), W))
    worst, flips = {}, 0
    for s in range(T):
        bw.set_state(gpos, gvel, cf)
        bw.step(dt, 1)
        g, v, f = bw.get_state()
        if model.nc:
            a = bw.constraints("active").cpu().numpy()
            br = bw.constraints("branch").cpu().numpy()
            flips += int((a != tr["active"][:, s]).sum()) + int((br*a != tr["branch"][:, s]).sum())
        worst["gvel"] = max(worst.get("gvel", 0), rel(v.T, tr["gvel"][:, s]))
        worst["gpos"] = max(worst.get("gpos", 0), rel(g.T, tr["gpos"][:, s]))
        if model.nrows:
            worst["cforce"] = max(worst.get("cforce", 0), rel(f.T, tr["cforce"][:, s]))
        gpos, gvel = tr["gpos"][:, s].T.copy(), tr["gvel"][:, s].T.copy()
        if model.nrows:
            cf = tr["cforce"][:, s].T.copy()
    assert flips == 0, "active set / branch flips with the group prepare stage: %d" % flips
    assert int((bw.status() & ~4).max()) == 0
    for k, v in worst.items():
        assert v <= REL_TOL, (k, v)


def test_group_prepare_large_batch_matches_lane_stages(torch_cuda):
    """4096 seeded falling humanoids, 100 free-running steps, world sorting on: the group stages
    against the lane-per-world stages (two different summation orders: not bit-identical, but within
    1e-9 after 100 steps for every world whose active sets agree; flips are counted)."""
    from arboris_b200 import scenarios
    from arboris_b200.flatten import flatten
    model = flatten(scenarios.BUILDERS["human36_contact"]())
    W = 4096
    gpos, gvel = scenarios.initial_states(model, "human36_contact", 0, W)
    out = []
    for grp in (1, 0):
        bw = _batch(model, W)
        bw.set_option("prepare_group", grp)
        bw.set_state(gpos, gvel)
        bw.step(1e-3, 60)
        out.append((bw.get_state(), bw.constraints("active").cpu().numpy()))
    same = (out[0][1] == out[1][1]).all(1)
    dv = np.abs(out[0][0][1] - out[1][0][1]).max(0)/np.maximum(np.abs(out[1][0][1]).max(0), 1e-3)
    print("group vs lane stages after 60 free-running steps: %d of %d worlds with identical active sets, "
          "max rel dv among them %.3g" % (same.sum(), W, dv[same].max()))
    assert same.sum() >= 0.99*W
    assert dv[same].max() < 1e-7
    assert np.isfinite(out[0][0][1]).all()
