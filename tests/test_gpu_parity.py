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
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
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
        if s in fs:
            i = fs.index(s)
            for k in ("mass", "nleffects", "impedance", "admittance"):
                worst[k] = max(worst.get(k, 0), rel(bw.matrix(k).cpu().numpy(), tr[k][:, i]))
        bw.update_constraints(dt)
        if model.nc:
            a = bw.constraints("active").cpu().numpy()
            br = bw.constraints("branch").cpu().numpy()
            flips += int((a != tr["active"][:, s]).sum()) + int((br*a != tr["branch"][:, s]).sum())
            if name != "ball_socket":
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
        if model.nrows and name != "ball_socket":
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
        assert int(bw.status().max()) == 0


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
