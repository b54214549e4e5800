"""simulate() on a BatchedWorld with the batched observers: the fused
step split around Observer.update (arb_step_begin / arb_step_end) and the trajectory file in
the reference's Hdf5Logger layout, checked against the reference's own tests/simplearm_flat.h5."""
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _world(scen):
    from arboris_b200 import scenarios
    from arboris_b200.flatten import flatten
    return flatten(scenarios.BUILDERS[scen]())


def test_simulate_logger_reproduces_reference_h5(tmp_path):
    """reference recipe tests/test_visu_collada.py:11-27: simplearm, timeline arange(0, 1, .01),
    body poses before each integrate -> tests/simplearm_flat.h5 (tolerance 1e-10)."""
    from arboris_b200.batch import BatchedWorld
    from arboris_b200.core import simulate
    from arboris_b200.observers import BatchedHdf5Logger
    from oracle import h5lite
    ref = np.load(os.path.join(GOLDEN, "reference_h5.npz"))
    model = _world("simplearm")
    bw = BatchedWorld(model, 5, device="cuda:0")
    p1, p5 = str(tmp_path/"one.h5"), str(tmp_path/"all.h5")
    obs = [BatchedHdf5Logger(p1, worlds=[3], squeeze=True), BatchedHdf5Logger(p5, group="/run/a")]
    timeline = np.arange(0, 1, 0.01)
    simulate(bw, timeline, obs)
    one = h5lite.read(p1)
    assert sorted(one) == ["gpositions", "gvelocities", "timeline", "transforms"]
    assert np.abs(one["timeline"] - ref["simplearm_flat/timeline"]).max() < 1e-12
    for name in ("Arm", "Forearm", "Hand"):
        assert one["transforms"][name].shape == (99, 4, 4)
        assert np.abs(one["transforms"][name] - ref["simplearm_flat/transforms/" + name]).max() < 1e-10
    assert sorted(one["gpositions"]) == ["Elbow", "Shoulder", "Wrist"]
    assert one["gvelocities"]["Elbow"].shape == (99, 1)
    allw = h5lite.read(p5)["run"]["a"]
    assert allw["transforms"]["Hand"].shape == (99, 5, 4, 4)
    assert np.abs(allw["transforms"]["Hand"] - one["transforms"]["Hand"][:, None]).max() == 0.
    assert abs(bw.current_time - timeline[-1]) < 1e-12


def test_split_step_equals_fused_step_and_phases():
    """begin_step + end_step == step == the four phase calls, on falling humanoids in contact;
    between the halves the state is untouched and poses / active sets are those of the step."""
    import torch
    from arboris_b200 import scenarios
    from arboris_b200.batch import BatchedWorld
    model = _world("human36_contact")
    W = 200
    gp, gv = scenarios.initial_states(model, "human36_contact", 0, W)
    a, b, c = (BatchedWorld(model, W, device="cuda:0") for _ in range(3))
    for w in (a, b, c):
        w.set_state(gp, gv)
        w.step(1e-3, 80)                 # sorted assignment in force on all three
    a.step(1e-3, 1)
    g0 = b.gpos.clone()
    b.begin_step(1e-3)
    assert bool((b.gpos == g0).all())
    pose_b = b.body("pose", 4)
    act_b = b.constraints("active")
    b.end_step(1e-3)
    assert bool((a.gvel == b.gvel).all()) and bool((a.gpos == b.gpos).all())
    assert bool((a.constraints("active") == act_b).all())
    c.update_dynamic(); c.update_controllers(1e-3); c.update_constraints(1e-3)
    assert (c.body("pose", 4) - pose_b).abs().max().item() < 1e-12
    assert bool((c.constraints("active") == act_b).all())
    c.integrate(1e-3)
    assert (c.gvel - b.gvel).abs().max().item() < 1e-8*max(1., c.gvel.abs().max().item())
    b.step(1e-3, 5)                      # sorting resumes after a split step
    a.step(1e-3, 5)
    assert bool((a.gvel == b.gvel).all())


def test_energy_monitor_matches_mass_matrix():
    """BatchedEnergyMonitor: sum_b T_b^T M_b T_b / 2 equals gvel^T M gvel / 2 of the assembled
    mass matrix (observers.py:42-44) and, without gravity work... free flight conserves
    E_c + E_p to O(dt)."""
    import torch
    from arboris_b200 import scenarios
    from arboris_b200.batch import BatchedWorld
    from arboris_b200.core import simulate
    from arboris_b200.observers import BatchedEnergyMonitor
    model = _world("human36_free")
    W = 64
    gp, gv = scenarios.initial_states(model, "human36_free", 0, W)
    bw = BatchedWorld(model, W, device="cuda:0")
    bw.set_state(gp, gv)
    bw.update_dynamic()
    M = bw.matrix("mass")
    v = bw.gvel.T.contiguous()
    ec_ref = 0.5*torch.einsum("wi,wij,wj->w", v, M, v).cpu().numpy()
    mon = BatchedEnergyMonitor()
    simulate(bw, np.arange(0, 0.0505, 1e-3), [mon])
    assert mon.kinetic_energy.shape == (50, W)
    assert np.abs(mon.kinetic_energy[0] - ec_ref).max() < 1e-10*np.abs(ec_ref).max()
    e = mon.mechanichal_energy
    assert np.abs(e[-1] - e[0]).max() < 2e-2*np.abs(e[0]).max()      # semi-implicit Euler, dt = 1 ms


@pytest.mark.parametrize("mode", ["streams", "serial", "serial-queued"])
def test_host_pipeline_equals_device_step(mode):
    """HostPipeline (column blocks of pinned host state; one stream each with arb_step_host_strided,
    or all kernels on one stream and the copies of arb_state_copy_host_strided on two others)
    gives bit for bit the states of the device-resident step, for a batch that does not split evenly."""
    import torch
    from arboris_b200 import scenarios
    from arboris_b200.batch import BatchedWorld, HostPipeline
    model = _world("human36_contact")
    W = 1000
    gp, gv = scenarios.initial_states(model, "human36_contact", 0, W)
    bw = BatchedWorld(model, W, device="cuda:0")
    bw.set_state(gp, gv)
    hg = torch.as_tensor(gp).pin_memory()
    hv = torch.as_tensor(gv).pin_memory()
    hf = torch.zeros((model.nrows, W), dtype=torch.float64).pin_memory()
    pipe = HostPipeline(model, W, chunks=7, device="cuda:0", mode=mode.split("-")[0])
    for _ in range(90):
        # ("serial-queued": calls queued without waiting, HostPipeline.wait at the end)
        pipe.step(hg.numpy(), hv.numpy(), hf.numpy(), 1e-3, 1, sync=not mode.endswith("queued"))
    pipe.wait()
    bw.step(1e-3, 90)
    g, v, f = bw.get_state()
    assert np.array_equal(g, hg.numpy()) and np.array_equal(v, hv.numpy()) and np.array_equal(f, hf.numpy())
    assert np.abs(f).max() > 0
