"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` from the REAL reference.

Run in the build container (needs ``/root/reference``):

    python oracle/make_goldens.py

It imports the unmodified reference through ``oracle/ref_loader.py`` (in-memory
py3 patches), builds the BASELINE.json scenarios with the
reference's own classes, steps them with the reference's own ``World`` methods
and records, per step, what the parity tests compare: M, N, B, Z, Y, gforce,
generalized velocities/positions, constraint active sets, solver branch ids,
forces and signed distances.  It also converts the reference's HDF5 goldens
(``tests/simplearm_flat.h5``, ``tests/human36.h5``) to ``.npz`` and stores the
flattened models, so that nothing on the GPU box needs the reference tree.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "arboris-python_b200"))

from oracle import ref_loader, h5lite  # noqa: E402

ref_loader.install()

import arboris.constraints as rcons  # noqa: E402
from arboris_b200 import flatten  # noqa: E402
from arboris_b200 import scenarios  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
BR_SEP, BR_STATIC, BR_SLIDING = 1, 2, 3

# ---- instrumentation of the reference (observation only) ----------------------
_branch_log = {}
_eig_called = [False]
_orig_eigvals = rcons.eigvals
_orig_sf_solve = rcons.SoftFingerContact.solve
_orig_jl_solve = rcons.JointLimits.solve


def _eigvals(B):
    _eig_called[0] = True
    return _orig_eigvals(B)


def _sf_solve(self, vel, admittance, dt):
    # same expression as constraints.py:780-781, evaluated before the call
    sep = self._sdist + dt*(vel - np.dot(admittance, self._force))[3] > 0
    _eig_called[0] = False
    out = _orig_sf_solve(self, vel, admittance, dt)
    _branch_log[id(self)] = BR_SEP if sep else (BR_SLIDING if _eig_called[0] else BR_STATIC)
    return out


def _jl_solve(self, vel, admittance, dt):
    pred = self._pos0 + dt*(vel - np.dot(admittance, self._force))
    _branch_log[id(self)] = 2 if (pred <= self._min) else (3 if (self._max <= pred) else 1)
    return _orig_jl_solve(self, vel, admittance, dt)


rcons.eigvals = _eigvals
rcons.SoftFingerContact.solve = _sf_solve
rcons.JointLimits.solve = _jl_solve


def set_state(world, model, gpos, gvel):
    for k, j in enumerate(world.iterjoints()):
        g = int(model.joint_gpos[k])
        if int(model.joint_type[k]) == 0:
            j.gpos = gpos[g:g + 16].reshape(4, 4).copy()
        else:
            j.gpos[:] = gpos[g:g + j.ndof]
    world._gvel[:] = gvel


def get_gpos(world, model):
    out = np.zeros(model.ngpos)
    for k, j in enumerate(world.iterjoints()):
        g = int(model.joint_gpos[k])
        v = np.asarray(j.gpos, dtype=float).reshape(-1)
        out[g:g + v.size] = v
    return out


def run_reference(scenario, world_ids, nsteps, dt, full_steps=(0,), reset_bs=True, keep=None,
                  fields=None, per_world=None):
    """Step the real reference; return dict of stacked per-step records.
    ``keep``: steps whose records are stored (default all) -- long free-running trajectories keep a
    few checkpoints only (``kept_steps`` in the output).  ``fields``: subset of the per-step
    records to store.  ``per_world(world, wid)``: hook run before a world is stepped (sets
    per-world controller parameters on the reference's own objects)."""
    w = scenarios.BUILDERS[scenario](reference=True)
    model = flatten(w)
    names = ("gpos", "gvel", "gforce_ctrl", "gforce", "active", "branch", "cforce", "sdist")
    rec = {k: [] for k in names if fields is None or k in fields}
    keepset = None if keep is None else set(int(k) for k in keep)
    full = {k: [] for k in ("mass", "nleffects", "viscosity", "impedance",
                            "admittance")}
    gpos_in, gvel_in = [], []
    cons = list(w._constraints)
    for wid in world_ids:
        gpos, gvel = scenarios.initial_state(model, scenario, wid)
        set_state(w, model, gpos, gvel)
        for c in cons:
            c._force[:] = 0.
        if per_world is not None:
            per_world(w, wid)
        gpos_in.append(gpos)
        gvel_in.append(gvel)
        r = {k: [] for k in rec}
        fl = {k: [] for k in full}
        for s in range(nsteps):
            _branch_log.clear()
            kept = keepset is None or s in keepset
            w.update_dynamic()
            w.update_controllers(dt)
            if kept and "gforce_ctrl" in r:
                r["gforce_ctrl"].append(w._gforce.copy())
            if s in full_steps:
                fl["mass"].append(w.mass.copy())
                fl["nleffects"].append(w.nleffects.copy())
                fl["viscosity"].append(w.viscosity.copy())
                fl["impedance"].append(w._impedance.copy())
                fl["admittance"].append(w._admittance.copy())
            w.update_constraints(dt)
            if kept:
                cf = np.zeros(model.nrows)
                sd = np.zeros(len(cons))
                for k, c in enumerate(cons):
                    r0 = int(model.cons_row[k])
                    f = np.asarray(c._force, dtype=float).reshape(-1)
                    cf[r0:r0 + f.size] = f
                    sd[k] = getattr(c, "_sdist", 0.) or 0.
                vals = {"gforce": w._gforce.copy(),
                        "active": [1 if (c.is_enabled() and c.is_active()) else 0 for c in cons],
                        "branch": [_branch_log.get(id(c), 0) for c in cons], "cforce": cf, "sdist": sd}
                for k2, v2 in vals.items():
                    if k2 in r:
                        r[k2].append(v2)
            w.integrate(dt)
            if kept:
                if "gvel" in r:
                    r["gvel"].append(w._gvel.copy())
                if "gpos" in r:
                    r["gpos"].append(get_gpos(w, model))
        for k in rec:
            rec[k].append(np.array(r[k]))
        for k in full:
            full[k].append(np.array(fl[k]))
    out = {k: np.array(v) for k, v in rec.items()}           # (W, T, ...)
    out.update({k: np.array(v) for k, v in full.items()})      # (W, len(full_steps), n, n)
    out["gpos_in"] = np.array(gpos_in)
    out["gvel_in"] = np.array(gvel_in)
    out["world_ids"] = np.array(world_ids)
    out["full_steps"] = np.array(full_steps)
    out["dt"] = np.array(dt)
    out["kept_steps"] = np.array(sorted(keepset) if keepset is not None else np.arange(nsteps))
    for k in ("active", "branch"):
        if k in out:
            out[k] = out[k].astype(np.int8)
    return model, out


def save(name, model, out, prefix="traj", with_model=True):
    if with_model:
        model.save(os.path.join(GOLD, "model_%s.npz" % name))
    np.savez_compressed(os.path.join(GOLD, "%s_%s.npz" % (prefix, name)), **out)
    print("wrote", name, {k: v.shape for k, v in out.items() if v.ndim > 1})


def simplearm_h5_recipe():
    """tests/test_visu_collada.py:11-27 -> tests/simplearm_flat.h5: body poses
    logged before each integrate, 99 steps at dt = 1e-2."""
    w = scenarios.simplearm_world(reference=True)
    import arboris.core
    timeline = np.arange(0, 1, 0.01)
    bodies = w.getbodies()
    poses = {k: [] for k in ("Arm", "Forearm", "Hand")}

    class Log(arboris.core.Observer):
        def init(self, world, timeline):
            pass

        def update(self, dt):
            for k in poses:
                poses[k].append(bodies[k].pose.copy())

        def finish(self):
            pass
    arboris.core.simulate(w, timeline, [Log()])
    return {k: np.array(v) for k, v in poses.items()}


def main():
    os.makedirs(GOLD, exist_ok=True)
    # 1. the reference's own HDF5 goldens, converted
    ref = {}
    for f in ("simplearm_flat", "simplearm_notflat", "human36"):
        t = h5lite.flat(h5lite.read(os.path.join(ref_loader.REFERENCE_ROOT, "tests", f + ".h5")))
        for k, v in t.items():
            ref[f + "/" + k] = v
    np.savez_compressed(os.path.join(GOLD, "reference_h5.npz"), **ref)
    mine = simplearm_h5_recipe()
    err = max(np.abs(mine[k] - ref["simplearm_flat/transforms/" + k]).max() for k in mine)
    print("real reference vs its own simplearm_flat.h5: max abs err %.3g" % err)
    assert err < 1e-12

    # 2. scenario trajectories from the real reference
    m, o = run_reference("simplearm", [0], 1000, 1e-3, full_steps=(0, 1, 500, 999))
    save("simplearm", m, o)
    m, o = run_reference("human36_free", [0, 1, 2, 3], 100, 1e-3, full_steps=(0, 1, 99))
    save("human36_free", m, o)
    m, o = run_reference("human36_contact", [0, 1, 2, 3], 300, 1e-3, full_steps=(0, 150))
    save("human36_contact", m, o)
    print("  contact branches seen:", np.unique(o["branch"], return_counts=True))
    m, o = run_reference("snake_loop", [0, 1], 200, 1e-3, full_steps=(0, 1))
    save("snake_loop", m, o)
    m, o = run_reference("ball_socket", [0, 1], 20, 1e-3, full_steps=(0,))
    save("ball_socket", m, o)
    m, o = run_reference("simplearm_limits", [0, 1], 300, 1e-3, full_steps=(0,))
    save("simplearm_limits", m, o)
    print("  limits branches seen:", np.unique(o["branch"], return_counts=True))
    balls()
    zoo()
    contact64()
    free_running()
    zoo_pd()


STATE_FIELDS = ("gpos", "gvel", "cforce", "active", "branch")


def contact64():
    """SURVEY.md 8(d) config 3 parity subset: 64 falling humanoids, the reference's own
    free-running trajectory (state, forces, active sets, branches after every step)."""
    m, o = run_reference("human36_contact", list(range(64)), 120, 1e-3, full_steps=(), fields=STATE_FIELDS)
    save("human36_contact64", m, o, with_model=False)
    print("  contact64 branches:", np.unique(o["branch"], return_counts=True))


def free_running():
    """north_star: trajectories within 1e-6 after 1000 steps.  1000 free-running steps of the
    real reference on configs 2 and 4, checkpoints every 100 steps."""
    keep = list(range(99, 1000, 100))
    m, o = run_reference("human36_free", [0, 1, 2, 3], 1000, 1e-3, full_steps=(), keep=keep,
                         fields=("gpos", "gvel"))
    save("human36_free", m, o, prefix="free", with_model=False)
    m, o = run_reference("snake_loop", [0, 1], 1000, 1e-3, full_steps=(), keep=keep,
                         fields=("gpos", "gvel", "cforce"))
    save("snake_loop", m, o, prefix="free", with_model=False)


def pd_params(wid, npd):
    """Per-world PD parameters of the zoo_pd fixture (rows as arb_model_pd_dofs)."""
    rng = np.random.default_rng(777 + wid)
    return {"kp": rng.uniform(5., 40., npd), "kd": rng.uniform(.3, 3., npd),
            "gpos_des": rng.uniform(-.3, .3, npd), "gvel_des": rng.uniform(-.2, .2, npd)}


def zoo_pd():
    """per-world controller parameters.  Every world of the zoo gets its own
    kp, kd (diagonal), gpos_des, gvel_des on the reference's ProportionalDerivativeController
    objects (controllers.py:113-131) before it is stepped."""
    import arboris.controllers as rctrl
    stash = {}

    def hook(world, wid):
        pds = [c for c in world._controllers if isinstance(c, rctrl.ProportionalDerivativeController)]
        npd = sum(c._cndof for c in pds)
        p = pd_params(wid, npd)
        stash[wid] = p
        r = 0
        for c in pds:
            n = c._cndof
            c.kp = np.diag(p["kp"][r:r + n])
            c.kd = np.diag(p["kd"][r:r + n])
            c.gpos_des = p["gpos_des"][r:r + n].copy()
            c.gvel_des = p["gvel_des"][r:r + n].copy()
            r += n
    ids = [0, 1, 2, 3]
    m, o = run_reference("zoo", ids, 200, 1e-3, full_steps=(0, 100), per_world=hook)
    for k in ("kp", "kd", "gpos_des", "gvel_des"):
        o["pd_" + k] = np.array([stash[w][k] for w in ids]).T        # (npd, W)
    save("zoo_pd", m, o, with_model=False)
    print("  zoo_pd: limit active in", int(o["active"].sum()), "world-steps")


def zoo():
    """every stock joint type, SubFrame joint frames, viscosity, two
    robots in one world, PD controller (diagonal gains), joint limits."""
    m, o = run_reference("zoo", [0, 1, 2], 300, 1e-3, full_steps=(0, 1, 150))
    save("zoo", m, o)
    print("  zoo: limit active in", int(o["active"].sum()), "world-steps; branches",
          np.unique(o["branch"], return_counts=True))


def balls():
    """plane/sphere, box/sphere, sphere/sphere, sphere/point contacts."""
    m, o = run_reference("balls", [0, 1, 2], 400, 1e-3, full_steps=(0,))
    save("balls", m, o)
    print("  balls: active per constraint:", o["active"].sum((0, 1)).tolist(),
          "branches:", np.unique(o["branch"], return_counts=True))


if __name__ == "__main__":
    only = {"balls": balls, "zoo": zoo, "contact64": contact64, "free_running": free_running,
            "zoo_pd": zoo_pd}
    if len(sys.argv) > 1 and all(a in only for a in sys.argv[1:]):
        for a in sys.argv[1:]:
            only[a]()
    else:
        main()
