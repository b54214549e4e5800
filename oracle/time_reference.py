"""TEST INFRASTRUCTURE ONLY -- one-off timing of the REAL reference (imported in memory through
oracle/ref_loader.py, build container only) next to the numpy oracle port that bench.py's CPU arm
runs on the GPU box (where the reference tree does not exist).  Same scenario, same seeded world,
one episode (250 steps at dt = 1 ms) on ONE core, BLAS threads = 1.

    python oracle/time_reference.py            # prints world-steps/s of both
"""
import os
import sys
import time

for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[v] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "arboris-python_b200"))

import numpy as np  # noqa: E402

from oracle import ref_loader  # noqa: E402

ref_loader.install()
from arboris_b200 import flatten, scenarios  # noqa: E402
from oracle.arboris_oracle import OracleWorld  # noqa: E402
from oracle.make_goldens import set_state  # noqa: E402

DT, EPISODE = 1e-3, 250


def time_real(scen, wid):
    w = scenarios.BUILDERS[scen](reference=True)
    model = flatten(w)
    gpos, gvel = scenarios.initial_state(model, scen, wid)
    set_state(w, model, gpos, gvel)
    for c in w._constraints:
        c._force[:] = 0.
    t0 = time.perf_counter()
    for _ in range(EPISODE):
        w.update_dynamic()
        w.update_controllers(DT)
        w.update_constraints(DT)
        w.integrate(DT)
    return EPISODE/(time.perf_counter() - t0)


def time_port(scen, wid):
    model = flatten(scenarios.BUILDERS[scen]())
    o = OracleWorld(model.to_dict())
    o.gpos[:], o.gvel[:] = scenarios.initial_state(model, scen, wid)
    t0 = time.perf_counter()
    for _ in range(EPISODE):
        o.step(DT)
    return EPISODE/(time.perf_counter() - t0)


if __name__ == "__main__":
    for scen in ("human36_contact", "human36_free"):
        real = [time_real(scen, w) for w in (0, 1, 2)]
        port = [time_port(scen, w) for w in (0, 1, 2)]
        print("%-16s real reference %.1f world-steps/s/core, oracle port %.1f (ratio port/real %.2f)"
              % (scen, np.mean(real), np.mean(port), np.mean(port)/np.mean(real)))
