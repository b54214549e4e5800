"""Probe: where does the end-to-end step (HostPipeline.step: pinned host state -> device, one step,
device -> host, every step) lose time against the device-resident step?  (gpurun, one GPU)

For K column blocks on K streams, per step and with a host synchronisation of all blocks every step:
  full    = HostPipeline.step                       (what bench.py reports as e2e)
  copies  = the same call with nsteps = 0           (H2D + D2H only)
  compute = BatchedWorld.step on every block        (kernels only, state resident)
"""
import sys, time, numpy as np, torch
sys.path[:0] = ['/root/repo', '/root/repo/arboris-python_b200']
from arboris_b200 import scenarios
from arboris_b200.batch import HostPipeline
from arboris_b200.flatten import flatten

DT = 1e-3
model = flatten(scenarios.BUILDERS['human36_contact']())
W = 262144
gp, gv = scenarios.initial_states(model, 'human36_contact', 0, 4096)
hg = torch.as_tensor(np.tile(gp, (1, W//4096))).pin_memory()
hv = torch.as_tensor(np.tile(gv, (1, W//4096))).pin_memory()
hf = torch.zeros((max(int(model.nrows), 1), W), dtype=torch.float64).pin_memory()
g0, v0 = hg.clone(), hv.clone()


def timeit(fn, steps=30):
    fn(); fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return 1e3*(time.perf_counter() - t0)/steps


import os
SORT = int(os.environ.get("PROBE_SORT_PERIOD", "-1"))
MODE = os.environ.get("PROBE_MODE", "streams")
NCS = int(os.environ.get("PROBE_COMPUTE_STREAMS", "1"))


def run(chunks):
    hg.copy_(g0); hv.copy_(v0); hf.zero_()
    pipe = HostPipeline(model, W, chunks=chunks, mode=MODE, compute_streams=NCS)
    if SORT >= 0:
        pipe.set_option("sort_period", SORT)
    hgn, hvn, hfn = hg.numpy(), hv.numpy(), hf.numpy()
    for _ in range(120):                      # into contact
        pipe.step(hgn, hvn, hfn, DT, 1)
    full = timeit(lambda: pipe.step(hgn, hvn, hfn, DT, 1))
    copies = timeit(lambda: pipe.step(hgn, hvn, hfn, DT, 0)) if MODE == "streams" else float("nan")

    def compute():
        for p in pipe.parts:
            p.step(DT, 1)
        for p in pipe.parts:
            p.synchronize()
    comp = timeit(compute)
    print('mode=%s/%d sort_period=%d ' % (MODE, NCS, SORT) + 'chunks=%s: full %.2f ms (%.3g w-s/s)  copies only %.2f ms  compute only %.2f ms'
          % (chunks, full, W/full*1e3, copies, comp), flush=True)
    pipe.close()


for ch in [int(x) if ',' not in x else [int(y) for y in x.split(',')] for x in sys.argv[1:]] or [1, 2, 4, 8, 16]:
    run(ch)
