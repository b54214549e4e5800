"""Host measurement (CPU build of the device routines, tests/hosttest): how the sliding-friction solves of
falling humanoids leave poly6_largest_root_fast.  usage: python profiles/fastroot_counters.py [W] [steps]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "arboris-python_b200"), os.path.join(ROOT, "tests", "hosttest")):
    sys.path.insert(0, p)
import ctypes as C
import numpy as np
import harness
from arboris_b200 import scenarios
from arboris_b200.flatten import flatten

W = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
model = flatten(scenarios.BUILDERS["human36_contact"]())
gp, gv = scenarios.initial_states(model, "human36_contact", 0, W)
hb = harness.HostBatch(model, W)
hb.set_coop(0)
hb.gpos[:], hb.gvel[:] = gp, gv
L = harness.lib()
L.ht_fastroot_hits.restype = C.c_long
L.ht_fastroot_fail.restype = C.c_long
L.ht_fastroot_fail.argtypes = [C.c_int]
t0 = time.time()
for s in range(T):
    hb.fused_step(1e-3)
hits = L.ht_fastroot_hits()
f = [L.ht_fastroot_fail(i) for i in range(8)]
tot = hits + sum(f[:5])
print("%d worlds x %d steps in %.1f s: %d sliding solves; certified by the fast path %d (%.3f %%)" %
      (W, T, time.time() - t0, tot, hits, 100.*hits/max(tot, 1)))
print("left the fast path: variance < 0 / NaN %d, left of a root %d, discriminant / slope %d, "
      "no convergence %d, certificate %d  (%.3f %% in all)" % (f[0], f[1], f[2], f[3], f[4], 100.*sum(f[:5])/max(tot, 1)))
print("recoveries: bracket refinements inside the fast path %d; of those that left it, no root t >= 0 proven by the "
      "positivity march %d, rigorous isolation %d (%.4f %% of the solves); Laguerre iterations per solve %.2f" %
      (f[5], f[6], sum(f[:5]) - f[6], 100.*(sum(f[:5]) - f[6])/max(tot, 1), f[7]/max(tot, 1)))
print("finite:", bool(np.isfinite(hb.gvel).all()), " checksum gvel %.17g" % float(np.abs(hb.gvel).sum()))
