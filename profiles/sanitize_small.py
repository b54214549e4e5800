"""Small run for compute-sanitizer (memcheck / racecheck): lane stages with the staged (TMA + mbarrier)
and the unstaged Gauss-Seidel kernel, and the group prepare stage; 100 falling humanoids (a batch that
is not a multiple of the warp size), a few steps with contacts active, world sorting on."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "arboris-python_b200"))
import numpy as np, torch
from arboris_b200 import scenarios
from arboris_b200.batch import BatchedWorld
from arboris_b200.flatten import flatten
model = flatten(scenarios.BUILDERS["human36_contact"]())
W = 100
gp, gv = scenarios.initial_states(model, "human36_contact", 0, W)
for grp, stage in ((0, 1), (0, 0), (1, 1)):
    bw = BatchedWorld(model, W, device="cuda:0")
    bw.set_option("prepare_group", grp)
    bw.set_option("gs_stage", stage)
    bw.set_state(gp, gv)
    bw.step(1e-3, 6)
    torch.cuda.synchronize()
    print("group" if grp else "lane", "staged" if stage else "unstaged", "active constraints:",
          int(bw.constraints("active").sum()), "sliding:", int((bw.constraints("branch") == 3).sum()),
          "finite:", bool(torch.isfinite(bw.gvel).all()))
