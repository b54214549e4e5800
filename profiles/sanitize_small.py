"""Small run for compute-sanitizer (memcheck / racecheck): lane stages and the group prepare stage,
64 falling humanoids, a few steps with contacts active, world sorting on."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "arboris-python_b200"))
import numpy as np, torch
from arboris_b200 import scenarios
from arboris_b200.batch import BatchedWorld
from arboris_b200.flatten import flatten
model = flatten(scenarios.BUILDERS["human36_contact"]())
W = 96
gp, gv = scenarios.initial_states(model, "human36_contact", 0, W)
for grp in (0, 1):
    bw = BatchedWorld(model, W, device="cuda:0")
    bw.set_option("prepare_group", grp)
    bw.set_state(gp, gv)
    bw.step(1e-3, 4)
    torch.cuda.synchronize()
    print("group" if grp else "lane", "active constraints:", int(bw.constraints("active").sum()), "finite:", bool(torch.isfinite(bw.gvel).all()))
