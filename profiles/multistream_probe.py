"""Probe: do the kernels of several batches on separate streams overlap?  (gpurun, one GPU)"""
import sys, time, numpy as np, torch
sys.path[:0]=['/root/repo','/root/repo/arboris-python_b200']
from arboris_b200 import scenarios
from arboris_b200.batch import BatchedWorld
from arboris_b200.flatten import flatten
model=flatten(scenarios.BUILDERS['human36_contact']())
gp,gv=scenarios.initial_states(model,'human36_contact',0,4096)
def run(K, Wtot=262144, steps=50, sort=2):
    W=Wtot//K
    parts=[BatchedWorld(model, W, device='cuda:0', stream=torch.cuda.Stream()) for _ in range(K)]
    g=np.tile(gp,(1,W//4096 if W>=4096 else 1))[:, :W]; v=np.tile(gv,(1,W//4096 if W>=4096 else 1))[:, :W]
    for p in parts: p.set_state(g,v); p.set_option('sort_period', sort)
    torch.cuda.synchronize()
    for p in parts: p.step(1e-3, 100)     # into contact
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for s in range(steps):
        for p in parts: p.step(1e-3, 1)
    torch.cuda.synchronize()
    dt=time.perf_counter()-t0
    print('sort_period=%d '%sort + 'K=%d streams x %d worlds: %.3g world-steps/s (%.2f ms/step)'%(K,W,Wtot*steps/dt,1e3*dt/steps), flush=True)
    for p in parts: p.close()
for sort in (0, 2):
    for K in (1, 4, 8):
        run(K, sort=sort)
