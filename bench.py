#!/usr/bin/env python
"""Benchmark of the batched Arboris step (BASELINE.json metric: world-steps/s,
human36, fp64, dt = 1 ms).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port)

A "step" is one pass of the hot path (update_dynamic -> update_controllers ->
update_constraints -> integrate) over the whole batch of worlds.  Prints ONE JSON
line (rank 0).  Under torchrun (N > 1) the worlds are sharded over the ranks with
no data-path collective; NCCL only carries the timing/diagnostics reduction.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "arboris-python_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

DT = 1e-3
# algorithmic flops per human36 world-step (1 FMA = 2 flop)
FLOP_PER_WORLD_STEP = {"human36_contact": 0.99e6, "human36_free": 360118.}
STATE_BYTES_PER_WORLD_STEP = 1504   # read+write of 94 doubles
WORKLOADS = {
    # BASELINE.json configs[4] (the configuration the metric's "1/2/4/8 B200" sweep is quoted on)
    "human36_contact_262144": ("human36_contact", 262144),
    # BASELINE.json configs[1]
    "human36_free_4096": ("human36_free", 4096),
    "human36_contact_16384": ("human36_contact", 16384),
}
EPISODE = 250      # steps after which a world is re-initialised (the uncontrolled humanoid
                   # collapses and the reference's own sliding solve diverges after ~0.35 s)
GROUPS = 10        # the batch is split into GROUPS contiguous blocks whose episodes are staggered
PHASE = EPISODE//GROUPS   # by PHASE steps, so that at ANY step the batch holds every age of the
                   # episode (free fall, first contacts, 8 contacts with ~20 % sliding) in equal
                   # shares: the measured rate does not depend on which steps are timed


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="human36_contact_262144", choices=sorted(WORKLOADS))
    ap.add_argument("--worlds", type=int, default=0, help="override the number of worlds (per GPU "
                    "with --scaling weak, in total with --scaling strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): every GPU steps the workload's number of worlds (worlds are "
                         "independent units: per-GPU work fixed, no collective); strong: the workload's "
                         "worlds are split over the GPUs (BASELINE.json configs[4] read literally)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.)
    ap.add_argument("--e2e-chunks", type=lambda t: [int(x) for x in t.split(",")] if "," in t else int(t), default=[1, 1, 2, 2, 2, 1, 1],
                    help="column blocks of the end-to-end pass: a number of equal blocks or their relative "
                         "sizes (a,b,c,...); 1 = one synchronous arb_step_host call per step")
    ap.add_argument("--e2e-mode", default="serial", choices=["serial", "streams"],
                    help="HostPipeline mode: 'serial' = the kernels of all blocks on --e2e-compute-streams "
                         "streams, block after block, copies on two more streams; 'streams' = one stream per block")
    ap.add_argument("--e2e-compute-streams", type=int, default=3)
    ap.add_argument("--opt", action="append", help="arb_batch_set_option switch, name=value (A/B runs)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference on the host cores (multiprocessing)
# ---------------------------------------------------------------------------------------
def _cpu_worker(args):
    scen, wid, seconds = args
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    from arboris_b200 import scenarios
    from arboris_b200.flatten import flatten
    from oracle.arboris_oracle import OracleWorld
    model = flatten(scenarios.BUILDERS[scen]())
    o = OracleWorld(model.to_dict())
    o.gpos[:], o.gvel[:] = scenarios.initial_state(model, scen, wid)
    o.step(DT)                                   # warm-up (imports, caches)
    o.gpos[:], o.gvel[:] = scenarios.initial_state(model, scen, wid)
    o.cforce[:] = 0.
    # one whole episode from the reset state = the age mix the GPU batch holds at any step
    n, t0 = 0, time.perf_counter()
    while n < EPISODE and time.perf_counter() - t0 < seconds:
        o.step(DT)
        n += 1
    return n, time.perf_counter() - t0


def cpu_baseline(scen, seconds):
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(scen, w, seconds) for w in range(cores)])
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    nmin = min(r[0] for r in res)
    return {"value": steps/wall, "unit": "world-steps/s", "cores": cores, "kind": "port",
            "sample": "%d worlds (one per core), each one episode of %s from its reset state "
                      "(%d of %d steps done within the %.0f s cap), numpy oracle port of the "
                      "reference, BLAS threads = 1" % (cores, scen, nmin, EPISODE, seconds),
            "per_core": steps/wall/cores}


# ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows)
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm)//2], "sm_max_mhz": float(self.rows[0][2]),
                "power_w_max": max(float(r[3]) for r in self.rows), "samples": len(self.rows),
                "reasons": sorted(reasons)}


class Episodes(object):
    """Staggered episodes over GROUPS contiguous blocks of the batch (see EPISODE).
    ``state`` is a triple of (elem, W) tensors (device tensors of the BatchedWorld, or
    pinned host tensors for the end-to-end pass), ``init`` the matching reset values."""

    def __init__(self, W, state, init, step_fn):
        from arboris_b200.shard import shard_range
        self.blocks = [shard_range(W, g, GROUPS) for g in range(GROUPS)]
        self.state, self.init, self.step_fn = state, init, step_fn
        self.t = -(EPISODE - PHASE)
        for g in range(GROUPS):
            self.reset(g)

    def reset(self, g):
        w0, w1 = self.blocks[g]
        if w1 > w0:
            for dst, src in zip(self.state, self.init):
                if src is None:
                    dst[:, w0:w1].zero_()
                else:
                    dst[:, w0:w1].copy_(src[:, w0:w1])

    def advance(self, k):
        while k > 0:
            if self.t % PHASE == 0:
                for g in range(GROUPS):
                    if (self.t + PHASE*g) % EPISODE == 0:
                        self.reset(g)
            c = min(k, PHASE - self.t % PHASE)
            self.step_fn(c)
            self.t += c
            k -= c

    def prime(self):
        """Untimed: bring group g to age PHASE*g."""
        self.advance(-self.t)


def run_ours(a):
    import torch
    import torch.distributed as dist
    from arboris_b200 import scenarios, _capi
    from arboris_b200.batch import BatchedWorld, HostPipeline
    from arboris_b200.flatten import flatten
    from arboris_b200.shard import env_rank, shard_range, reduce_report

    rank, world_size, local = env_rank()
    torch.cuda.set_device(local)
    # stdout carries exactly ONE line (the JSON, rank 0): anything a library prints there
    # meanwhile (NCCL's version banner at the first collective) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scen, total = WORKLOADS[a.workload]
    if a.worlds:
        total = a.worlds
    if a.scaling == "weak":
        total *= world_size
    w0, w1 = shard_range(total, rank, world_size)   # contiguous block of worlds per GPU
    W = w1 - w0
    model = flatten(scenarios.BUILDERS[scen]())
    # synthetic seeded initial states; 4096 distinct worlds tiled over the shard
    nseed = min(W, 4096)
    gp, gv = scenarios.initial_states(model, scen, w0 % 4096, w0 % 4096 + nseed)
    reps = (W + nseed - 1)//nseed
    gp = np.tile(gp, (1, reps))[:, :W]
    gv = np.tile(gv, (1, reps))[:, :W]
    bw = BatchedWorld(model, W, device="cuda:%d" % local)
    for opt in (a.opt or []):                      # kernel A/B switches: --opt gs_coop=1
        name, val = opt.split("=")
        bw.set_option(name, int(val))
    gpos0 = torch.as_tensor(gp, device=bw.device)
    gvel0 = torch.as_tensor(gv, device=bw.device)
    ep = Episodes(W, (bw.gpos, bw.gvel, bw.cforce), (gpos0, gvel0, None),
                  lambda c: bw.step(DT, c))

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ep.prime()                      # untimed: stagger the episodes
    warm = max(a.warmup, 3)
    sampler = ClockSampler(local)   # samples cover the warm-up and the timed region (same load)
    sampler.start()
    ep.advance(warm)
    barrier()
    l0 = bw.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ep.advance(a.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = bw.launch_count() - l0
    if not sampler.rows:            # a very short timed region: keep the load up until one sample exists
        t_end = time.time() + 10.
        while not sampler.rows and time.time() < t_end:
            ep.advance(PHASE)
            torch.cuda.synchronize()
    sampler.stop_flag = True
    nonfinite = int((~torch.isfinite(bw.gvel)).any(0).sum())

    # ---- per-stage device time (diagnostic pass, CUDA events around every kernel) ----------
    bw.set_option("time_stages", 1)
    ep.advance(PHASE)
    torch.cuda.synchronize()
    st = bw.stage_ms()
    bw.set_option("time_stages", 0)

    # ---- end to end through the host-buffer entry point (arb_step_host) --------------------
    e2e = None
    if not a.no_e2e:
        nrows = max(int(model.nrows), 1)
        hg = torch.empty(gp.shape, dtype=torch.float64).pin_memory()
        hv = torch.empty(gv.shape, dtype=torch.float64).pin_memory()
        hf = torch.zeros((nrows, W), dtype=torch.float64).pin_memory()
        hg.copy_(bw.gpos); hv.copy_(bw.gvel); hf.copy_(bw.cforce)    # continue the same episodes
        hgn, hvn, hfn = hg.numpy(), hv.numpy(), hf.numpy()
        hep = Episodes.__new__(Episodes)
        hep.blocks, hep.t = ep.blocks, ep.t
        hep.state, hep.init = (hg, hv, hf), (torch.as_tensor(gp), torch.as_tensor(gv), None)

        pipe = None
        if isinstance(a.e2e_chunks, list) or a.e2e_chunks > 1:
            # the public end-to-end call: column blocks of the host state, one stream each, so
            # that copies and kernels of different blocks overlap (batch.HostPipeline)
            pipe = HostPipeline(model, W, chunks=a.e2e_chunks, device=bw.device, mode=a.e2e_mode,
                                compute_streams=a.e2e_compute_streams)
            for opt in (a.opt or []):
                name, val = opt.split("=")
                pipe.set_option(name, int(val))

        def host_steps(c):
            for _ in range(c):
                if pipe is not None:
                    pipe.step(hgn, hvn, hfn if model.nrows else None, DT, 1)
                else:
                    bw.step_host(hgn, hvn, hfn, DT, 1)    # H2D state, 1 step, D2H state, sync
        hep.step_fn = host_steps
        k_e2e = max(3, min(a.steps, 100))
        hep.advance(max(3, min(a.warmup, 8)))    # the blocks' first sorts happen here
        barrier()
        t0 = time.perf_counter()
        hep.advance(k_e2e)
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        bytes_in = (hgn.nbytes + hvn.nbytes + (hfn.nbytes if model.nrows else 0))
        e2e = [t_e2e/k_e2e, bytes_in, bytes_in, k_e2e]

    # max over ranks of the timed region; totals over ranks
    (ms_all, e2e_ms), (nonfinite, launches, total_worlds) = reduce_report(
        [ms, e2e[0]*1e3 if e2e else 0.], [nonfinite, launches, W], device=bw.device)
    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return
    total_worlds = int(total_worlds)
    value = total_worlds*a.steps/(ms_all*1e-3)
    lib = _capi.load()
    peak = C_double()
    lib.arb_measure_fp64_peak(local, peak.ref())
    fp64_peak = peak.value
    flop = FLOP_PER_WORLD_STEP[scen]
    achieved = (value/world_size)*flop        # per GPU, flop/s
    peaks, traffic = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(scen, {})
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.))
    nst = max(st["steps"], 1)
    stage = {k: st[k]/nst for k in ("prepare", "gs", "finish")}
    tot_stage = sum(stage.values()) or 1.
    dom = max(stage, key=stage.get)
    dram_per_world = traffic.get("dram_bytes_per_world_step")
    out = {
        "metric": "world-steps/s (human36, fp64, dt=1ms)", "value": value, "unit": "world-steps/s",
        "n_gpus": world_size, "steps": a.steps, "warmup": warm, "ms_per_step": ms_all/a.steps,
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (seeded random initial states, SURVEY.md 8(d))",
        "config": {"workload": a.workload, "scenario": scen, "worlds_total": total_worlds,
                   "worlds_per_gpu": W, "dt": DT, "constraints": int(model.nc),
                   "episode_steps": EPISODE, "episode_groups": GROUPS,
                   "episodes": "worlds restart from their seeded state every %d steps; %d blocks of "
                               "the batch are staggered by %d steps so every timed step sees the "
                               "whole episode's mix of contact states" % (EPISODE, GROUPS, PHASE),
                   "l2": "state of all worlds (%.0f MB) and per-world scratch exceed L2; no flush needed"
                         % (W*STATE_BYTES_PER_WORLD_STEP/2/1e6),
                   "parallelism": "worlds sharded over %d GPU(s) (%s scaling: %s), no collective on the "
                                  "step path" % (world_size, a.scaling,
                                                 "every GPU holds the workload's number of worlds"
                                                 if a.scaling == "weak" else
                                                 "the workload's worlds are split over the GPUs")},
        "gpu_launches": int(launches),
        "nonfinite_worlds": int(nonfinite),
        "roofline": {"bound": "fp64", "achieved": achieved/1e12, "peak": fp64_peak/1e12,
                     "unit": "TFLOP/s", "frac": achieved/fp64_peak if fp64_peak else None,
                     "traffic": dram_per_world*W if dram_per_world else None,
                     "traffic_note": traffic.get("source"),
                     "peak_source": "DFMA micro-benchmark measured in this run (arb_measure_fp64_peak); "
                                    "MEASURED_PEAKS.json has no fp64 entry",
                     "flop_per_world_step": flop,
                     "per_launch": "one (prepare, gs, finish) triple = one step of the %d worlds of a GPU; "
                                   "achieved = worlds x flop_per_world_step / step time" % W,
                     "stage_ms": stage, "dominant_kernel": "k_fused_" + dom,
                     "dominant_share": stage[dom]/tot_stage,
                     "hbm": {"achieved": value/world_size*STATE_BYTES_PER_WORLD_STEP/1e9,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": value/world_size*STATE_BYTES_PER_WORLD_STEP/1e9/hbm_peak,
                             "dram_achieved": (dram_per_world*value/world_size/1e9
                                               if dram_per_world else None)}},
        "clocks": sampler.summary(),
    }
    if e2e:
        out["e2e"] = {"value": total_worlds/(e2e_ms*1e-3), "unit": "world-steps/s",
                      "h2d_bytes_per_step": e2e[1]*world_size, "d2h_bytes_per_step": e2e[2]*world_size,
                      "steps": e2e[3], "chunks": a.e2e_chunks,
                      "mode": a.e2e_mode,
                      "how": ("HostPipeline.step (%d column blocks of the pinned host state; %s): host -> "
                              "device, 1 step, device -> host, all blocks synchronised, every step; same "
                              "staggered episodes as the timed region"
                              % (len(a.e2e_chunks) if isinstance(a.e2e_chunks, list) else a.e2e_chunks,
                                 ("kernels of all blocks on %d streams, block after block, "
                                  "arb_state_copy_host_strided copies on two more streams ordered by events"
                                  % a.e2e_compute_streams) if a.e2e_mode == "serial"
                                 else "arb_step_host_strided, one stream per block"))
                             if (isinstance(a.e2e_chunks, list) or a.e2e_chunks > 1) else
                             "arb_step_host: pinned host state -> device, 1 step, device -> host, "
                             "synchronised, every step; same staggered episodes as the timed region"}
    if not a.no_cpu_baseline and world_size == 1:      # rank 0 at N = 1 only
        out["cpu_baseline"] = cpu_baseline(scen, a.cpu_seconds)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(out), flush=True)
    os.dup2(2, 1)
    if world_size > 1:
        dist.destroy_process_group()


class C_double(object):
    def __init__(self):
        import ctypes
        self._c = ctypes.c_double(0.)
        self._ct = ctypes

    def ref(self):
        return self._ct.byref(self._c)

    value = property(lambda self: self._c.value)


def run_reference(a):
    """Reference arm: the reference's CPU implementation of the path (the numpy oracle
    port -- the Python reference cannot travel to the GPU box) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scen, total = WORKLOADS[a.workload]
    if a.worlds:
        total = a.worlds
    if a.scaling == "weak":
        total *= max(a.gpus, 1)
    base = cpu_baseline(scen, a.cpu_seconds)
    v = base["value"]
    out = {
        "impl": "reference", "metric": "world-steps/s (human36, fp64, dt=1ms)", "value": v,
        "unit": "world-steps/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3/v if v else None, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (seeded random initial states)",
        "config": {"workload": a.workload, "scenario": scen, "worlds_total": total, "dt": DT},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
