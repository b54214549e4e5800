"""bench.py's host-side helpers: the SURVEY.md 8(d) flop formulas (pinned to the survey's own
worked numbers) and the config object both arms print."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from arboris_b200 import scenarios  # noqa: E402
from arboris_b200.flatten import flatten  # noqa: E402


def test_flop_formulas_reproduce_survey_numbers():
    """human36 free step 360 118 flop; with 8 active contacts the increment is
    65 856 + 198 912 + 9 744 + 353 280; snake9 111 080."""
    m = flatten(scenarios.human36_contact_world())
    assert bench.flop_free(m) == 360118.
    act = np.zeros((3, 10))
    act[0, :8] = 1          # the 8 contacts, no knee limit
    act[2, :] = 1           # everything
    inc = bench.flop_contact_increment(m, act)
    assert inc[0] == 65856 + 198912 + 9744 + 353280
    assert inc[1] == 0.     # no active constraint: no contact work
    assert inc[2] > inc[0]
    s = flatten(scenarios.snake_loop_world())
    # (the survey's snake has no constraints and 9 links + free base: k_b = 6..15)
    assert bench.flop_free(s) == 111080.


def test_both_arms_print_the_same_config_keys():
    class A(object):
        workload = "human36_contact_262144"
    m = flatten(scenarios.human36_contact_world())
    ours = bench.config_of(A, "human36_contact", 262144, 262144, 1, "weak", m)
    ref = bench.config_of(A, "human36_contact", 262144, 262144, 1, "weak", m)
    assert set(ours) == set(ref) and ours["distinct_worlds"] == 262144 and ours["constraints"] == 10


def test_auto_chunks_follow_batch_size():
    from arboris_b200.batch import HostPipeline
    assert HostPipeline.auto_chunks(32768) == 1
    assert HostPipeline.auto_chunks(131072) == (1, 2, 1)
    assert HostPipeline.auto_chunks(262144) == (1, 1, 2, 2, 2, 1, 1)


def test_traffic_json_follows_from_the_committed_ncu_summaries():
    """profiles/traffic.json (what bench.py reports as roofline.traffic and frac_executed) can be
    recomputed from the ncu summaries committed beside it (profiles/make_traffic.py)."""
    import importlib.util
    import json
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prof = os.path.join(root, "profiles")
    doc = json.load(open(os.path.join(prof, "traffic.json")))["human36_contact"]
    tag = re.search(r"profiles/(\w+)_ncu_", doc["source"]).group(1)
    spec = importlib.util.spec_from_file_location("make_traffic", os.path.join(prof, "make_traffic.py"))
    mt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mt)
    W = 262144
    tot_b = tot_f = 0.
    for st in ("prepare", "gs", "finish"):
        r = mt.read(os.path.join(prof, "%s_ncu_%s_%d.txt" % (tag, st, W)))
        b = (r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"])/W
        f = (2*r["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
             + r["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
             + r["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"])/W
        assert abs(doc["per_kernel_bytes_per_world"][r["kernel"]] - b) <= 1e-9*b
        assert abs(doc["per_kernel_executed_fp64_flop_per_world"][r["kernel"]] - f) <= 1e-9*f
        tot_b += b
        tot_f += f
    assert abs(doc["dram_bytes_per_world_step"] - tot_b) <= 1e-9*tot_b
    assert abs(doc["executed_fp64_flop_per_world_step"] - tot_f) <= 1e-9*tot_f
