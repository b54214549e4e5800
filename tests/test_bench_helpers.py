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
