"""The C-ABI shared library loads on a CPU-only box and exports every symbol
include/arboris_b200.h declares (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def _header_functions():
    src = open(os.path.join(ROOT, "include", "arboris_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(arb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from arboris_b200 import _capi
    return _capi.load()


def test_exports_match_header(lib):
    from arboris_b200 import _capi
    declared = _header_functions()
    bound = sorted(name for name, _, _ in _capi.SYMBOLS)
    assert declared == bound
    for name in declared:
        assert hasattr(lib, name), name


def test_model_create_validates_without_gpu(lib):
    """arb_model_create is host-only: good models are accepted, broken ones give an
    error code and a message (no exception crosses the ABI)."""
    from arboris_b200 import _capi
    model, _ = load_golden("human36_contact")
    desc, keep = _capi.make_desc(model)
    h = C.c_void_p()
    assert lib.arb_model_create(C.byref(desc), C.byref(h)) == 0
    lib.arb_model_destroy(h)
    bad = np.array(model.joint_type, dtype=np.int32).copy()
    bad[3] = 42
    desc.joint_type = bad.ctypes.data_as(_capi.c_i32p)
    assert lib.arb_model_create(C.byref(desc), C.byref(h)) < 0
    assert b"joint type" in lib.arb_last_error()
    assert lib.arb_model_create(None, C.byref(h)) < 0


def test_model_create_validates_controller_and_limit_tables(lib):
    """descriptor contents that come through the public ABI are checked before any
    kernel can index with them: PD blob offset / size, its dof and gpos maps, and the dof / gpos index a
    JointLimits row carries for its joint.  Plus the per-world PD parameter rows the ABI reports."""
    from arboris_b200 import _capi
    model, _ = load_golden("zoo")
    h = C.c_void_p()

    def create(**edit):
        import copy
        m = copy.copy(model)
        for k, v in edit.items():
            setattr(m, k, v)
        desc, keep = _capi.make_desc(m)
        rc = lib.arb_model_create(C.byref(desc), C.byref(h))
        return rc, lib.arb_last_error()

    rc, _ = create()
    assert rc == 0
    dofs = (C.c_int32*16)()
    npd = lib.arb_model_pd_dofs(h, dofs, 16)
    assert npd == 5 and sorted(dofs[i] for i in range(npd)) == sorted(set(dofs[i] for i in range(npd)))
    lib.arb_model_destroy(h)
    pd = [a for a in range(len(model.ctrl_type)) if int(model.ctrl_type[a]) == 1][0]
    ci = np.array(model.ctrl_int, dtype=np.int32).copy()
    ci[pd][1] = int(np.asarray(model.ctrl_blob).size) - 2          # blob offset past the end
    rc, msg = create(ctrl_int=ci)
    assert rc < 0 and b"blob" in msg
    blob = np.array(model.ctrl_blob, dtype=float).copy()
    blob[int(model.ctrl_int[pd][1])] = 10_000.                      # dof map entry out of range
    rc, msg = create(ctrl_blob=blob)
    assert rc < 0 and b"out of range" in msg
    lim = [c for c in range(len(model.cons_type)) if int(model.cons_type[c]) == 0][0]
    ki = np.array(model.cons_int, dtype=np.int32).copy()
    ki[lim][1] += 1                                                  # dof index that is not the joint's
    rc, msg = create(cons_int=ki)
    assert rc < 0 and b"JointLimits" in msg


def test_step_without_gpu_fails_loudly():
    """No CUDA device -> the product raises; it never falls back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from arboris_b200.batch import BatchedWorld
    model, _ = load_golden("simplearm")
    with pytest.raises(RuntimeError):
        BatchedWorld(model, 4)
    from arboris_b200 import scenarios
    w = scenarios.simplearm_world()
    with pytest.raises(RuntimeError):
        w.update_dynamic()
