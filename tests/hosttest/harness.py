"""TEST INFRASTRUCTURE ONLY: ctypes driver of the CPU build of the per-world
routines (see arb_hosttest.cpp).  Mirrors the phase API on host numpy arrays with
the device layouts ([elem][W])."""
import ctypes as C
import os
import subprocess

import numpy as np

from arboris_b200 import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "arb_hosttest.cpp")
OUT = os.path.join(HERE, "_build", "libarb_hosttest.so")
CSRC = os.path.join(ROOT, "arboris-python_b200", "csrc")


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                           "-Wno-unknown-pragmas", "-o", OUT, SRC])
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ht_create.restype = C.c_void_p
        L.ht_create.argtypes = [C.POINTER(_capi.ModelDesc), C.c_int64, C.c_char_p, C.c_int]
        L.ht_destroy.argtypes = [C.c_void_p]
        L.ht_bind.argtypes = [C.c_void_p] * 4
        for f in ("ht_update_dynamic",):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("ht_update_controllers", "ht_update_constraints", "ht_integrate", "ht_fused_step"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_double]
        L.ht_array.restype = C.POINTER(C.c_double)
        L.ht_array.argtypes = [C.c_void_p, C.c_char_p]
        L.ht_iarray.restype = C.POINTER(C.c_int)
        L.ht_iarray.argtypes = [C.c_void_p, C.c_char_p]
        L.ht_fused_ints.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        _lib = L
    return _lib


class HostBatch(object):
    def __init__(self, model, W):
        self.L = lib()
        self.model, self.W = model, W
        self.desc, self._keep = _capi.make_desc(model)
        err = C.create_string_buffer(256)
        self.h = self.L.ht_create(C.byref(self.desc), W, err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.gpos = np.zeros((model.ngpos, W))
        self.gvel = np.zeros((model.ndof, W))
        self.cforce = np.zeros((max(model.nrows, 1), W))
        self.L.ht_bind(self.h, self.gpos.ctypes.data, self.gvel.ctypes.data, self.cforce.ctypes.data)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ht_destroy(self.h)
            self.h = None

    def aligned(self):
        """(flags of the generator bodies, flags of the constraints): contact-aligned blocks."""
        g = np.zeros(64, np.int32)
        c = np.zeros(max(len(self.model.cons_type), 1), np.int32)
        self.L.ht_aligned.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        n = self.L.ht_aligned(self.h, g.ctypes.data, c.ctypes.data)
        return g[:n].copy(), c[:len(self.model.cons_type)].copy()

    def bind_params(self, kp, kd, qd, dqd):
        """per-world PD parameters, (npd, W) C-contiguous arrays (kept alive by the caller)"""
        self._params = [np.ascontiguousarray(a, dtype=np.float64) for a in (kp, kd, qd, dqd)]
        self.L.ht_bind_params.argtypes = [C.c_void_p] * 5
        self.L.ht_bind_params(self.h, *[a.ctypes.data for a in self._params])

    def update_dynamic(self):
        self.L.ht_update_dynamic(self.h)

    def update_controllers(self, dt):
        self.L.ht_update_controllers(self.h, dt)

    def update_constraints(self, dt):
        self.L.ht_update_constraints(self.h, dt)

    def integrate(self, dt):
        self.L.ht_integrate(self.h, dt)

    def set_coop(self, v):
        """1 (default): block-cooperative Gauss-Seidel form; 0: per-lane form."""
        self.L.ht_set_coop.argtypes = [C.c_void_p, C.c_int]
        self.L.ht_set_coop(self.h, int(v))

    def fused_step(self, dt):
        self.L.ht_fused_step(self.h, dt)

    def fused_step_group(self, dt, write_poses=0):
        """group prepare stage (lanes emulated) + per-lane Gauss-Seidel + K-matrix finish"""
        self.L.ht_fused_step_group.argtypes = [C.c_void_p, C.c_double, C.c_int]
        self.L.ht_fused_step_group(self.h, dt, int(write_poses))

    def group_doubles(self):
        self.L.ht_group_doubles.argtypes = [C.c_void_p]
        return int(self.L.ht_group_doubles(self.h))

    def arr(self, name, *shape):
        p = self.L.ht_array(self.h, name.encode())
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(p, shape=(n * self.W,)).reshape(shape + (self.W,))

    def iarr(self, name, *shape):
        if name in ("factive", "fbranch"):      # tiled fused scratch: copied out as [elem][W]
            n = int(np.prod(shape))
            out = np.zeros((n, self.W), dtype=np.int32)
            self.L.ht_fused_ints(self.h, name.encode(), n, out.ctypes.data_as(C.POINTER(C.c_int)))
            return out.reshape(shape + (self.W,))
        p = self.L.ht_iarray(self.h, name.encode())
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(p, shape=(n * self.W,)).reshape(shape + (self.W,))
