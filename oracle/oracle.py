"""ctypes binding of the CPU oracle (oracle/opt_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product (opt_b200) never does.

The calling convention mirrors Opt.h (reference API/release/include/Opt.h:35-71) with HOST pointers:
`problemparams` is a positional list in the order the .t spec declares -- numpy arrays for images
and graph index arrays, python floats for `Param(..., float, i)` scalars, python ints for graph
edge counts.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libopt_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("opt_oracle.cpp", "opt_oracle_energies2.inc", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.oracle_create.restype = ctypes.c_void_p
        L.oracle_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint), ctypes.c_int]
        L.oracle_free.argtypes = [ctypes.c_void_p]
        L.oracle_set_solver_parameter.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        for f in (L.oracle_init, L.oracle_solve):
            f.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
        L.oracle_step.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
        L.oracle_step.restype = ctypes.c_int
        L.oracle_current_cost.argtypes = [ctypes.c_void_p]
        L.oracle_current_cost.restype = ctypes.c_double
        L.oracle_num_unknowns.argtypes = [ctypes.c_void_p]
        L.oracle_num_unknowns.restype = ctypes.c_long
        L.oracle_eval_jtf.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_double), ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_apply_jtj.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_get_vector.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        L.oracle_get_vector.restype = ctypes.c_int
        L.oracle_get_scalar.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.oracle_get_scalar.restype = ctypes.c_double
        L.oracle_set_max_linear_iterations.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oracle_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


_INT_PARAMS = ("nIterations", "lIterations", "residual_reset_period")


class ParamPack:
    """Builds the `void**` array from python values and keeps the backing storage alive."""

    def __init__(self, values):
        self.keep = []
        n = len(values)
        self.arr = (ctypes.c_void_p * n)()
        for i, v in enumerate(values):
            if isinstance(v, np.ndarray):
                assert v.flags["C_CONTIGUOUS"]
                self.keep.append(v)
                self.arr[i] = v.ctypes.data
            elif isinstance(v, (int, np.integer)) and not isinstance(v, bool):
                c = ctypes.c_int(int(v)); self.keep.append(c)
                self.arr[i] = ctypes.cast(ctypes.pointer(c), ctypes.c_void_p).value
            elif isinstance(v, (float, np.floating)):
                c = ctypes.c_float(float(v)); self.keep.append(c)
                self.arr[i] = ctypes.cast(ctypes.pointer(c), ctypes.c_void_p).value
            else:
                raise TypeError(f"param {i}: unsupported {type(v)}")

    @property
    def ptr(self):
        return ctypes.cast(self.arr, ctypes.POINTER(ctypes.c_void_p))


class Oracle:
    def __init__(self, energy, solverkind, dims, double_precision=False):
        self.L = lib()
        d = (ctypes.c_uint * len(dims))(*[int(x) for x in dims])
        self.h = self.L.oracle_create(energy.encode(), solverkind.encode(), d, int(bool(double_precision)))
        if not self.h:
            raise ValueError(f"oracle: unknown energy {energy}")
        self.dtype = np.float64 if double_precision else np.float32

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_free(self.h); self.h = None

    def set_solver_parameter(self, name, value):
        c = ctypes.c_int(int(value)) if name in _INT_PARAMS else ctypes.c_float(float(value))
        self.L.oracle_set_solver_parameter(self.h, name.encode(), ctypes.byref(c))

    def init(self, params):
        pk = ParamPack(params); self.L.oracle_init(self.h, pk.ptr)

    def step(self, params):
        pk = ParamPack(params); return self.L.oracle_step(self.h, pk.ptr)

    def solve(self, params):
        pk = ParamPack(params); self.L.oracle_solve(self.h, pk.ptr)
        return self.cost()

    def cost(self):
        return self.L.oracle_current_cost(self.h)

    def num_unknowns(self):
        return self.L.oracle_num_unknowns(self.h)

    def eval_jtf(self, params):
        """returns (cost, JtF, diag(JtJ)) at the current unknowns; flat vectors, unknown arrays concatenated in spec order"""
        pk = ParamPack(params)
        # binding happens inside; num_unknowns is known after the first bind -> allocate generously via a probe call
        cost = ctypes.c_double(0)
        self.L.oracle_eval_jtf(self.h, pk.ptr, ctypes.byref(cost), None, None)
        n = self.num_unknowns()
        g = np.zeros(n, self.dtype); pre = np.zeros(n, self.dtype)
        self.L.oracle_eval_jtf(self.h, pk.ptr, ctypes.byref(cost), g.ctypes.data, pre.ctypes.data)
        return cost.value, g, pre

    def apply_jtj(self, params, p):
        pk = ParamPack(params)
        p = np.ascontiguousarray(p, self.dtype)
        out = np.zeros_like(p)
        self.L.oracle_apply_jtj(self.h, pk.ptr, p.ctypes.data, out.ctypes.data)
        return out

    def get_vector(self, name):
        out = np.zeros(self.num_unknowns(), self.dtype)
        rc = self.L.oracle_get_vector(self.h, name.encode(), out.ctypes.data)
        if rc != 0:
            raise KeyError(name)
        return out

    def get_scalar(self, name):
        return self.L.oracle_get_scalar(self.h, name.encode())

    def set_max_linear_iterations(self, k):
        self.L.oracle_set_max_linear_iterations(self.h, int(k))


def num_threads():
    return lib().oracle_num_threads()
