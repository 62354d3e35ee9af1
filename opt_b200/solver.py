"""Python mirror of the reference harness class OptSolver (examples/shared/OptSolver.h:40-96) and of
NamedParameters::data() (examples/shared/NamedParameters.h:34-47): positional `void*` tables with
DEVICE pointers for images / graph index arrays and HOST pointers for scalars and edge counts."""
import ctypes
import os

import numpy as np

from . import capi

_INT_PARAMS = ("nIterations", "lIterations", "residual_reset_period")
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def spec_path(name):
    """path of an energy spec shipped with this repository (specs/<name>.t)"""
    return os.path.join(_ROOT, "specs", name if name.endswith(".t") else name + ".t")


class ParamTable:
    """void** built from torch CUDA tensors (device ptr), python floats (host float*), ints (host int*)."""

    def __init__(self, values):
        self.keep = list(values)
        self.arr = (ctypes.c_void_p * len(values))()
        for i, v in enumerate(values):
            if hasattr(v, "data_ptr"):
                if not v.is_cuda:
                    raise capi.OptError(f"problem parameter {i}: images and graph indices must be CUDA tensors")
                if not v.is_contiguous():
                    raise capi.OptError(f"problem parameter {i}: tensor must be contiguous")
                self.arr[i] = v.data_ptr()
            elif isinstance(v, (int, np.integer)) and not isinstance(v, bool):
                c = ctypes.c_int(int(v)); self.keep.append(c)
                self.arr[i] = ctypes.cast(ctypes.pointer(c), ctypes.c_void_p).value
            elif isinstance(v, (float, np.floating)):
                c = ctypes.c_float(float(v)); self.keep.append(c)
                self.arr[i] = ctypes.cast(ctypes.pointer(c), ctypes.c_void_p).value
            else:
                raise TypeError(f"problem parameter {i}: unsupported {type(v)}")

    @property
    def ptr(self):
        return ctypes.cast(self.arr, ctypes.POINTER(ctypes.c_void_p))


class OptSolver:
    """OptSolver(dimensions, terraFile, optName, doublePrecision) -- same argument meaning as the
    reference constructor (OptSolver.h:43-56)."""

    def __init__(self, dimensions, terra_file, opt_name="gaussNewtonGPU", double_precision=False, verbosity=0, timing=0):
        self.L = capi.lib()
        ip = capi.Opt_InitializationParameters(int(bool(double_precision)), int(verbosity), int(timing), 0)
        self.state = self.L.Opt_NewState(ip)
        if not self.state:
            raise capi.OptError("Opt_NewState failed (no CUDA device?)")
        self.problem = self.L.Opt_ProblemDefine(self.state, os.fsencode(terra_file), opt_name.encode())
        dims = (ctypes.c_uint * len(dimensions))(*[int(d) for d in dimensions])
        self.plan = self.L.Opt_ProblemPlan(self.state, self.problem, dims)
        if not self.plan:
            raise capi.OptError(f"Opt_ProblemPlan failed for {terra_file} ({opt_name})")
        self.double_precision = bool(double_precision)
        self.np_dtype = np.float64 if double_precision else np.float32

    def close(self):
        if getattr(self, "plan", None):
            self.L.Opt_PlanFree(self.state, self.plan); self.plan = None
        if getattr(self, "problem", None):
            self.L.Opt_ProblemDelete(self.state, self.problem); self.problem = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Opt.h
    def set_solver_parameter(self, name, value):
        c = ctypes.c_int(int(value)) if name in _INT_PARAMS else ctypes.c_float(float(value))
        self.L.Opt_SetSolverParameter(self.state, self.plan, name.encode(), ctypes.cast(ctypes.pointer(c), ctypes.c_void_p))

    def set_solver_parameters(self, params):
        for k, v in (params or {}).items():
            self.set_solver_parameter(k, v)

    def init(self, problem_params):
        t = ParamTable(problem_params); self.L.Opt_ProblemInit(self.state, self.plan, t.ptr)

    def step(self, problem_params):
        t = ParamTable(problem_params); return self.L.Opt_ProblemStep(self.state, self.plan, t.ptr)

    def solve(self, solver_params, problem_params):
        """OptSolver::solve (OptSolver.h:72-91): set parameters, Opt_ProblemSolve, return the final cost."""
        self.set_solver_parameters(solver_params)
        t = ParamTable(problem_params)
        self.L.Opt_ProblemSolve(self.state, self.plan, t.ptr)
        return self.current_cost()

    def current_cost(self):
        return self.L.Opt_ProblemCurrentCost(self.state, self.plan)

    # ---- OptB200.h
    def num_unknowns(self):
        return self.L.OptB200_PlanNumUnknowns(self.plan)

    def unknown_layout(self):
        off = (ctypes.c_long * 8)(); el = (ctypes.c_long * 8)(); ch = (ctypes.c_int * 8)()
        k = self.L.OptB200_PlanUnknownLayout(self.plan, 8, off, el, ch)
        return [(off[i], el[i], ch[i]) for i in range(k)]

    def unflatten(self, flat):
        """flat padded plan vector -> list of per-unknown-array numpy views (elements, channels)"""
        return [flat[o:o + e * c].reshape(e, c) for (o, e, c) in self.unknown_layout()]

    def flatten(self, arrays):
        flat = np.zeros(self.num_unknowns(), self.np_dtype)
        for a, (o, e, c) in zip(arrays, self.unknown_layout()):
            flat[o:o + e * c] = np.asarray(a, self.np_dtype).reshape(-1)
        return flat

    def eval_jtf(self, problem_params):
        t = ParamTable(problem_params)
        n = self.num_unknowns()
        r = np.zeros(n, self.np_dtype); pre = np.zeros(n, self.np_dtype)
        cost = ctypes.c_double(0); rdp = ctypes.c_double(0)
        self.L.OptB200_EvalJTF(self.plan, t.ptr, ctypes.byref(cost), r.ctypes.data, pre.ctypes.data, ctypes.byref(rdp))
        return cost.value, r, pre, rdp.value

    def apply_jtj(self, problem_params, x):
        t = ParamTable(problem_params)
        x = np.ascontiguousarray(x, self.np_dtype); out = np.zeros_like(x)
        dot = ctypes.c_double(0)
        self.L.OptB200_ApplyJTJ(self.plan, t.ptr, x.ctypes.data, out.ctypes.data, ctypes.byref(dot))
        return out, dot.value

    def set_max_linear_iterations(self, k):
        self.L.OptB200_SetMaxLinearIterations(self.plan, int(k))

    def read_vector(self, name):
        out = np.zeros(self.num_unknowns(), self.np_dtype)
        if self.L.OptB200_ReadVector(self.plan, name.encode(), out.ctypes.data) != 0:
            raise KeyError(name)
        return out

    def read_scalar(self, name):
        return self.L.OptB200_ReadScalar(self.plan, name.encode())

    def launches_last_step(self):
        return self.L.OptB200_LaunchesLastStep(self.plan)

    def pcg_loop_ms_last_step(self):
        return self.L.OptB200_PcgLoopMsLastStep(self.plan)

    def comm_handles(self):
        """CUDA IPC handles of this plan's scalar mailbox block, z, p and delta (OptB200_PlanCommHandles: 4 x 64 bytes)"""
        buf = ctypes.create_string_buffer(1024)
        n = self.L.OptB200_PlanCommHandles(self.plan, buf, 1024)
        if n <= 0:
            raise capi.OptError("OptB200_PlanCommHandles failed")
        return buf.raw[:n]

    def connect(self, rank, world, all_handles, owned_begin, owned_end, upper_rows, lower_rows):
        rc = self.L.OptB200_PlanConnect(self.plan, int(rank), int(world), all_handles, int(owned_begin), int(owned_end), int(upper_rows), int(lower_rows))
        if rc != 0:
            raise capi.OptError(f"OptB200_PlanConnect failed ({rc})")

    def profile_kernels(self, problem_params, reps=20):
        """{kernel name: average ms per launch}, CUDA events on the plan's stream (OptB200_ProfileKernels)"""
        t = ParamTable(problem_params)
        names = (ctypes.c_char_p * 16)(); ms = (ctypes.c_float * 16)()
        k = self.L.OptB200_ProfileKernels(self.plan, t.ptr, int(reps), 16, names, ms)
        return {names[i].decode(): float(ms[i]) for i in range(k)}


def spec_energy_name(filename):
    """front-end probe: (ok, energy-name-or-message); needs no GPU"""
    L = capi.lib()
    buf = ctypes.create_string_buffer(512)
    rc = L.OptB200_SpecEnergyName(os.fsencode(filename), buf, 512)
    return rc == 0, buf.value.decode()


def frontend_compile(filename, double_precision=False):
    """general front end, host only: (ok, message) of interpreting the .t file and NVRTC-compiling its residual code"""
    L = capi.lib()
    buf = ctypes.create_string_buffer(16384)
    rc = L.OptB200_FrontendCompile(os.fsencode(filename), int(bool(double_precision)), buf, 16384)
    return rc == 0, buf.value.decode(errors="replace")
