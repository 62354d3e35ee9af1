"""ctypes binding of libOpt.so (include/Opt.h, include/OptB200.h, include/OptB200Examples.h).  Plain C ABI, no torch types."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OptError(RuntimeError):
    pass


class Opt_InitializationParameters(ctypes.Structure):   # reference Opt.h:10-30
    _fields_ = [("doublePrecision", ctypes.c_int), ("verbosityLevel", ctypes.c_int),
                ("collectPerKernelTimingInfo", ctypes.c_int), ("threadsPerBlock", ctypes.c_int)]


# every symbol the two public headers declare: (name, restype, argtypes)
_VP, _VPP, _CP = ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p
SYMBOLS = [
    ("Opt_NewState", _VP, [Opt_InitializationParameters]),
    ("Opt_ProblemDefine", _VP, [_VP, _CP, _CP]),
    ("Opt_ProblemDelete", None, [_VP, _VP]),
    ("Opt_ProblemPlan", _VP, [_VP, _VP, ctypes.POINTER(ctypes.c_uint)]),
    ("Opt_PlanFree", None, [_VP, _VP]),
    ("Opt_SetSolverParameter", None, [_VP, _VP, _CP, _VP]),
    ("Opt_ProblemSolve", None, [_VP, _VP, _VPP]),
    ("Opt_ProblemInit", None, [_VP, _VP, _VPP]),
    ("Opt_ProblemStep", ctypes.c_int, [_VP, _VP, _VPP]),
    ("Opt_ProblemCurrentCost", ctypes.c_double, [_VP, _VP]),
    ("OptB200_Version", _CP, []),
    ("OptB200_PlanNumUnknowns", ctypes.c_long, [_VP]),
    ("OptB200_PlanUnknownLayout", ctypes.c_int, [_VP, ctypes.c_int, ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_int)]),
    ("OptB200_EvalJTF", None, [_VP, _VPP, ctypes.POINTER(ctypes.c_double), _VP, _VP, ctypes.POINTER(ctypes.c_double)]),
    ("OptB200_ApplyJTJ", None, [_VP, _VPP, _VP, _VP, ctypes.POINTER(ctypes.c_double)]),
    ("OptB200_SetMaxLinearIterations", None, [_VP, ctypes.c_int]),
    ("OptB200_ReadVector", ctypes.c_int, [_VP, _CP, _VP]),
    ("OptB200_ReadScalar", ctypes.c_double, [_VP, _CP]),
    ("OptB200_PlanSetStream", None, [_VP, _VP]),
    ("OptB200_LaunchesLastStep", ctypes.c_int, [_VP]),
    ("OptB200_PcgLoopMsLastStep", ctypes.c_float, [_VP]),
    ("OptB200_ProfileKernels", ctypes.c_int, [_VP, _VPP, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_float)]),
    ("OptB200_SpecEnergyName", ctypes.c_int, [_CP, ctypes.c_char_p, ctypes.c_int]),
    ("OptB200_FrontendCompile", ctypes.c_int, [_CP, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]),
    ("OptB200_PlanCommHandles", ctypes.c_int, [_VP, _VP, ctypes.c_int]),
    ("OptB200_PlanConnect", ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    # include/OptB200Examples.h: the example harnesses' per-outer-iteration pre-steps on the device
    ("OptB200_ImageWarpingConstraints", ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, _VP, ctypes.c_int, ctypes.c_float, ctypes.c_int, _VP]),
    ("OptB200_ArapConstraints", ctypes.c_int, [_VP, _VP, ctypes.c_int, _VP, _VP, ctypes.c_int, ctypes.c_float, ctypes.c_int, _VP]),
    ("OptB200_VertexNormals", ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP, ctypes.c_int, _VP, _VP]),
    ("OptB200_PointGridCreate", _VP, [_VP, ctypes.c_int, ctypes.c_float, _VP]),
    ("OptB200_PointGridFree", None, [_VP]),
    ("OptB200_RnaCorrespondences", ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int, _VP, _VP, _VP, _VP, ctypes.c_int, _VP]),
]


def lib_path():
    # OPTB_LIBOPT: an alternative build of the same library (kernel-variant A/B runs, tools/build_variants.sh); never a fallback
    return os.environ.get("OPTB_LIBOPT") or os.path.join(_HERE, "lib", "libOpt.so")


def lib():
    """Loads libOpt.so.  Raises OptError (never falls back to anything) when it has not been built."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise OptError(f"{p} not found: run `python -c 'import __graft_entry__ as g; g.build()'` first; "
                           "opt_b200 has no CPU or PyTorch fallback path")
        L = ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)       # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB
