"""opt_b200 -- Blackwell-native drop-in for the solver core of niessner/Opt.

The product is `lib/libOpt.so` (C ABI of include/Opt.h + include/OptB200.h, hand-written sm_100a
kernels, built by `__graft_entry__.build()`).  This package is the thin Python host side: a ctypes
binding (`capi`) and a mirror of the reference's C++ harness class `OptSolver`
(examples/shared/OptSolver.h:40-96) so tests and benchmarks read like the reference's callers.
There is no CPU fallback: importing `capi` without the built library raises.
"""
from .capi import lib, lib_path, OptError  # noqa: F401
from .solver import OptSolver, spec_path  # noqa: F401

__version__ = "0.1.0"
