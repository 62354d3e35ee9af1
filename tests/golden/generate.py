#!/usr/bin/env python
"""Generates tests/golden/*.npz: per energy, the cost history and final unknowns of an fp64 oracle solve
on a small deterministic instance (inputs are regenerated from opt_b200.workloads with the recorded
arguments, so only outputs are stored).  The reference itself cannot be executed (no Terra/Lua), so these
vectors pin the ORACLE against regressions; the oracle in turn is pinned against the reference's
hand-written equations, finite differences and the analytic curve-fit answer (tests/test_oracle.py).
Run from the repo root:  python tests/golden/generate.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from opt_b200 import workloads  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

CASES = {
    # name: (energy, solverkind, dims, generator, unknown positions, nIterations, lIterations)
    "image_warping_gn": ("image_warping", "gaussNewtonGPU", [32, 24], lambda: workloads.image_warping(32, 24, dtype=np.float64, mask="blob", seed=9), [0, 1], 4, 40),
    "image_warping_lm": ("image_warping", "LMGPU", [32, 24], lambda: workloads.image_warping(32, 24, dtype=np.float64, mask="random", seed=9), [0, 1], 4, 40),
    "arap_gn": ("arap_mesh_deformation", "gaussNewtonGPU", [9 * 7], lambda: workloads.arap_mesh(9, 7, dtype=np.float64, n_handles=6), [2, 3], 4, 40),
    "poisson_gn": ("poisson_image_editing", "gaussNewtonGPU", [20, 16], lambda: workloads.poisson(20, 16, dtype=np.float64), [0], 2, 60),
    "laplacian_gn": ("laplacian", "gaussNewtonGPU", [16, 12], lambda: workloads.laplacian(16, 12, dtype=np.float64), [0], 2, 40),
    "curve_gn": ("curveFitting", "gaussNewtonGPU", [128, 1], lambda: workloads.curve_fitting(128), [0], 10, 10),
    "sfs_lm": ("shape_from_shading", "LMGPU", [24, 20], lambda: workloads.shape_from_shading(24, 20, dtype=np.float64), [16], 5, 10),
    "rna_lm": ("robust_nonrigid_alignment", "LMGPU", [6 * 5], lambda: workloads.robust_nonrigid(6, 5, dtype=np.float64), [2, 3, 4], 4, 30),
}


def run(name):
    energy, kind, dims, gen, upos, nit, lit = CASES[name]
    p = gen()
    o = Oracle(energy, kind, dims, True)
    o.set_solver_parameter("nIterations", nit); o.set_solver_parameter("lIterations", lit)
    o.init(p)
    costs = [o.cost()]
    while o.step(p):
        costs.append(o.cost())
    out = {"costs": np.array(costs), "nIterations": nit, "lIterations": lit}
    for k, pos in enumerate(upos):
        out[f"unknown{k}"] = p[pos]
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        out = run(name)
        np.savez_compressed(os.path.join(here, name + ".npz"), **out)
        print(name, out["costs"][0], "->", out["costs"][-1], f"({len(out['costs']) - 1} steps)")
