#!/usr/bin/env python
"""Diagnostic matrix for shape_from_shading LM on one GPU: a few small solves under every implementation toggle, printed next to the
oracle's numbers (run once per library build: OPTB_LIBOPT selects it).  Not a test; tools/gpu1_sfs_diag.sh drives it."""
import contextlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opt_b200 import OptSolver, spec_path, workloads   # noqa: E402
from oracle.oracle import Oracle                        # noqa: E402


@contextlib.contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update(kw)
    try:
        yield
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)


def costs(W, H, dp, nit, lit, e):
    host = workloads.shape_from_shading(W, H, dtype=np.float64 if dp else np.float32)
    dev = workloads.to_device(host)
    with env(**e):
        s = OptSolver([W, H], spec_path("shape_from_shading"), "LMGPU", double_precision=dp)
    s.set_solver_parameters({"nIterations": nit, "lIterations": lit})
    s.init(dev)
    out = [s.current_cost()]
    while s.step(dev):
        out.append(s.current_cost())
    s.close()
    return out


def oracle_costs(W, H, dp, nit, lit):
    host = workloads.shape_from_shading(W, H, dtype=np.float64 if dp else np.float32)
    o = Oracle("shape_from_shading", "LMGPU", [W, H], dp)
    o.set_solver_parameter("nIterations", nit); o.set_solver_parameter("lIterations", lit)
    o.init(host)
    out = [o.cost()]
    while o.step(host):
        out.append(o.cost())
    return out


def k_steps_err(W, H, k, e):
    host = workloads.shape_from_shading(W, H)
    dev = workloads.to_device(workloads.clone_params(host))
    with env(**e):
        s = OptSolver([W, H], spec_path("shape_from_shading"), "LMGPU")
    s.set_solver_parameter("nIterations", 1); s.set_solver_parameter("lIterations", 50); s.set_max_linear_iterations(k)
    s.init(dev); s.step(dev)
    d = s.read_vector("delta"); s.close()
    o = Oracle("shape_from_shading", "LMGPU", [W, H])
    oh = workloads.clone_params(host)
    o.set_solver_parameter("nIterations", 1); o.set_solver_parameter("lIterations", 50); o.set_max_linear_iterations(k)
    o.init(oh); o.step(oh)
    b = o.get_vector("delta")
    return float(np.abs(d[:b.size] - b).max() / max(1e-30, np.abs(b).max()))


CONFIGS = [("default", {}), ("generic", {"OPTB_GENERIC": "1"}), ("no_head_in_loop", {"OPTB_SFS_NO_HEAD_IN_LOOP": "1"}), ("two_launch_precompute", {"OPTB_SFS_NO_FUSED_PRECOMPUTE": "1"}),
           ("no_resident", {"OPTB_SFS_NO_RESIDENT": "1"}), ("no_model", {"OPTB_SFS_NO_FUSED_MODEL": "1"}), ("no_update", {"OPTB_SFS_NO_FUSED_UPDATE": "1"}),
           ("per_step", {"OPTB_PERSISTENT": "0", "OPTB_SFS_NO_FUSED_INIT": "1", "OPTB_SFS_NO_FUSED_COST": "1"}),
           ("no_iter_graph", {"OPTB_NO_ITER_GRAPH": "1"}), ("no_graph", {"OPTB_NO_GRAPH": "1"}), ("nt1024", {"OPTB_SFS_LOOP_NT": "1024"}),
           ("no_fused_init", {"OPTB_SFS_NO_FUSED_INIT": "1"})]


def main():
    which = sys.argv[1:] or [c[0] for c in CONFIGS]
    print(json.dumps({"oracle": True, "fp64_24x20_l10": oracle_costs(24, 20, True, 5, 10), "fp32_26x22_l25": oracle_costs(26, 22, False, 4, 25)[-1],
                      "fp64_26x22_l25": oracle_costs(26, 22, True, 4, 25)[-1]}), flush=True)
    for name, e in CONFIGS:
        if name not in which:
            continue
        try:
            rec = {"config": name, "fp64_24x20_l10": costs(24, 20, True, 5, 10, e), "fp32_26x22_l25": costs(26, 22, False, 4, 25, e)[-1],
                   "fp64_26x22_l25": costs(26, 22, True, 4, 25, e)[-1], "k3_64x48_delta_err": k_steps_err(64, 48, 3, e), "k12_64x48_delta_err": k_steps_err(64, 48, 12, e),
                   "twice_200x150": [costs(200, 150, False, 5, 10, e)[-1] for _ in range(2)]}
        except Exception as ex:   # noqa: BLE001
            rec = {"config": name, "error": repr(ex)}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
