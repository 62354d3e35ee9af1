"""No GPU needed: the per-vertex math of the hand-written robust_nonrigid_alignment kernel set (opt_b200/csrc/rna_device.cuh --
the functions energy_rna.cu's kernels call) runs on the CPU (tests/host/rna_host.cpp) and is compared with the oracle:
cost, J^T F, diag(J^T J), J^T J x, and the model cost through  model(x) = cost + x.J^T F + 1/2 x.J^T J x  (no Exclude here,
so the identity is exact)."""
import os
import subprocess

import numpy as np
import pytest

from opt_b200 import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dp", [True, False])
@pytest.mark.parametrize("nx,ny", [(8, 7), (13, 9), (2, 1), (1, 1), (3, 2)])
def test_rna_per_vertex_math_matches_oracle(tmp_path, nx, ny, dp):
    from oracle.oracle import Oracle
    dt = np.float64 if dp else np.float32
    host = workloads.robust_nonrigid(nx, ny, dtype=dt, perturb=0.1)
    N = nx * ny
    exe = tmp_path / "rna_host"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-DHOST_T=" + ("double" if dp else "float"), "-o", str(exe), os.path.join(ROOT, "tests", "host", "rna_host.cpp")])
    data = tmp_path / "params"; data.mkdir()
    for k, v in enumerate(host):
        if isinstance(v, np.ndarray):
            np.ascontiguousarray(v).tofile(str(data / f"p{k}.bin"))
        elif isinstance(v, (int, np.integer)):
            np.array([v], np.int32).tofile(str(data / f"p{k}.bin"))
        else:
            np.array([v], np.float32).tofile(str(data / f"p{k}.bin"))
    rng = np.random.default_rng(3)
    x = rng.normal(size=7 * N).astype(dt)
    x.tofile(str(data / "x.bin"))
    res = tmp_path / "out.bin"
    r = subprocess.run([str(exe), str(data), str(res), str(N)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("COST "), r.stdout + r.stderr
    tok = r.stdout.split()
    cost, model = float(tok[1]), float(tok[3])
    o = Oracle("robust_nonrigid_alignment", "LMGPU", [N], dp)
    ocost, g, diag = o.eval_jtf(host)
    ref = o.apply_jtj(host, x)
    out = np.fromfile(str(res), np.float64).reshape(3, -1)
    tol = 1e-10 if dp else 3e-5
    assert abs(cost - ocost) <= (1e-11 if dp else 2e-5) * max(1.0, abs(ocost)), (cost, ocost)
    assert np.abs(out[0] - g).max() <= tol * max(1.0, np.abs(g).max())
    assert np.abs(out[1] - diag).max() <= tol * max(1.0, np.abs(diag).max())
    assert np.abs(out[2] - ref).max() <= tol * max(1.0, np.abs(ref).max())
    xd = x.astype(np.float64)
    quad = ocost + float(xd @ g.astype(np.float64)) + 0.5 * float(xd @ ref.astype(np.float64))
    assert abs(model - quad) <= (1e-9 if dp else 2e-4) * max(1.0, abs(quad)), (model, quad)
