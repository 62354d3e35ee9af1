"""No GPU needed: the per-pixel math of the hand-written shape_from_shading kernel set (opt_b200/csrc/sfs_device.cuh --
the functions energy_sfs.cu's kernels call) runs on the CPU (tests/host/sfs_host.cpp) and is compared with the oracle:
cost, J^T F, diag(J^T J), J^T J x, and the model cost through the identity  model(x) = cost + x.J^T F + 1/2 x.J^T J x."""
import os
import subprocess

import numpy as np
import pytest

from opt_b200 import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dp", [True, False])
@pytest.mark.parametrize("W,H,seed", [(26, 22, 0), (40, 31, 1), (3, 3, 0), (2, 5, 0), (1, 1, 0), (7, 3, 2)])
def test_sfs_per_pixel_math_matches_oracle(tmp_path, W, H, seed, dp):
    from oracle.oracle import Oracle
    dt = np.float64 if dp else np.float32
    try:
        host = workloads.shape_from_shading(W, H, dtype=dt, seed=seed)
    except TypeError:
        host = workloads.shape_from_shading(W, H, dtype=dt)
    exe = tmp_path / "sfs_host"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-DHOST_T=" + ("double" if dp else "float"), "-o", str(exe), os.path.join(ROOT, "tests", "host", "sfs_host.cpp")])
    data = tmp_path / "params"; data.mkdir()
    for k, v in enumerate(host):
        if isinstance(v, np.ndarray):
            np.ascontiguousarray(v).tofile(str(data / f"p{k}.bin"))
        else:
            np.array([v], np.float32).tofile(str(data / f"p{k}.bin"))
    rng = np.random.default_rng(3)
    x = rng.normal(size=W * H).astype(dt)
    x.tofile(str(data / "x.bin"))
    res = tmp_path / "out.bin"
    r = subprocess.run([str(exe), str(data), str(res), str(W), str(H)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("COST "), r.stdout + r.stderr
    tok = r.stdout.split()
    cost, model, n_inc, n_flat = float(tok[1]), float(tok[3]), int(tok[5]), int(tok[7])
    if W * H > 100:
        assert n_inc > 0.5 * W * H and n_flat > 0      # the workload exercises every term
    o = Oracle("shape_from_shading", "LMGPU", [W, H], dp)
    ocost, g, diag = o.eval_jtf(host)
    ref = o.apply_jtj(host, x)
    out = np.fromfile(str(res), np.float64).reshape(3, -1)
    tol = 1e-10 if dp else 3e-5
    assert abs(cost - ocost) <= (1e-11 if dp else 2e-5) * max(1.0, abs(ocost)), (cost, ocost)
    assert np.abs(out[0] - g).max() <= tol * max(1.0, np.abs(g).max())
    assert np.abs(out[1] - diag).max() <= tol * max(1.0, np.abs(diag).max())
    assert np.abs(out[2] - ref).max() <= tol * max(1.0, np.abs(ref).max())
    # model cost: 1/2 |F + J x|^2 over the counted rows.  Rows centred at excluded pixels feed J^T F / J^T J but not the
    # cost, so the quadratic identity only holds where no such row is non-zero; check it against the oracle's pieces then.
    xa = np.where(np.asarray(host[17]).reshape(-1) > 0, x, 0).astype(np.float64)
    quad = ocost + float(xa @ g.astype(np.float64)) + 0.5 * float(xa @ ref.astype(np.float64))
    if abs(model - quad) > (1e-9 if dp else 1e-3) * max(1.0, abs(quad)):
        # legitimate only when rows centred at excluded pixels are non-zero: then model <= quad (it drops non-negative terms)
        assert model <= quad * (1 + 1e-9) + 1e-9
