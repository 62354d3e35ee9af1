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


def _dump_params(host, data):
    for k, v in enumerate(host):
        if isinstance(v, np.ndarray):
            np.ascontiguousarray(v).tofile(str(data / f"p{k}.bin"))
        else:
            np.array([v], np.float32).tofile(str(data / f"p{k}.bin"))


@pytest.mark.parametrize("head_in_loop", [0, 1])
@pytest.mark.parametrize("dp", [True, False])
@pytest.mark.parametrize("W,H,tw,th,NT,k", [(40, 31, 32, 8, 96, 5), (45, 37, 16, 12, 64, 7), (26, 22, 32, 24, 256, 3), (70, 19, 8, 4, 11, 9),
                                            (5, 4, 32, 8, 128, 4), (64, 48, 64, 16, 352, 9), (33, 9, 32, 4, 64, 1), (40, 31, 32, 8, 96, 0),
                                            (40, 31, 32, 8, 96, 10), (45, 37, 16, 12, 64, 13), (64, 48, 32, 35, 400, 20), (26, 22, 32, 4, 43, 21), (64, 48, 32, 35, 400, 40)])
def test_sfs_tile_algorithms_match_oracle(tmp_path, W, H, tw, th, NT, k, dp, head_in_loop):
    """sfs_tile.cuh: the fused PCGInit1 + CtC + FinalizeDiagonal pass and the tile-resident PCG loop (the bodies of k_sfsInitTiled and
    k_sfsLoopResident), emulated CTA by CTA on the CPU, against the oracle's LM state after k PCG steps"""
    from oracle.oracle import Oracle
    dt = np.float64 if dp else np.float32
    host = workloads.shape_from_shading(W, H, dtype=dt)
    exe = tmp_path / "sfs_tile_host"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-DHOST_T=" + ("double" if dp else "float"), "-o", str(exe), os.path.join(ROOT, "tests", "host", "sfs_tile_host.cpp")])
    data = tmp_path / "params"; data.mkdir()
    _dump_params(host, data)
    res = tmp_path / "out.bin"
    r = subprocess.run([str(exe), str(data), str(res), str(W), str(H), str(tw), str(th), str(NT), str(k), "1", str(head_in_loop)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("RZ0 "), r.stdout + r.stderr
    tok = r.stdout.split()
    val = {tok[i]: float(tok[i + 1]) for i in range(0, len(tok), 2)}
    # the tile gathers against the per-pixel gathers of sfs_device.cuh: same products, same order
    ref_scale = 1.0
    assert val["JTFERR"] <= (1e-12 if dp else 1e-5) * 1e3 and val["DIAGERR"] <= (1e-12 if dp else 1e-5) * 1e3, r.stdout
    # precompute + cost in one tiled pass (k_sfsPrecomputeTiled): B, flags and row coefficients identical to the two per-pixel passes
    assert val["PREERR"] == 0.0 and abs(val["COST"] - val["COSTREF"]) <= (1e-13 if dp else 1e-6) * max(1.0, abs(val["COSTREF"])), r.stdout
    assert val["HEADERR"] <= (1e-12 if dp else 1e-6) * max(1.0, abs(val["RZ0"])), r.stdout       # the head inside the loop kernel against the stand-alone one
    out = np.fromfile(str(res), np.float64).reshape(8, -1)
    o = Oracle("shape_from_shading", "LMGPU", [W, H], dp)
    ref = workloads.clone_params(host)
    o.set_solver_parameter("nIterations", 1); o.set_solver_parameter("lIterations", 50)
    o.set_max_linear_iterations(0)
    o.init(ref); o.step(ref)
    tol0 = 1e-10 if dp else 3e-5
    for idx, name in ((0, "r"), (1, "CtC"), (2, "preconditioner"), (3, "p")):
        b = o.get_vector(name).astype(np.float64)
        assert np.abs(out[idx] - b).max() <= tol0 * max(1e-30, np.abs(b).max()), (name, float(np.abs(out[idx] - b).max()), float(np.abs(b).max()))
    if k == 0:
        return
    o2 = Oracle("shape_from_shading", "LMGPU", [W, H], dp)
    ref = workloads.clone_params(host)
    o2.set_solver_parameter("nIterations", 1); o2.set_solver_parameter("lIterations", 50)
    o2.set_max_linear_iterations(k)
    o2.init(ref); o2.step(ref)
    r0max = float(np.abs(out[0]).max())
    tol = 1e-8 if dp else 3e-4
    for idx, name in ((4, "delta"), (5, "r"), (6, "p"), (7, "z")):
        b = o2.get_vector(name).astype(np.float64)
        assert not np.any(out[idx] == -777.0), name                     # the epilogue rewrote every pixel
        # r = b - A delta after a reset: a few fp32 ulp of |J^T F|; z = r and p = z + beta p carry the same absolute noise
        floor = (1e-12 if dp else 2e-5) * r0max if name in ("r", "z", "p") else 0.0
        assert np.abs(out[idx] - b).max() <= max(tol * max(1e-30, np.abs(b).max()), floor), (name, k, float(np.abs(out[idx] - b).max()), float(np.abs(b).max()))
    assert abs(val["MODEL"] - val["MODELREF"]) <= (1e-12 if dp else 2e-6) * max(1.0, abs(val["MODELREF"])), r.stdout     # model cost inside the loop kernel
    if k == 40:
        assert val["BROKE"] >= 0 and val["DONE"] < 40        # this case ends through the zeta early-out (:1093-1102), after two residual resets
    assert abs(val["Q"] - o2.get_scalar("q")) <= (1e-8 if dp else 1e-3) * max(1e-30, abs(o2.get_scalar("q")))
