"""CPU: the LM accept / reject / trust-region decision of opt_b200/csrc/lm_decide.cuh (run on the DEVICE by k_lmDecide when OPTB_DEVICE_LM=1
keeps the outer iteration off the host) compiled for the host and compared, bit for bit, with a transcription of the reference's lines
(API/src/solverGPUGaussNewton.t:1119-1157) in numpy scalars of the same widths: opt_float for costs and radius, double for the cubic of
the step quality.  Also replays the oracle's own LM trajectory on shape_from_shading through it: the sequence of verdicts must be the
oracle's accept / reject sequence."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_lines(T, new_cost, model_cost, prev_cost, radius, factor, min_rel, f_tol, max_r, min_r):
    """:1105-1157, opt_float = T; returns (verdict, prevCost, radius, factor)"""
    with np.errstate(all="ignore"):
        model_cost_change = T(prev_cost - model_cost)                     # computeModelCostChange :949-954
        cost_change = T(prev_cost - new_cost)
        relative_decrease = T(cost_change / model_cost_change)
        if cost_change >= 0 and relative_decrease > min_rel:
            if cost_change <= T(prev_cost * f_tol):
                return 2, prev_cost, radius, factor
            tmp = 1.0 - np.float64(2.0 * np.float64(relative_decrease) - 1.0) ** 3.0
            radius = T(np.float64(radius) / max(1.0 / 3.0, tmp))
            radius = T(min(np.float64(radius), np.float64(max_r)))
            return 0, new_cost, radius, T(2.0)
        radius = T(radius / factor)
        factor = T(T(2.0) * factor)
        if radius <= min_r:
            return 3, prev_cost, radius, factor
        return 1, prev_cost, radius, factor


def _run(tmp_path, T, cases):
    exe = tmp_path / ("lm_decide_" + T.__name__)
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-ffp-contract=off", "-DHOST_T=" + ("double" if T is np.float64 else "float"), "-o", str(exe),
                           os.path.join(ROOT, "tests", "host", "lm_decide_host.cpp")])
    a = np.ascontiguousarray(np.asarray(cases, T).reshape(-1, 9))
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    a.tofile(str(inp))
    r = subprocess.run([str(exe), str(inp), str(out), str(len(a))], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
    return np.fromfile(str(out), np.float64).reshape(-1, 4)


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_decision_equals_the_reference_lines(tmp_path, T):
    rng = np.random.default_rng(17)
    # solver defaults (solverGPUGaussNewton.t:26-39): min_relative_decrease 1e-3, function_tolerance 1e-6, max radius 1e16, min radius 1e-32
    lim = [T(1e-3), T(1e-6), T(1e16), T(1e-32)]
    cases = []
    for _ in range(4000):
        prev = T(10.0 ** rng.uniform(-3, 6))
        kind = rng.integers(0, 6)
        model = T(prev * rng.uniform(0.0, 1.0))
        if kind == 0:   new = T(prev * rng.uniform(0.0, 1.0))                       # decrease of any quality
        elif kind == 1: new = T(prev * (1.0 + rng.uniform(0.0, 2.0)))               # increase: reject
        elif kind == 2:                                                             # around the function tolerance, predicted about as small
            new = T(prev * (1.0 - rng.uniform(0.0, 3e-6))); model = T(prev - (prev - new) * rng.uniform(0.5, 2.0))
        elif kind == 3: new = T(model + (prev - model) * rng.uniform(-0.01, 0.01))  # step quality around 1 (radius grows 3x)
        elif kind == 4: new = T(prev - (prev - model) * rng.uniform(0.0, 2e-3))     # around min_relative_decrease
        else:           new, model = prev, prev                                     # 0 / 0
        radius = T(10.0 ** rng.uniform(-33, 16)); factor = T(2.0 ** rng.integers(1, 12))
        cases.append([new, model, prev, radius, factor] + lim)
    got = _run(tmp_path, T, cases)
    seen = set()
    for c, g in zip(cases, got):
        v, p, r, f = _reference_lines(T, *[T(x) for x in c])
        seen.add(v)
        assert int(g[0]) == v, (c, g, v)
        assert np.float64(p) == g[1] and np.float64(r) == g[2] and np.float64(f) == g[3], (c, g, (v, p, r, f))
    assert seen == {0, 1, 2, 3}


def test_replays_the_oracle_lm_trajectory(tmp_path, built):
    """shape_from_shading LM on the CPU oracle, one Opt_ProblemStep at a time: feeding the oracle's (cost, model cost) of every iteration
    through lm_decide reproduces its accepted costs and trust-region radii"""
    from opt_b200 import workloads
    from oracle.oracle import Oracle
    W, H = 40, 31
    host = workloads.shape_from_shading(W, H)
    o = Oracle("shape_from_shading", "LMGPU", [W, H])
    o.set_solver_parameter("nIterations", 12); o.set_solver_parameter("lIterations", 10)
    p = workloads.clone_params(host)
    o.init(p)
    T = np.float32
    state = [T(o.cost()), T(1e4), T(2.0)]             # trust_region_radius / radius_decrease_factor defaults (:33-34)
    lim = [T(1e-3), T(1e-6), T(1e16), T(1e-32)]
    n_acc = n_rej = 0
    for _ in range(12):
        more = o.step(p)
        new, model = T(o.get_scalar("newCost")), T(o.get_scalar("modelCost"))
        got = _run(tmp_path, T, [[new, model] + state + lim])[0]
        v = int(got[0])
        state = [T(got[1]), T(got[2]), T(got[3])]
        assert abs(float(state[0]) - o.cost()) <= 1e-6 * abs(o.cost()), (state, o.cost())
        assert abs(float(state[1]) - o.get_scalar("trust_region_radius")) <= 1e-6 * o.get_scalar("trust_region_radius")
        n_acc += v == 0; n_rej += v == 1
        assert (v in (0, 1)) == bool(more)
        if not more:
            break
    assert n_acc >= 3
