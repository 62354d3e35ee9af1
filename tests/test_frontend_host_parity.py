"""No GPU needed: the code the .t front end GENERATES for every energy the oracle knows is executed on the CPU -- through the
generic engine's own device bodies (geActiveBody / geFillBody), compiled for the host behind shims by
tests/host/spec_eval_host.cpp -- and its cost, J^T F and diag(J^T J) are compared with the CPU oracle in fp64.
This pins interpreter + analysis + code generation + dual-number Jacobians + Exclude / bounds rules without a device;
the GPU tests (test_gpu_generic_energies.py, @frontend cases) then only have to add the launch plumbing."""
import os
import subprocess

import numpy as np
import pytest

from opt_b200 import spec_path, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = "/usr/bin/g++"

CASES = {
    # energy: (dims, host-parameter generator (fp64), positions of the unknown arrays)
    "image_warping": ([24, 18], lambda: workloads.image_warping(24, 18, dtype=np.float64, mask="random", seed=4, perturb=0.3), [0, 1]),
    "arap_mesh_deformation": ([11 * 9], lambda: workloads.arap_mesh(11, 9, dtype=np.float64, n_handles=7, perturb=0.2), [2, 3]),
    "poisson_image_editing": ([21, 17], lambda: workloads.poisson(21, 17, dtype=np.float64), [0]),
    "curveFitting": ([96, 1], lambda: workloads.curve_fitting(96, dtype=np.float64), [0]),
    "shape_from_shading": ([26, 22], lambda: workloads.shape_from_shading(26, 22, dtype=np.float64), [16]),
    "robust_nonrigid_alignment": ([8 * 7], lambda: workloads.robust_nonrigid(8, 7, dtype=np.float64, perturb=0.1), [2, 3, 4]),
    # laplacian.t declares its arrays `float`: single precision on both sides
    "laplacian": ([19, 13], lambda: workloads.laplacian(19, 13, dtype=np.float32), [0]),
}
FP32 = {"laplacian"}
# degenerate sizes: single pixels / rows, images without an interior, a mesh of two vertices, a graph of three samples
TINY = {
    "image_warping": ([5, 1], lambda: workloads.image_warping(5, 1, dtype=np.float64, perturb=0.3), [0, 1]),
    "poisson_image_editing": ([1, 1], lambda: workloads.poisson(1, 1, dtype=np.float64), [0]),
    "laplacian": ([1, 3], lambda: workloads.laplacian(1, 3, dtype=np.float32), [0]),
    "arap_mesh_deformation": ([2], lambda: workloads.arap_mesh(2, 1, dtype=np.float64, n_handles=1, perturb=0.2), [2, 3]),
    "shape_from_shading": ([3, 3], lambda: workloads.shape_from_shading(3, 3, dtype=np.float64), [16]),
    "robust_nonrigid_alignment": ([2], lambda: workloads.robust_nonrigid(2, 1, dtype=np.float64, perturb=0.1), [2, 3, 4]),
    "curveFitting": ([3, 1], lambda: workloads.curve_fitting(3, dtype=np.float64), [0]),
}


@pytest.fixture(scope="module")
def fe_probe(tmp_path_factory, frontend_obj):
    exe = tmp_path_factory.mktemp("fe") / "fe_probe"
    subprocess.check_call([CXX, "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tools", "fe_probe.cpp"), frontend_obj])
    return str(exe)


@pytest.fixture(scope="module")
def frontend_obj(tmp_path_factory):
    """frontend.cpp compiled once for the whole module: every harness below links it"""
    obj = tmp_path_factory.mktemp("feobj") / "frontend.o"
    subprocess.check_call([CXX, "-O1", "-std=c++17", "-c", "-o", str(obj), os.path.join(ROOT, "opt_b200", "csrc", "frontend.cpp")])
    return str(obj)


_HARNESS = {}   # (spec file, precision) -> compiled harness, shared by the cases that use the same generated code


def _harness(fe_probe, frontend_obj, spec, dp, tmp_path_factory):
    key = (os.path.abspath(spec), dp)
    if key not in _HARNESS:
        d = tmp_path_factory.mktemp("harness")
        inc = d / "generated.inc"
        out = subprocess.run([fe_probe, "--emit", str(inc), spec], capture_output=True, text=True)
        assert out.returncode == 0 and inc.exists(), out.stdout + out.stderr
        exe = d / "spec_eval_host"
        subprocess.check_call([CXX, "-O1", "-std=c++17", f'-DGEN_SOURCE="{inc}"', "-DHOST_T=" + ("double" if dp else "float"), "-o", str(exe),
                               os.path.join(ROOT, "tests", "host", "spec_eval_host.cpp"), frontend_obj])
        _HARNESS[key] = str(exe)
    return _HARNESS[key]


def _write_params(host, d):
    for k, v in enumerate(host):
        if isinstance(v, np.ndarray):
            np.ascontiguousarray(v).tofile(os.path.join(d, f"p{k}.bin"))
        elif isinstance(v, (int, np.integer)) and not isinstance(v, bool):
            np.array([v], np.int32).tofile(os.path.join(d, f"p{k}.bin"))
        else:
            np.array([v], np.float32).tofile(os.path.join(d, f"p{k}.bin"))     # Param(..., float, k): the callers pass float* (oracle.ParamPack does the same)


REFERENCE = "/root/reference"


def _reference_spec(energy):
    if energy == "curveFitting":
        return os.path.join(REFERENCE, "tests", "minimal_graph_only", "curveFitting.t")
    if energy == "laplacian":
        return os.path.join(REFERENCE, "tests", "minimal", "laplacian.t")
    return os.path.join(REFERENCE, "examples", energy, energy + ".t")


@pytest.mark.parametrize("source", ["specs", "reference", "specs-tiny"])
@pytest.mark.parametrize("energy", sorted(CASES))
def test_generated_code_matches_oracle_on_the_host(fe_probe, frontend_obj, tmp_path_factory, tmp_path, energy, source):
    """source = "specs": this repository's own statement of the energy; "reference": the reference's .t file itself, read
    where it lies (only when the reference tree is mounted) -- the oracle restates exactly that file"""
    from oracle.oracle import Oracle
    dims, gen, upos = (TINY if source == "specs-tiny" else CASES)[energy]
    host = gen()
    spec = spec_path(energy) if source.startswith("specs") else _reference_spec(energy)
    if source == "reference" and not os.path.exists(spec):
        pytest.skip("reference tree not mounted")
    dp = energy not in FP32
    exe = _harness(fe_probe, frontend_obj, spec, dp, tmp_path_factory)
    data = tmp_path / "params"; data.mkdir()
    _write_params(host, str(data))
    res = tmp_path / "out.bin"
    r = subprocess.run([str(exe), spec, str(data), str(res)] + [str(x) for x in dims], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("COST "), r.stdout + r.stderr
    cost = float(r.stdout.split()[1])
    o = Oracle(energy, "gaussNewtonGPU", dims, dp)
    ocost, g, diag = o.eval_jtf(host)
    both = np.fromfile(str(res), np.float64)          # the harness accumulates and writes in double either way
    assert both.size == 2 * g.size, (both.size, g.size)
    jtf, dg = both[:g.size], both[g.size:]
    tol = 1e-10 if dp else 2e-5
    assert abs(cost - ocost) <= (1e-11 if dp else 1e-5) * max(1.0, abs(ocost)), (cost, ocost)
    assert np.abs(jtf - g).max() <= tol * max(1.0, np.abs(g).max())
    assert np.abs(dg - diag).max() <= tol * max(1.0, np.abs(diag).max())


GRAPH2D_SPEC = """
local W, H = Dim("W", 0), Dim("H", 1)
local X = Unknown("X", opt_float2, { W, H }, 0)
local T = Array("T", opt_float2, { W, H }, 1)
local G = Graph("G", 2, "a", { W, H }, 3, "b", { W, H }, 4)
local w = Param("w", float, 5)
UsePreconditioner(true)
Energy(X(0, 0) - T(0, 0))
Energy(w * (X(G.a) - X(G.b) - 0.5 * T(G.a)))
"""


def test_graph_over_a_2d_index_space_on_the_host(fe_probe, frontend_obj, tmp_path_factory, tmp_path):
    """graph vertices of a 2-D index space are (d0, d1) int32 pairs (the reference's Index struct, o.t:355-387); the generated
    code, run on the CPU, must match a direct numpy statement of the energy"""
    W, H, E = 9, 7, 40
    rng = np.random.default_rng(4)
    X = rng.normal(size=(H, W, 2)); T = rng.normal(size=(H, W, 2))
    a = np.stack([rng.integers(0, W, E), rng.integers(0, H, E)], -1).astype(np.int32)      # (x, y) pairs
    b = np.stack([rng.integers(0, W, E), rng.integers(0, H, E)], -1).astype(np.int32)
    w = 0.7
    spec = tmp_path / "graph2d.t"; spec.write_text(GRAPH2D_SPEC)
    exe = _harness(fe_probe, frontend_obj, str(spec), True, tmp_path_factory)
    data = tmp_path / "params"; data.mkdir()
    _write_params([X.reshape(-1, 2), T.reshape(-1, 2), E, a, b, w], str(data))
    res = tmp_path / "out.bin"
    r = subprocess.run([str(exe), str(spec), str(data), str(res), str(W), str(H)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("COST "), r.stdout + r.stderr
    cost = float(r.stdout.split()[1])
    w32 = float(np.float32(w))
    r0 = X - T
    xa, xb, ta = X[a[:, 1], a[:, 0]], X[b[:, 1], b[:, 0]], T[a[:, 1], a[:, 0]]
    r1 = w32 * (xa - xb - 0.5 * ta)
    ref_cost = 0.5 * (r0 ** 2).sum() + 0.5 * (r1 ** 2).sum()
    g = r0.copy()
    np.add.at(g, (a[:, 1], a[:, 0]), w32 * r1)
    np.add.at(g, (b[:, 1], b[:, 0]), -w32 * r1)
    jtf = np.fromfile(str(res), np.float64)[:W * H * 2].reshape(H, W, 2)
    assert abs(cost - ref_cost) <= 1e-12 * ref_cost
    assert np.abs(jtf - g).max() <= 1e-12 * np.abs(g).max()


@pytest.mark.parametrize("energy", ["volumetric_mesh_deformation", "cotangent_mesh_smoothing", "embedded_mesh_deformation", "optical_flow", "intrinsic_image_decomposition"])
def test_specs_without_an_oracle_energy_equal_the_reference_text(fe_probe, frontend_obj, energy):
    """the five example energies the oracle does not restate: specs/<energy>.t (this repository's own statement) and the reference's
    .t file must be the SAME energy -- both through the front end, generated code run on the CPU on identical random inputs built from
    the parameter table (tools/spec_equiv.py), cost / J^T F / diag(J^T J) equal in fp64.  Needs the reference tree."""
    ref = os.path.join(REFERENCE, "examples", energy, energy + ".t")
    if not os.path.exists(ref):
        pytest.skip("reference tree not mounted")
    import importlib.util
    sp = importlib.util.spec_from_file_location("spec_equiv", os.path.join(ROOT, "tools", "spec_equiv.py"))
    mod = importlib.util.module_from_spec(sp); sp.loader.exec_module(mod)
    assert mod.equivalent(spec_path(energy), ref, verbose=False, probe=fe_probe, frontend_obj=frontend_obj)
