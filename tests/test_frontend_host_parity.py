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


@pytest.fixture(scope="module")
def fe_probe(tmp_path_factory):
    exe = tmp_path_factory.mktemp("fe") / "fe_probe"
    subprocess.check_call([CXX, "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tools", "fe_probe.cpp"), os.path.join(ROOT, "opt_b200", "csrc", "frontend.cpp")])
    return str(exe)


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


@pytest.mark.parametrize("source", ["specs", "reference"])
@pytest.mark.parametrize("energy", sorted(CASES))
def test_generated_code_matches_oracle_on_the_host(fe_probe, tmp_path, energy, source):
    """source = "specs": this repository's own statement of the energy; "reference": the reference's .t file itself, read
    where it lies (only when the reference tree is mounted) -- the oracle restates exactly that file"""
    from oracle.oracle import Oracle
    dims, gen, upos = CASES[energy]
    host = gen()
    spec = spec_path(energy) if source == "specs" else _reference_spec(energy)
    if source == "reference" and not os.path.exists(spec):
        pytest.skip("reference tree not mounted")
    inc = tmp_path / "generated.inc"
    out = subprocess.run([fe_probe, "--emit", str(inc), spec], capture_output=True, text=True)
    assert out.returncode == 0 and inc.exists(), out.stdout + out.stderr
    exe = tmp_path / "spec_eval_host"
    dp = energy not in FP32
    subprocess.check_call([CXX, "-O1", "-std=c++17", f'-DGEN_SOURCE="{inc}"', "-DHOST_T=" + ("double" if dp else "float"), "-o", str(exe), os.path.join(ROOT, "tests", "host", "spec_eval_host.cpp"),
                           os.path.join(ROOT, "opt_b200", "csrc", "frontend.cpp")])
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
