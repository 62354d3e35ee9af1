"""CPU-side checks: the oracle reproduces the committed golden vectors; libOpt.so loads and exports every
symbol include/*.h declares; the spec front end accepts the registered energies and rejects anything
else; the product refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import importlib.util
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden_module():
    spec = importlib.util.spec_from_file_location("golden_generate", os.path.join(ROOT, "tests", "golden", "generate.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["image_warping_gn", "image_warping_lm", "arap_gn", "poisson_gn", "laplacian_gn", "curve_gn", "sfs_lm", "rna_lm"])
def test_oracle_reproduces_golden(built, name):
    g = _golden_module()
    ref = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    out = g.run(name)
    assert len(out["costs"]) == len(ref["costs"])
    assert np.allclose(out["costs"], ref["costs"], rtol=1e-9, atol=1e-12)
    k = 0
    while f"unknown{k}" in ref:
        finite = np.isfinite(ref[f"unknown{k}"])
        assert np.allclose(out[f"unknown{k}"][finite], ref[f"unknown{k}"][finite], rtol=1e-8, atol=1e-10), (name, k)
        k += 1


def _declared_symbols():
    names = []
    for h in ("Opt.h", "OptB200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(Opt(?:B200)?_[A-Za-z]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(built):
    from opt_b200 import capi
    L = capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    bound = {s[0] for s in capi.SYMBOLS}
    for n in declared:
        assert hasattr(L, n), f"{n} declared in include/ but not exported by libOpt.so"
        assert n in bound, f"{n} declared in include/ but not bound in opt_b200/capi.py"
    nm = subprocess.check_output(["nm", "-D", "--defined-only", capi.lib_path()], text=True)
    for n in declared:
        assert re.search(rf"\bT {n}\b", nm), n
    assert b"sm_100a" in L.OptB200_Version()


def test_struct_layout_matches_reference_header():
    from opt_b200 import capi
    # reference Opt.h:10-30: four ints, passed by value
    assert ctypes.sizeof(capi.Opt_InitializationParameters) == 16
    assert [f[0] for f in capi.Opt_InitializationParameters._fields_] == ["doublePrecision", "verbosityLevel", "collectPerKernelTimingInfo", "threadsPerBlock"]


def test_spec_front_end_accepts_registered_and_rejects_others(built, tmp_path):
    from opt_b200 import spec_path
    from opt_b200.solver import spec_energy_name
    for name in ("image_warping", "arap_mesh_deformation", "poisson_image_editing", "laplacian", "curveFitting"):
        ok, energy = spec_energy_name(spec_path(name))
        assert ok and energy == name, (name, energy)
    # comments / whitespace / number spelling do not matter ...
    src = open(spec_path("image_warping")).read()
    p = tmp_path / "reformatted.t"
    p.write_text("-- a comment\n" + src.replace("0.0", ".0").replace("\n", "\n   "))
    assert spec_energy_name(str(p)) == (True, "image_warping")
    # ... but any change of the energy does
    p2 = tmp_path / "changed.t"
    p2.write_text(src.replace("w_regSqrt * (", "2 * w_regSqrt * ("))
    ok, msg = spec_energy_name(str(p2))
    assert not ok and "not one of the energies" in msg
    ok, msg = spec_energy_name(str(tmp_path / "missing.t"))
    assert not ok and "cannot open" in msg


def test_reference_specs_are_registered(built):
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not mounted")
    from opt_b200.solver import spec_energy_name
    for rel, name in (("examples/image_warping/image_warping.t", "image_warping"), ("examples/arap_mesh_deformation/arap_mesh_deformation.t", "arap_mesh_deformation"),
                      ("tests/minimal/laplacian.t", "laplacian"), ("tests/minimal_graph_only/curveFitting.t", "curveFitting")):
        assert spec_energy_name(os.path.join(ref, rel)) == (True, name)


def test_no_cpu_fallback(built):
    """without a CUDA device Opt_NewState fails loudly (NULL + message); nothing is computed on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from opt_b200 import OptSolver, spec_path
    from opt_b200.capi import OptError
    with pytest.raises(OptError):
        OptSolver([8, 8], spec_path("image_warping"))
