"""Runs the reference's own test programs (built UNCHANGED from /root/reference/tests/*/main.cpp by
tests/build_reference_tests.sh, linked against include/Opt.h + libOpt.so) on the GPU box: SURVEY.md 2 #24
"must link+run unchanged".  They read their specs by relative name, so each runs in a scratch directory
holding this repository's equivalent spec under the expected file name."""
import os
import re
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "_refbin")


def _run(name, spec, tmp_path, timeout=300):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    shutil.copy(os.path.join(ROOT, "specs", spec), tmp_path / spec)
    env = dict(os.environ); env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "opt_b200", "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return r.stdout


def test_minimal(built, tmp_path):
    out = _run("minimal", "laplacian.t", tmp_path)
    assert "cost" in out.lower()
    assert any(f.suffix == ".png" for f in tmp_path.iterdir())      # the program writes its result images


def test_minimal_graph_only(built, tmp_path):
    out = _run("minimal_graph_only", "curveFitting.t", tmp_path)
    m = re.search(r"Result\s+([-\d.e+]+),\s*([-\d.e+]+)", out)
    assert m, out[-500:]
    assert abs(float(m.group(1)) - 100.0) < 1e-3 and abs(float(m.group(2)) - 102.0) < 1e-3     # printed with 6 significant digits


def test_create_delete_cycle(built, tmp_path):
    _run("create_delete_cycle", "laplacian.t", tmp_path, timeout=600)
