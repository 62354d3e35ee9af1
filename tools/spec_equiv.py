#!/usr/bin/env python
"""Checks that two .t files state the SAME energy: both run through the front end, their generated code is executed on the CPU
(tests/host/spec_eval_host.cpp: the generic engine's device bodies compiled for the host) on identical random inputs built from the
parameter table, and cost / J^T F / diag(J^T J) are compared in fp64.  Used when a spec in specs/ is re-stated:
    python tools/spec_equiv.py specs/optical_flow.t /root/reference/examples/optical_flow/optical_flow.t"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = "/usr/bin/g++"


def build_probe(tmp):
    exe = os.path.join(tmp, "fe_probe")
    subprocess.check_call([CXX, "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "fe_probe.cpp"), os.path.join(ROOT, "opt_b200", "csrc", "frontend.cpp")])
    return exe


def meta(probe, spec):
    out = subprocess.run([probe, "--meta", spec], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError(out.stdout + out.stderr)
    m = {"images": [], "params": [], "graphs": [], "maxdim": 0}
    for line in out.stdout.splitlines():
        t = line.split()
        if not t:
            continue
        if t[0] == "MAXDIM":
            m["maxdim"] = int(t[1])
        elif t[0] == "IMAGE":
            m["images"].append({"type": t[2], "channels": int(t[3]), "index": int(t[4]), "unknown": int(t[5]), "dims": [int(x) for x in t[6:]]})
        elif t[0] == "PARAM":
            m["params"].append({"type": t[2], "index": int(t[3])})
        elif t[0] == "GRAPH":
            m["graphs"].append({"count": int(t[2]), "slots": []})
        elif t[0] == "SLOT":
            m["graphs"][-1]["slots"].append({"index": int(t[2]), "dims": [int(x) for x in t[3:]]})
    return m


def layout(m):   # what must agree between the two files: types, channels, positions, index spaces (names are free)
    return (sorted((i["index"], i["type"], i["channels"], i["unknown"], tuple(i["dims"])) for i in m["images"]),
            sorted((p["index"], p["type"]) for p in m["params"]),
            sorted((g["count"], tuple((s["index"], tuple(s["dims"])) for s in g["slots"])) for g in m["graphs"]))


def single_precision(m):   # a spec that declares a `float` unknown only exists in single precision (o.t:651-656)
    return any(im["unknown"] and im["type"] == "float" for im in m["images"])


def random_params(m, dims, rng, n_edges=37):
    files = {}
    real = np.float32 if single_precision(m) else np.float64
    for im in m["images"]:
        n = int(np.prod([dims[d] for d in im["dims"]]))
        if im["type"] == "uint8":
            a = (rng.random((n, im["channels"])) < 0.8).astype(np.uint8)
        elif im["type"] == "float":
            a = rng.normal(size=(n, im["channels"])).astype(np.float32)
        else:
            a = (rng.normal(size=(n, im["channels"])) * 0.5 + (1.0 if not im["unknown"] else 0.3)).astype(real)
            if not im["unknown"] and im["channels"] == 1:
                a[rng.random(n) < 0.5] = 0.0      # single-channel data arrays are often masks / validity maps tested against 0
        files[im["index"]] = a
    for p in m["params"]:
        files[p["index"]] = np.array([0.3 + rng.random()], np.float32 if p["type"] == "float" else np.float64)
    for g in m["graphs"]:
        files[g["count"]] = np.array([n_edges], np.int32)
        for s in g["slots"]:
            cols = [rng.integers(0, dims[d], n_edges) for d in s["dims"]]
            files[s["index"]] = np.stack(cols, -1).astype(np.int32)
    return files


def evaluate(probe, spec, files, dims, tmp, tag, fp32=False, frontend_obj=None):
    inc = os.path.join(tmp, f"gen_{tag}.inc")
    subprocess.check_call([probe, "--emit", inc, spec], stdout=subprocess.DEVNULL)
    exe = os.path.join(tmp, f"eval_{tag}")
    subprocess.check_call([CXX, "-O1", "-std=c++17", f'-DGEN_SOURCE="{inc}"', "-DHOST_T=" + ("float" if fp32 else "double"), "-o", exe, os.path.join(ROOT, "tests", "host", "spec_eval_host.cpp"),
                           frontend_obj or os.path.join(ROOT, "opt_b200", "csrc", "frontend.cpp")])
    data = os.path.join(tmp, f"data_{tag}"); os.makedirs(data, exist_ok=True)
    for k, a in files.items():
        np.ascontiguousarray(a).tofile(os.path.join(data, f"p{k}.bin"))
    res = os.path.join(tmp, f"out_{tag}.bin")
    r = subprocess.run([exe, spec, data, res] + [str(d) for d in dims], capture_output=True, text=True, timeout=600)
    if r.returncode != 0 or not r.stdout.startswith("COST "):
        raise RuntimeError(r.stdout + r.stderr)
    return float(r.stdout.split()[1]), np.fromfile(res, np.float64)


def equivalent(a, b, dims=None, seed=0, verbose=True, probe=None, frontend_obj=None):
    """probe / frontend_obj: prebuilt tools/fe_probe and frontend.o (the test suite builds them once); built here when absent"""
    with tempfile.TemporaryDirectory() as tmp:
        probe = probe or build_probe(tmp)
        ma, mb = meta(probe, a), meta(probe, b)
        if layout(ma) != layout(mb):
            if verbose:
                print("parameter tables differ:\n ", layout(ma), "\n ", layout(mb))
            return False
        dims = dims or [11, 7, 5][: ma["maxdim"] + 1]
        files = random_params(ma, dims, np.random.default_rng(seed))
        fp32 = single_precision(ma)
        ca, va = evaluate(probe, a, files, dims, tmp, "a", fp32, frontend_obj)
        cb, vb = evaluate(probe, b, files, dims, tmp, "b", fp32, frontend_obj)
        tol = 1e-5 if fp32 else 1e-11
        ok = abs(ca - cb) <= tol * max(1.0, abs(ca)) and va.shape == vb.shape and np.abs(va - vb).max() <= tol * max(1.0, np.abs(va).max())
        if verbose:
            print(f"{'EQUIVALENT' if ok else 'DIFFERENT '}  cost {ca:.12g} vs {cb:.12g}; max |J^T F, diag| difference {np.abs(va - vb).max() if va.shape == vb.shape else 'shape'}  ({os.path.basename(a)})")
        return ok


if __name__ == "__main__":
    sys.exit(0 if equivalent(sys.argv[1], sys.argv[2]) else 1)
