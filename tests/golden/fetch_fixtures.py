#!/usr/bin/env python
"""Copies the input fixtures the parity tests load (byte-identical) from the reference tree into tests/golden/fixtures/.
They are DATA files the reference ships for its examples (public domain / freely redistributable, see
/root/reference/examples/data/copyright.txt, reproduced in fixtures/README.md), not reference source code.
/root/reference does not exist on the GPU box, so the tests read the committed copies.
    python tests/golden/fetch_fixtures.py"""
import os
import shutil

SRC = "/root/reference/examples/data"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures")
FILES = ["cat512_mask.png", "cat512.constraints", "small_armadillo.ply", "small_armadillo.mrk", "poisson0.png", "poisson1.png", "poisson_mask.png",
         "shape_from_shading/default.SFSSolverParameters", "shape_from_shading/default_initialUnknown.imagedump",
         "shape_from_shading/default_maskEdgeMap.imagedump", "shape_from_shading/default_targetDepth.imagedump",
         "shape_from_shading/default_targetIntensity.imagedump"]

if __name__ == "__main__":
    for f in FILES:
        d = os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), d)
        print(f, os.path.getsize(d))
