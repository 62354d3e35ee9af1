"""libOptIO.so (include/OptB200IO.h) against independent decoders of the same files: numpy.fromfile for .imagedump, a zlib + numpy
PNG decoder, a struct-based PLY parser; plus write -> read round trips and the example set-ups built from the shipped fixtures."""
import os
import struct
import zlib

import numpy as np
import pytest

from opt_b200 import examples, io

F = examples.FIXTURES


def _png_reference_decode(path):
    """straight-line PNG decoder (8-bit, non-interlaced) -- independent of io.cpp"""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, pal = 8, b"", None
    while pos < len(b):
        n, t = struct.unpack(">I4s", b[pos:pos + 8]); d = b[pos + 8:pos + 8 + n]; pos += 12 + n
        if t == b"IHDR":
            w, h, depth, ctype, _, _, inter = struct.unpack(">IIBBBBB", d)
        elif t == b"PLTE":
            pal = np.frombuffer(d, np.uint8).reshape(-1, 3)
        elif t == b"IDAT":
            idat += d
    assert depth == 8 and inter == 0
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * ch)
    out = np.zeros((h, w * ch), np.int32)
    for y in range(h):
        ft, line = raw[y, 0], raw[y, 1:].astype(np.int32)
        up = out[y - 1] if y else np.zeros(w * ch, np.int32)
        cur = np.zeros(w * ch, np.int32)
        for i in range(w * ch):
            a = cur[i - ch] if i >= ch else 0; bb = up[i]; c = up[i - ch] if i >= ch else 0
            if ft == 0: v = line[i]
            elif ft == 1: v = line[i] + a
            elif ft == 2: v = line[i] + bb
            elif ft == 3: v = line[i] + (a + bb) // 2
            else:
                p = a + bb - c; pa, pb, pc = abs(p - a), abs(p - bb), abs(p - c)
                v = line[i] + (a if (pa <= pb and pa <= pc) else (bb if pb <= pc else c))
            cur[i] = v & 255
        out[y] = cur
    px = out.reshape(h, w, ch).astype(np.uint8)
    rgba = np.full((h, w, 4), 255, np.uint8)
    if ctype == 0: rgba[..., :3] = px
    elif ctype == 2: rgba[..., :3] = px
    elif ctype == 3: rgba[..., :3] = pal[px[..., 0]]
    elif ctype == 4: rgba[..., :3] = px[..., :1]; rgba[..., 3] = px[..., 1]
    else: rgba = px
    return rgba


def test_png_reader_matches_independent_decoder(built):
    a = io.read_png(os.path.join(F, "cat512_mask.png"))
    b = _png_reference_decode(os.path.join(F, "cat512_mask.png"))
    assert a.shape == (512, 512, 4) and a.dtype == np.uint8 and np.array_equal(a, b)


@pytest.mark.parametrize("ch", [1, 3, 4])
def test_png_round_trip(built, tmp_path, ch):
    rng = np.random.default_rng(ch)
    img = rng.integers(0, 256, (37, 53, ch) if ch > 1 else (37, 53), dtype=np.uint8)
    p = str(tmp_path / "x.png")
    io.write_png(p, img)
    back = io.read_png(p)
    ref = _png_reference_decode(p)
    assert np.array_equal(back, ref)
    if ch == 1:
        assert np.array_equal(back[..., 0], img) and (back[..., 3] == 255).all()
    else:
        assert np.array_equal(back[..., :ch], img)


def test_imagedump_reader(built, tmp_path):
    p = os.path.join(F, "shape_from_shading", "default_targetDepth.imagedump")
    raw = open(p, "rb").read()
    w, h, c, t = struct.unpack("<4i", raw[:16])
    assert (w, h, c, t) == (640, 480, 1, 0)
    ref = np.frombuffer(raw[16:], np.float32).reshape(h, w)
    a = io.read_imagedump(p, clamp_infinity=False)
    assert a.shape == (480, 640) and np.array_equal(a, ref, equal_nan=True)
    b = io.read_imagedump(p)            # the loader's clamp: -inf -> -10000, +inf -> FLT_MAX (SimpleBuffer.cpp:29-40)
    assert np.isfinite(b).all() and np.array_equal(b[np.isfinite(ref)], ref[np.isfinite(ref)]) and (b[np.isneginf(ref)] == -10000.0).all()
    e = io.read_imagedump(os.path.join(F, "shape_from_shading", "default_maskEdgeMap.imagedump"))
    assert e.dtype == np.uint8 and e.shape == (960, 640)          # row and column maps stuck together
    q = str(tmp_path / "y.imagedump")
    x = np.random.default_rng(0).normal(size=(7, 5, 2)).astype(np.float32)
    io.write_imagedump(q, x)
    assert np.array_equal(io.read_imagedump(q), x)
    with pytest.raises(io.IOError_):
        io.read_imagedump(str(tmp_path / "missing.imagedump"))


def test_sfs_parameter_file(built):
    P = io.read_sfs_parameters(os.path.join(F, "shape_from_shading", "default.SFSSolverParameters"))
    raw = np.frombuffer(open(os.path.join(F, "shape_from_shading", "default.SFSSolverParameters"), "rb").read()[:144], np.float32)
    assert P["weightFitting"] == raw[0] == 100.0 and P["weightRegularizer"] == 100.0 and P["weightShading"] == 1.0
    assert abs(P["fx"] - 574.05) < 0.01 and (P["ux"], P["uy"]) == (320.0, 240.0)
    assert np.array_equal(P["deltaTransform"], np.eye(4, dtype=np.float32)) and np.array_equal(P["lightingCoefficients"], raw[27:36])
    assert P["unused"] == [1, 10, 8]


def test_mesh_readers(built, tmp_path):
    p = os.path.join(F, "small_armadillo.ply")
    v, f = io.read_mesh(p)
    raw = open(p, "rb").read()
    body = raw[raw.index(b"end_header\n") + len(b"end_header\n"):]
    vr = np.frombuffer(body[:130 * 12], np.float32).reshape(130, 3)
    fr = np.array([struct.unpack_from("<B3i", body, 130 * 12 + 13 * k)[1:] for k in range(256)], np.int32)
    assert np.array_equal(v, vr) and np.array_equal(f, fr)
    q = str(tmp_path / "a.ply")
    io.write_ply(q, v, f)                                   # ascii
    v2, f2 = io.read_mesh(q)
    assert np.allclose(v2, v, rtol=1e-7, atol=0) and np.array_equal(f2, f)
    obj = str(tmp_path / "a.obj")
    with open(obj, "w") as fh:
        fh.write("# comment\n" + "".join(f"v {x:.9g} {y:.9g} {z:.9g}\n" for x, y, z in v) + "vn 0 0 1\n" + "".join(f"f {a + 1}//1 {b + 1}//1 {c + 1}//1\n" for a, b, c in f))
    v3, f3 = io.read_mesh(obj)
    assert np.allclose(v3, v, rtol=1e-7, atol=0) and np.array_equal(f3, f)
    off = str(tmp_path / "a.off")
    with open(off, "w") as fh:
        fh.write(f"OFF\n{len(v)} {len(f)} 0\n" + "".join(f"{x:.9g} {y:.9g} {z:.9g}\n" for x, y, z in v) + "".join(f"3 {a} {b} {c}\n" for a, b, c in f))
    v4, f4 = io.read_mesh(off)
    assert np.allclose(v4, v, rtol=1e-7, atol=0) and np.array_equal(f4, f)
    head, tail = io.mesh_neighbour_edges(v, f)
    und = {(min(a, b), max(a, b)) for tri in f for a, b in ((tri[0], tri[1]), (tri[1], tri[2]), (tri[2], tri[0]))}
    assert len(head) == 2 * len(und) == 2 * 384 and set(zip(head.tolist(), tail.tolist())) == {(a, b) for a, b in und} | {(b, a) for a, b in und}
    assert (np.diff(head) >= 0).all()                       # grouped by head vertex like createGraphFromNeighborLists
    with pytest.raises(io.IOError_):
        io.read_mesh(str(tmp_path / "a.stl"))


def test_side_files(built):
    pos, rad, idx = io.read_markers(os.path.join(F, "small_armadillo.mrk"))
    assert idx.tolist() == [68, 358, 7] and np.allclose(pos[2], [0.0, 50.0, -50.0]) and np.allclose(rad, 0.0224524)
    rows = io.read_pair_constraints(os.path.join(F, "cat512.constraints"))
    assert rows.shape == (9, 4) and rows[0].tolist() == [30, 132, 59, 44] and rows[-1].tolist() == [92, 192, 84, 192]


def test_example_setups_from_fixtures(built):
    """the three reference apps' inputs, built from their own files, are well-formed problems for the oracle"""
    from oracle.oracle import Oracle
    p, (W, H) = examples.image_warping_from_files(os.path.join(F, "cat512_mask.png"), os.path.join(F, "cat512.constraints"))
    assert (W, H) == (512, 512) and int((p[4] == 0).sum()) == 101406 and int((p[3][:, 0] >= 0).sum()) == 9
    c0 = Oracle("image_warping", "gaussNewtonGPU", [W, H]).eval_jtf(p)[0]
    assert np.isfinite(c0) and c0 > 0
    p, (W, H) = examples.shape_from_shading_from_files(os.path.join(F, "shape_from_shading", "default"))
    assert (W, H) == (640, 480) and p[19].dtype == np.uint8 and int((p[17] > 0).sum()) == 192162
    c0 = Oracle("shape_from_shading", "LMGPU", [W, H]).eval_jtf(p)[0]
    assert np.isfinite(c0) and c0 > 0
    p, N = examples.arap_from_files(os.path.join(F, "small_armadillo.ply"), os.path.join(F, "small_armadillo.mrk"))
    assert N == 386 and p[6] == 2304 and int(np.isfinite(p[5][:, 0]).sum()) == 3       # sqrt3-subdivided once: 130 + 256 vertices
    c0 = Oracle("arap_mesh_deformation", "gaussNewtonGPU", [N]).eval_jtf(p)[0]
    assert np.isfinite(c0) and c0 > 0
