"""ctypes binding of libOptIO.so (include/OptB200IO.h): the file formats of the reference's examples and fixtures as numpy arrays."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class IOError_(RuntimeError):
    pass


class _Image(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("channels", ctypes.c_int), ("type", ctypes.c_int), ("data", ctypes.c_void_p)]


class _Mesh(ctypes.Structure):
    _fields_ = [("nVertices", ctypes.c_int), ("nFaces", ctypes.c_int), ("vertices", ctypes.POINTER(ctypes.c_float)),
                ("faceStart", ctypes.POINTER(ctypes.c_int)), ("faceIndices", ctypes.POINTER(ctypes.c_int))]


def lib():
    global _LIB
    if _LIB is None:
        p = os.path.join(_HERE, "lib", "libOptIO.so")
        if not os.path.exists(p):
            raise IOError_(f"{p} is missing: run __graft_entry__.build()")
        L = ctypes.CDLL(p)
        L.OptIO_LastError.restype = ctypes.c_char_p
        for name in ("OptIO_ReadImageDump", "OptIO_WriteImageDump", "OptIO_ReadPNG", "OptIO_WritePNG", "OptIO_ReadSFSParameters", "OptIO_ReadMesh",
                     "OptIO_WritePLY", "OptIO_MeshNeighbourEdges", "OptIO_ReadMarkers", "OptIO_ReadPairConstraints"):
            getattr(L, name).restype = ctypes.c_int
        L.OptIO_FreeInts.argtypes = [ctypes.c_void_p]
        _LIB = L
    return _LIB


def _check(rc, what):
    if rc < 0:
        raise IOError_(f"{what}: {lib().OptIO_LastError().decode()}")
    return rc


def _image_to_numpy(img):
    n = img.width * img.height * img.channels
    ct = ctypes.c_float if img.type == 0 else ctypes.c_ubyte
    a = np.ctypeslib.as_array(ctypes.cast(img.data, ctypes.POINTER(ct)), shape=(n,)).copy()
    lib().OptIO_FreeImage(ctypes.byref(img))
    return a.reshape(img.height, img.width, img.channels) if img.channels > 1 else a.reshape(img.height, img.width)


def read_imagedump(path, clamp_infinity=True):
    """(H, W[, C]) float32 or uint8 array; clamp_infinity as the reference's loader (SimpleBuffer.cpp:29-40)"""
    img = _Image()
    _check(lib().OptIO_ReadImageDump(os.fsencode(path), int(bool(clamp_infinity)), ctypes.byref(img)), path)
    return _image_to_numpy(img)


def _as_image(a):
    a = np.ascontiguousarray(a)
    if a.dtype not in (np.float32, np.uint8):
        raise TypeError("float32 or uint8 expected")
    h, w = a.shape[:2]
    c = a.shape[2] if a.ndim == 3 else 1
    return _Image(w, h, c, 0 if a.dtype == np.float32 else 1, a.ctypes.data), a


def write_imagedump(path, a):
    img, keep = _as_image(a)
    _check(lib().OptIO_WriteImageDump(os.fsencode(path), ctypes.byref(img)), path)


def read_png(path):
    """(H, W, 4) uint8 RGBA, like LodePNG::load -> ColorImageR8G8B8A8"""
    img = _Image()
    _check(lib().OptIO_ReadPNG(os.fsencode(path), ctypes.byref(img)), path)
    return _image_to_numpy(img)


def write_png(path, a):
    img, keep = _as_image(np.asarray(a, np.uint8))
    _check(lib().OptIO_WritePNG(os.fsencode(path), ctypes.byref(img)), path)


def read_sfs_parameters(path):
    """dict of the fields of TerraSolverParameters (examples/shape_from_shading/src/TerraSolverParameters.h:7-45)"""
    out = (ctypes.c_float * 40)()
    _check(lib().OptIO_ReadSFSParameters(os.fsencode(path), out), path)
    f = np.array(out[:36], np.float32)
    names = ["weightFitting", "weightRegularizer", "weightPrior", "weightShading", "weightShadingStart", "weightShadingIncrement", "weightBoundary", "fx", "fy", "ux", "uy"]
    d = {n: float(f[i]) for i, n in enumerate(names)}
    d["deltaTransform"] = f[11:27].reshape(4, 4).copy(); d["lightingCoefficients"] = f[27:36].copy(); d["unused"] = [int(out[36 + i]) for i in range(3)]
    return d


def read_mesh(path):
    """(vertices (N,3) float32, faces: list of int arrays -- or an (F,k) int32 array when every face has k vertices)"""
    m = _Mesh()
    _check(lib().OptIO_ReadMesh(os.fsencode(path), ctypes.byref(m)), path)
    v = np.ctypeslib.as_array(m.vertices, shape=(max(m.nVertices, 1) * 3,))[:m.nVertices * 3].copy().reshape(-1, 3)
    fs = np.ctypeslib.as_array(m.faceStart, shape=(m.nFaces + 1,)).copy()
    fi = np.ctypeslib.as_array(m.faceIndices, shape=(max(int(fs[-1]), 1),))[:int(fs[-1])].copy()
    lib().OptIO_FreeMesh(ctypes.byref(m))
    sizes = np.diff(fs)
    faces = fi.reshape(-1, int(sizes[0])).astype(np.int32) if len(sizes) and (sizes == sizes[0]).all() else [fi[fs[k]:fs[k + 1]] for k in range(len(sizes))]
    return v, faces


def _as_mesh(v, faces):
    v = np.ascontiguousarray(v, np.float32)
    if isinstance(faces, np.ndarray):
        fs = (np.arange(faces.shape[0] + 1) * faces.shape[1]).astype(np.int32); fi = np.ascontiguousarray(faces.reshape(-1), np.int32)
    else:
        fs = np.concatenate([[0], np.cumsum([len(f) for f in faces])]).astype(np.int32); fi = np.ascontiguousarray(np.concatenate(faces), np.int32) if len(faces) else np.zeros(1, np.int32)
    m = _Mesh(len(v), len(fs) - 1, v.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), fs.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), fi.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return m, (v, fs, fi)


def write_ply(path, v, faces):
    m, keep = _as_mesh(v, faces)
    _check(lib().OptIO_WritePLY(os.fsencode(path), ctypes.byref(m)), path)


def mesh_neighbour_edges(v, faces):
    """directed edge lists (head, tail) int32: one edge per (vertex, neighbour), grouped by head vertex (OptGraph.h:64-75)"""
    m, keep = _as_mesh(v, faces)
    h = ctypes.POINTER(ctypes.c_int)(); t = ctypes.POINTER(ctypes.c_int)(); n = ctypes.c_int(0)
    _check(lib().OptIO_MeshNeighbourEdges(ctypes.byref(m), ctypes.byref(h), ctypes.byref(t), ctypes.byref(n)), "mesh_neighbour_edges")
    head = np.ctypeslib.as_array(h, shape=(max(n.value, 1),))[:n.value].copy(); tail = np.ctypeslib.as_array(t, shape=(max(n.value, 1),))[:n.value].copy()
    lib().OptIO_FreeInts(h); lib().OptIO_FreeInts(t)
    return head.astype(np.int32), tail.astype(np.int32)


def read_markers(path):
    """(positions (M,3) float32, radii (M,), vertex indices (M,) int32) of a .mrk file (LandMarkSet.h:16-49)"""
    n = _check(lib().OptIO_ReadMarkers(os.fsencode(path), 0, None, None, None), path)
    pos = np.zeros((n, 3), np.float32); rad = np.zeros(n, np.float32); idx = np.zeros(n, np.int32)
    _check(lib().OptIO_ReadMarkers(os.fsencode(path), n, pos.ctypes.data_as(ctypes.c_void_p), rad.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p)), path)
    return pos, rad, idx


def read_pair_constraints(path):
    """(M,4) int32 rows x y targetX targetY of a .constraints file (image_warping/src/main.cpp:4-27)"""
    n = _check(lib().OptIO_ReadPairConstraints(os.fsencode(path), 0, None), path)
    rows = np.zeros((n, 4), np.int32)
    _check(lib().OptIO_ReadPairConstraints(os.fsencode(path), n, rows.ctypes.data_as(ctypes.c_void_p)), path)
    return rows
