"""Problem set-ups of the reference's example programs, built from their own input FILES through opt_b200.io -- the part of
examples/*/src/main.cpp + CombinedSolver.h that turns files into the positional parameter table of the .t spec (host numpy here;
workloads.to_device makes the CUDA-tensor form).  Same constants as the reference apps; what they do with mLib/OpenMesh/LodePNG is
done with the minimal readers of include/OptB200IO.h."""
import os

import numpy as np

from . import io

FIXTURES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fixtures")


def image_warping_from_files(mask_png, constraints_file, alpha=1.0, downsample=1, dtype=np.float32, w_fit=100.0, w_reg=0.01):
    """examples/image_warping/src/main.cpp:55-108 + CombinedSolver.h:125-207: Mask = red channel of <name>_mask.png (0 = active),
    UrShape = Offset0 = pixel grid, Angle0 = 0, Constraints = -1 except the marker pixels (moved towards their targets by `alpha`,
    the continuation parameter of preNonlinearSolve) and the image border (pinned to itself); weights passed as square roots.
    Returns (params, (W, H))."""
    m = io.read_png(mask_png)[::downsample, ::downsample, 0].astype(dtype)
    H, W = m.shape
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    ur = np.stack([xs, ys], -1).astype(dtype).reshape(-1, 2).copy()
    rows = [list(r // downsample) for r in io.read_pair_constraints(constraints_file)]
    for y in range(H):
        for x in range(W):
            if y == 0 or x == 0 or y == H - 1 or x == W - 1:
                rows.append([x, y, x, y])
    cons = -np.ones((W * H, 2), dtype)
    for x, y, tx, ty in rows:
        if m[y, x] == 0:
            cons[y * W + x] = [(1.0 - alpha) * x + alpha * tx, (1.0 - alpha) * y + alpha * ty]
    return [ur.copy(), np.zeros(W * H, dtype), ur, cons, m.reshape(-1).copy(), float(np.sqrt(np.float32(w_fit))), float(np.sqrt(np.float32(w_reg)))], (W, H)


def shape_from_shading_from_files(prefix, dtype=np.float32):
    """examples/shape_from_shading/src/SFSSolverInput.h:22-73 + main.cpp:21-28: four .imagedump files and the parameter dump under
    `prefix` (e.g. .../shape_from_shading/default); the unknown starts from <prefix>_initialUnknown, the row / column edge masks are
    the two halves of <prefix>_maskEdgeMap.  Returns (params, (W, H))."""
    X = io.read_imagedump(prefix + "_initialUnknown.imagedump"); D = io.read_imagedump(prefix + "_targetDepth.imagedump")
    Im = io.read_imagedump(prefix + "_targetIntensity.imagedump"); E = io.read_imagedump(prefix + "_maskEdgeMap.imagedump", clamp_infinity=False)
    P = io.read_sfs_parameters(prefix + ".SFSSolverParameters")
    H, W = X.shape
    e = E.reshape(-1)
    eR, eC = e[:W * H].copy(), e[W * H:2 * W * H].copy()
    scal = [P["weightFitting"], P["weightRegularizer"], P["weightShading"], P["fx"], P["fy"], P["ux"], P["uy"]] + [float(v) for v in P["lightingCoefficients"]]
    c = lambda a: np.ascontiguousarray(a.reshape(-1).astype(dtype))
    return [float(v) for v in scal] + [c(X), c(D), c(Im), eR, eC], (W, H)


def poisson_from_files(image0_png, image1_png, mask_png, dtype=np.float32):
    """examples/poisson_image_editing/src/main.cpp:5-67 + CombinedSolver.h:66-90 (BASELINE.json configs[0]): X0 = the base image
    (RGB as floats 0..255, alpha forced to 255), T = the image to insert (same treatment, placed at offset 0 on a black canvas of the
    base image's size), M = 0 where the mask file's red channel is 255 (the blend region: those pixels are solved for) and 255 elsewhere
    (excluded).  The app solves with nIterations = 1, lIterations = 100.  Returns (params, (W, H))."""
    a = io.read_png(image0_png).astype(dtype); b = io.read_png(image1_png).astype(dtype); m = io.read_png(mask_png)[..., 0]
    H, W = a.shape[:2]
    X = a[..., :4].copy(); X[..., 3] = 255
    T = np.zeros_like(X); T[..., 3] = 255
    h, w = min(H, b.shape[0]), min(W, b.shape[1])
    T[:h, :w, :3] = b[:h, :w, :3]
    M = np.zeros((H, W), dtype)
    hm, wm = min(H, m.shape[0]), min(W, m.shape[1])
    M[:hm, :wm] = m[:hm, :wm]                              # imageR32MaskLarge: the mask file on a zero canvas
    M = np.where(M == 255, 0, 255).astype(dtype)
    return [np.ascontiguousarray(X.reshape(-1, 4)), np.ascontiguousarray(T.reshape(-1, 4)), np.ascontiguousarray(M.reshape(-1))], (W, H)


def sqrt3_subdivide(v, faces):
    """One step of Kobbelt's sqrt(3) subdivision of a closed triangle mesh -- what the reference app applies to its input once
    before solving (OpenMesh::Subdivider::Uniform::Sqrt3T, examples/arap_mesh_deformation/src/main.cpp:60-73; the marker file indexes
    the SUBDIVIDED mesh): a new vertex at every face centroid (index N + f), old vertices relaxed towards the mean of their old
    neighbours with weight (4 - 2 cos(2 pi / n)) / 9, every old edge flipped to connect the two adjacent centroids."""
    v = np.asarray(v, np.float64); faces = np.asarray(faces, np.int64)
    N, F = len(v), len(faces)
    nbrs = [set() for _ in range(N)]
    edge_faces = {}
    for f, (a, b, c) in enumerate(faces):
        for x, y in ((a, b), (b, c), (c, a)):
            nbrs[x].add(y); nbrs[y].add(x)
            edge_faces.setdefault((min(x, y), max(x, y)), []).append((f, x, y))
    cent = v[faces].mean(1)
    newv = np.empty((N + F, 3))
    for i in range(N):
        n = len(nbrs[i]); w = (4.0 - 2.0 * np.cos(2.0 * np.pi / n)) / 9.0
        newv[i] = (1.0 - w) * v[i] + w * v[list(nbrs[i])].mean(0)
    newv[N:] = cent
    newf = []
    for (lo, hi), fl in edge_faces.items():
        if len(fl) != 2:
            raise ValueError("sqrt3_subdivide: closed manifold triangle mesh expected")
        (f1, x, y), (f2, _, _) = fl                      # in f1 the edge runs x -> y
        newf.append([x, N + f2, N + f1]); newf.append([y, N + f1, N + f2])
    return newv.astype(np.float32), np.asarray(newf, np.int32)


def arap_from_files(mesh_file, marker_file, alpha=1.0, dtype=np.float32, w_fit=4.0, w_reg=1.0, subdivisions=1):
    """examples/arap_mesh_deformation/src/main.cpp:37-75,:93-96 + CombinedSolver.h:58-62,:90-148,:150-172: the mesh is subdivided
    once (sqrt3), Offset0 = UrShape = its vertices, Angle0 = 0, Constraints = -inf except the marker vertices (moved from their rest
    position towards the marker position by `alpha`), graph = one directed edge per (vertex, neighbour).  Returns (params, N)."""
    v, faces = io.read_mesh(mesh_file)
    for _ in range(subdivisions):
        v, faces = sqrt3_subdivide(v, faces)
    pos, _, idx = io.read_markers(marker_file)
    N = len(v)
    cons = np.full((N, 3), -np.inf, dtype)
    for k, vi in enumerate(idx):
        cons[vi] = (1.0 - alpha) * v[vi] + alpha * pos[k]
    head, tail = io.mesh_neighbour_edges(v, faces)
    c = lambda a: np.ascontiguousarray(a.astype(dtype))
    return [float(np.sqrt(np.float32(w_fit))), float(np.sqrt(np.float32(w_reg))), c(v), np.zeros((N, 3), dtype), c(v), cons, int(len(head)), head, tail], N


def read_tets(ele_file):
    """TetGen .ele file as the reference reads it (robust_nonrigid_alignment/src/main.cpp:17-31): `count 4 0`, then `index a b c d` per line"""
    with open(ele_file) as f:
        tok = f.read().split("#")[0].split()
    n = int(tok[0])
    return np.asarray(tok[3:3 + 5 * n], np.int64).reshape(n, 5)[:, 1:].astype(np.int32)


def vertex_normals(v, faces):
    """what OpenMesh's update_normals() leaves at the vertices (the harness's source of normals): the normalised sum of the unit normals
    of the incident faces"""
    v = np.asarray(v, np.float64); f = np.asarray(faces, np.int64)
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    fn /= np.maximum(np.linalg.norm(fn, axis=1, keepdims=True), 1e-30)
    n = np.zeros_like(v)
    for k in range(3):
        np.add.at(n, f[:, k], fn)
    return (n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)).astype(np.float32)


def rna_correspondences(src, src_normals, tgt, tgt_normals, position_threshold, cos_normal_threshold=0.7, max_k=20):
    """setConstraints of the harness (robust_nonrigid_alignment/src/CombinedSolver.h:261-346) on the host: for every source vertex the
    max_k nearest target vertices in order of distance; stop at the first one farther than position_threshold; take the first whose normal
    agrees (dot > cos_normal_threshold).  Returns (constraints with -inf rows where none was found, constraint normals)."""
    from scipy.spatial import cKDTree
    d, idx = cKDTree(np.asarray(tgt, np.float64)).query(np.asarray(src, np.float64), k=max_k)
    cons = np.full((len(src), 3), -np.inf, np.float32); cn = np.zeros((len(src), 3), np.float32)
    for i in range(len(src)):
        for dist, t in zip(d[i], idx[i]):
            if dist > position_threshold:
                break
            if float(np.dot(tgt_normals[t], src_normals[i])) > cos_normal_threshold:
                cons[i] = tgt[t]; cn[i] = tgt_normals[t]
                break
    return cons, cn


def rna_from_files(source_mesh, target_mesh, tet_file=None, dtype=np.float64, w_fit=10.0, w_reg=64.0):
    """examples/robust_nonrigid_alignment/src/main.cpp:33-63 + CombinedSolver.h:73-118,:142-170,:261-346,:349-424 for ONE target frame at its
    first outer iteration: Offset0 = UrShape = the source vertices, Angle0 = 0, RobustWeights0 = 1, graph = one directed edge per (vertex,
    neighbour) -- neighbours through the tetrahedra when a .ele file is given (sorted, as the harness's std::set walks them), else through
    the surface edges --, Constraints / ConstraintNormals = the correspondences of setConstraints with positionThreshold = 5 x the average
    edge length, weights w_fit = 10, w_reg = 64 (the start of the harness's continuation 64 -> 4) passed as square roots.  The harness
    also displaces 5 % of the targets by noise drawn from std::normal_distribution (implementation-defined sequence): not reproduced here.
    Returns (params, N, info)."""
    v, faces = io.read_mesh(source_mesh)
    tv, tfaces = io.read_mesh(target_mesh)
    N = len(v)
    if tet_file:
        tets = read_tets(tet_file)
        pairs = np.concatenate([np.stack([tets[:, i], tets[:, (i + j) % 4]], 1) for i in range(4) for j in range(1, 4)], 0)
        pairs = np.unique(pairs, axis=0)                         # (head, tail) sorted by head, then tail: std::set order per vertex
        head, tail = pairs[:, 0].astype(np.int32).copy(), pairs[:, 1].astype(np.int32).copy()
    else:
        head, tail = io.mesh_neighbour_edges(v, faces)
    f = np.asarray(faces, np.int64)
    und = np.unique(np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0), axis=1), axis=0)
    avg_edge = float(np.linalg.norm(v[und[:, 0]].astype(np.float64) - v[und[:, 1]], axis=1).mean())
    sn, tn = vertex_normals(v, faces), vertex_normals(tv, tfaces)
    cons, cn = rna_correspondences(v, sn, tv, tn, 5.0 * avg_edge)
    c = lambda a: np.ascontiguousarray(np.asarray(a).astype(dtype))
    params = [float(np.sqrt(np.float32(w_fit))), float(np.sqrt(np.float32(w_reg))), c(v), np.zeros((N, 3), dtype), np.ones((N, 1), dtype), c(v), c(cons), c(cn),
              int(len(head)), head, tail]
    return params, N, {"average_edge_length": avg_edge, "correspondences": int(np.isfinite(cons[:, 0]).sum()), "edges": int(len(head))}


# ------------------------------------------------------------------------------------------------ per-outer-iteration pre-steps on the GPU
# (include/OptB200Examples.h; thin wrappers over torch CUDA tensors -- device memory plumbing only)
def _ptr(t):
    return t.data_ptr() if t is not None else None


def _dp(t):
    import torch
    return 1 if t.dtype == torch.float64 else 0


def image_warping_constraints_gpu(constraints, mask, W, H, rows, alpha):
    """setConstraintImage(alpha) on the device: `constraints` (W*H*2) is rewritten in place from the marker rows (n x 4 int32 CUDA tensor)"""
    from . import capi
    rc = capi.lib().OptB200_ImageWarpingConstraints(_ptr(constraints), _ptr(mask), W, H, _ptr(rows), int(rows.shape[0]), float(alpha), _dp(constraints), None)
    if rc != 0:
        raise capi.OptError(f"OptB200_ImageWarpingConstraints failed ({rc})")


def arap_constraints_gpu(constraints, current, idx, targets, alpha):
    from . import capi
    rc = capi.lib().OptB200_ArapConstraints(_ptr(constraints), _ptr(current), int(current.shape[0]), _ptr(idx), _ptr(targets), int(idx.shape[0]), float(alpha), _dp(constraints), None)
    if rc != 0:
        raise capi.OptError(f"OptB200_ArapConstraints failed ({rc})")


def vertex_normals_gpu(positions, faces):
    import torch
    from . import capi
    out = torch.empty((positions.shape[0], 3), dtype=torch.float32, device=positions.device)
    rc = capi.lib().OptB200_VertexNormals(_ptr(positions), _dp(positions), int(positions.shape[0]), _ptr(faces), int(faces.shape[0]), _ptr(out), None)
    if rc != 0:
        raise capi.OptError(f"OptB200_VertexNormals failed ({rc})")
    return out


class PointGrid:
    """search structure over the target vertices of robust_nonrigid_alignment (replaces the harness's nanoflann k-d tree)"""

    def __init__(self, target_pos, cell_size):
        from . import capi
        self.L = capi.lib(); self.tp = target_pos
        self.h = self.L.OptB200_PointGridCreate(_ptr(target_pos), int(target_pos.shape[0]), float(cell_size), None)
        if not self.h:
            raise capi.OptError("OptB200_PointGridCreate failed")

    def close(self):
        if self.h:
            self.L.OptB200_PointGridFree(self.h); self.h = None

    def __del__(self):
        self.close()

    def correspondences(self, target_normals, source_pos, source_normals, constraints, constraint_normals, robust_weights, prev_constraints,
                        position_threshold, cos_normal_threshold=0.7, max_k=20):
        """setConstraints of the harness: fills constraints / constraint_normals / robust_weights in place, returns the number of changed constraints"""
        from . import capi
        rc = self.L.OptB200_RnaCorrespondences(self.h, _ptr(self.tp), _ptr(target_normals), _ptr(source_pos), _ptr(source_normals), int(source_pos.shape[0]),
                                               float(position_threshold), float(cos_normal_threshold), int(max_k), _ptr(constraints), _ptr(constraint_normals),
                                               _ptr(robust_weights), _ptr(prev_constraints), _dp(source_pos), None)
        if rc < 0:
            raise capi.OptError(f"OptB200_RnaCorrespondences failed ({rc})")
        return rc
