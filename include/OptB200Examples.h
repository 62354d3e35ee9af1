/* OptB200Examples.h -- the example programs' per-outer-iteration pre-steps on the GPU.
 *
 * Once the inner solve takes a millisecond, what the reference's example harnesses do on the CPU between solves dominates end to
 * end: they rebuild a constraint image / target array in a host loop and upload it before every outer iteration
 * (examples/image_warping/src/CombinedSolver.h:181-207 setConstraintImage, examples/arap_mesh_deformation/src/CombinedSolver.h:90-108
 * setConstraints), and robust_nonrigid_alignment searches a k-d tree per vertex on one core and uploads targets and normals
 * (examples/robust_nonrigid_alignment/src/CombinedSolver.h:261-346 setConstraints: nanoflann kNN, normal test, change count).
 * The same steps here read and write DEVICE buffers only -- the arrays that are bound into the Opt problem parameter table --
 * so an outer loop of solves never leaves the device.  C ABI, plain pointers; `doublePrecision` selects float/double for the
 * opt_float arrays, exactly like Opt_InitializationParameters.doublePrecision; every call is asynchronous on `stream` (a cudaStream_t,
 * NULL = legacy default stream) except where a count is returned to the host.  All return 0 on success (negative: bad argument).
 */
#ifndef OPTB200_EXAMPLES_H
#define OPTB200_EXAMPLES_H
#ifdef __cplusplus
extern "C" {
#endif

/* image_warping: Constraints(x,y) = (-1,-1) everywhere, then for each marker row k = (x, y, tx, ty) with Mask(x,y) == 0:
 * Constraints(x,y) = (1-alpha)(x,y) + alpha(tx,ty).  dRows: n x 4 int32 on the device (markers AND the pinned border pixels,
 * image_warping/src/main.cpp:98-108).  dConstraints: W*H*2 opt_float, dMask: W*H opt_float. */
int OptB200_ImageWarpingConstraints(void* dConstraints, const void* dMask, int W, int H, const int* dRows, int n, float alpha, int doublePrecision, void* stream);

/* arap_mesh_deformation: Constraints(v) = -inf for all v, then for marker k: Constraints(idx[k]) = (1-alpha) Current(idx[k]) + alpha target[k]
 * (the harness interpolates from the CURRENT deformed position m_result.point).  dCurrent / dConstraints: N*3 opt_float;
 * dIdx: n int32; dTargets: n*3 float. */
int OptB200_ArapConstraints(void* dConstraints, const void* dCurrent, int N, const int* dIdx, const float* dTargets, int n, float alpha, int doublePrecision, void* stream);

/* area-free vertex normals as OpenMesh::update_normals computes them (normalised sum of the unit normals of the incident faces);
 * dPositions: N*3 (opt_float when doublePrecision applies, see below), dFaces: nFaces*3 int32, dNormals: N*3 float out. */
int OptB200_VertexNormals(const void* dPositions, int positionsAreDouble, int N, const int* dFaces, int nFaces, float* dNormals, void* stream);

/* robust_nonrigid_alignment, setConstraints: an opaque search structure over the target vertices (uniform grid, cell size = the
 * position threshold), then per source vertex: its maxK nearest target vertices within positionThreshold in order of distance, the
 * first whose normal satisfies dot(nTarget, nSource) > cosNormalThreshold becomes the correspondence (Constraints = its position,
 * ConstraintNormals = its normal), otherwise Constraints = -inf.  Where the constraint differs from dPrevConstraints the robust weight is
 * reset to 1 and the change is counted (the harness stops early when fewer than 6 change).  Returns the number of changed
 * constraints (>= 0; this one value is read back), negative on error.
 * dSourcePos: N*3 opt_float (the Offset unknown), dSourceNormals: N*3 float; outputs dConstraints / dConstraintNormals: N*3 opt_float,
 * dRobustWeights: N opt_float, dPrevConstraints: N*3 float (updated in place; initialise with anything that is not a position). */
typedef struct OptB200_PointGrid OptB200_PointGrid;
OptB200_PointGrid* OptB200_PointGridCreate(const float* dTargetPos, int M, float cellSize, void* stream);
void OptB200_PointGridFree(OptB200_PointGrid* g);
int OptB200_RnaCorrespondences(const OptB200_PointGrid* grid, const float* dTargetPos, const float* dTargetNormals,
                               const void* dSourcePos, const float* dSourceNormals, int N, float positionThreshold, float cosNormalThreshold, int maxK,
                               void* dConstraints, void* dConstraintNormals, void* dRobustWeights, float* dPrevConstraints, int doublePrecision, void* stream);

#ifdef __cplusplus
}
#endif
#endif
