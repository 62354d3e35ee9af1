/* OptB200.h -- extension entry points of libOpt.so beyond the reference's Opt.h.
 *
 * Plain C ABI (pointers, ints, doubles).  Nothing here is needed by a reference caller; these exist so
 * that (a) the parity tests can compare individual hot-path kernels -- PCGInit1 / PCGStep1 / cost,
 * reference API/src/solverGPUGaussNewton.t:361-434,580-592 -- against the CPU oracle, (b) bench.py can
 * read device-side timings and launch counts, and (c) a multi-GPU host can shard a problem
 * (SURVEY.md section 8e).  All vectors use the plan's flat layout: the spec's unknown arrays
 * concatenated in declaration order, each starting at a 64-scalar-aligned offset
 * (OptB200_PlanUnknownLayout).
 */
#pragma once

#ifdef __cplusplus
extern "C" {
#endif
#include "Opt.h"

/* library identification: returns "opt_b200 <version> sm_100a" */
const char* OptB200_Version(void);

/* padded length (in scalars) of every PCG vector of the plan */
long OptB200_PlanNumUnknowns(Opt_Plan* plan);
/* fills up to maxArrays entries; returns the number of unknown arrays of the spec */
int OptB200_PlanUnknownLayout(Opt_Plan* plan, int maxArrays, long* offsets, long* elements, int* channels);

/* PCGInit1 at the current unknowns.  Host outputs (may be NULL): cost = 1/2 sum r^2,
 * negJtF = r0 = -J^T F, preInv = the inverted Jacobi preconditioner, rDotP = r0 . (preInv*r0). */
void OptB200_EvalJTF(Opt_Plan* plan, void** problemparams, double* cost, void* negJtF, void* preInv, double* rDotP);
/* PCGStep1: out = J^T J x linearised at the current unknowns; x and out are HOST vectors in the flat layout. */
void OptB200_ApplyJTJ(Opt_Plan* plan, void** problemparams, const void* x, void* out, double* xDotOut);

/* Debug hook for iteration-level parity: when k >= 0, Opt_ProblemStep runs PCGInit1 and only k PCG
 * iterations, leaves delta/r/z/p/Ap_X in place for OptB200_ReadVector and does NOT update the unknowns. */
void OptB200_SetMaxLinearIterations(Opt_Plan* plan, int k);
/* name in {delta,r,z,p,Ap_X,preconditioner,CtC,b}; returns 0 on success */
int OptB200_ReadVector(Opt_Plan* plan, const char* name, void* hostOut);
/* name in {scanAlphaDenominator,rz0,rz1,q,flags,trust_region_radius,cost,modelCost,nIter,deviceIterations}
 * (nIter: nonlinear iterations completed in the current solve, solverGPUGaussNewton.t:1173; deviceIterations: how many of them the last
 * Opt_ProblemSolve ran with the LM decision on the device, OPTB_DEVICE_LM=1 -- 0 otherwise) */
double OptB200_ReadScalar(Opt_Plan* plan, const char* name);

/* Order the plan's work after `stream` (a cudaStream_t) instead of the legacy NULL stream. */
void OptB200_PlanSetStream(Opt_Plan* plan, void* stream);
/* kernels launched by the last Opt_ProblemStep, and device time (ms, CUDA events) of its inner PCG loop */
int OptB200_LaunchesLastStep(Opt_Plan* plan);
float OptB200_PcgLoopMsLastStep(Opt_Plan* plan);
/* Replays each inner-loop kernel `reps` times between two CUDA events on the plan's stream and
 * returns the average duration per launch (ms).  names[i] point at static strings.  The PCG state
 * is scratch afterwards -- call Opt_ProblemInit before solving again.  Returns the entry count. */
int OptB200_ProfileKernels(Opt_Plan* plan, void** problemparams, int reps, int maxOut, const char** names, float* avgMs);
/* ---- multi-GPU slab decomposition (one process per GPU) --------------------------
 * Each rank plans its LOCAL domain = owned slab + 1 halo row towards every neighbour and binds local
 * buffers.  OptB200_PlanCommHandles writes 4 CUDA IPC handles (256 bytes: scalar mailbox block, z, p, delta) and returns the byte count; the host
 * all-gathers them (any transport, e.g. torch.distributed) and passes the rank-ordered concatenation to
 * OptB200_PlanConnect together with this rank's owned local rows [ownedBegin, ownedEnd) and the LOCAL row
 * counts of the upper / lower neighbour (0 = none).  Afterwards Opt_ProblemInit/Step must be called
 * collectively by all ranks: the PCG scalars are summed through peer-memory mailboxes and the halo rows of
 * z / p travel by direct NVLink stores, all issued from inside the solver's kernels.  Returns 0 on success. */
int OptB200_PlanCommHandles(Opt_Plan* plan, void* out, int outBytes);
int OptB200_PlanConnect(Opt_Plan* plan, int rank, int world, const void* allHandles, int ownedBegin, int ownedEnd, int upperRows, int lowerRows);
/* Front-end probe (no GPU needed): writes the registry name of the energy in `filename` to `out` and
 * returns 0, or writes the rejection message and returns 1. */
int OptB200_SpecEnergyName(const char* filename, char* out, int outLen);
/* General front end, host only (no GPU needed): runs the .t file (Lua subset + Opt DSL), generates the residual code
 * and compiles it with NVRTC for sm_100a.  Returns 0 and a summary in msg, or nonzero and the error text. */
int OptB200_FrontendCompile(const char* filename, int doublePrecision, char* msg, int msgLen);

#ifdef __cplusplus
}
#endif
