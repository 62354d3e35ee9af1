/* Opt.h -- C boundary of opt_b200 (libOpt.so).
 *
 * Drop-in for the reference's public header (reference: API/release/include/Opt.h:1-71): same type
 * names, same struct layout (4 ints), same ten entry points with the same argument meaning, so
 * the reference's callers -- examples/shared/OptSolver.h:43-91, examples/shared/OptUtils.h:47-108,
 * tests/minimal{,_graph_only}/main.cpp -- compile and link against libOpt.so without modification.  They include
 * this header inside `extern "C" { }` themselves, so it carries no linkage block of its own.
 *
 * Implementation: opt_b200/csrc/api.cpp (host, C++) over hand-written sm_100a kernels.
 * Everything declared here is plain C: opaque pointers, ints, void* tables.  No torch types.
 */
#pragma once

/* ---- opaque handles -------------------------------------------------------------------------- */
typedef struct Opt_State   Opt_State;    /* one solver context; not thread-safe                    */
typedef struct Opt_Problem Opt_Problem;  /* an energy spec (.t file) + solver kind                 */
typedef struct Opt_Plan    Opt_Plan;     /* a problem instantiated for concrete dimensions         */

/* ---- context options (reference Opt.h:10-30).  All-zero is the fast default. ----------------- */
struct Opt_InitializationParameters {
    int doublePrecision;            /* nonzero: unknowns and every opt_float array are fp64        */
    int verbosityLevel;             /* 0 silent, 1 per-iteration costs, >=2 also plan diagnostics  */
    int collectPerKernelTimingInfo; /* nonzero: per-kernel cudaEvent timing table at solve end     */
    int threadsPerBlock;            /* block size hint; not a positive multiple of 32 => 256       */
};
typedef struct Opt_InitializationParameters Opt_InitializationParameters;

/* Create an independent context on the CURRENT CUDA device (reference createwrapper.t:124-211). */
Opt_State* Opt_NewState(Opt_InitializationParameters params);

/* Register the energy in `filename` for solver `solverkind` ("gaussNewtonGPU" or "LMGPU",
 * reference o.t:122,315).  Cheap: the spec is only read at Opt_ProblemPlan (o.t:2521-2534). */
Opt_Problem* Opt_ProblemDefine(Opt_State* state, const char* filename, const char* solverkind);
void Opt_ProblemDelete(Opt_State* state, Opt_Problem* problem);

/* Instantiate the problem for `dimensions` (indexed by the spec's Dim(name, i)); allocates all
 * solver temporaries.  Returns NULL (after printing why) if the spec cannot be planned. */
Opt_Plan* Opt_ProblemPlan(Opt_State* state, Opt_Problem* problem, unsigned int* dimensions);
void Opt_PlanFree(Opt_State* state, Opt_Plan* plan);

/* Set a solver parameter by name; `value` points at an int (nIterations, lIterations,
 * residual_reset_period) or a float (the nine LM parameters), read at call time
 * (reference solverGPUGaussNewton.t:1205-1221).  Unknown names only warn. */
void Opt_SetSolverParameter(Opt_State* state, Opt_Plan* plan, const char* name, void* value);

/* problemparams: positional table in the order the spec declares -- device pointers for
 * Unknown/Array images, host pointers for scalar Params, and for a Graph a host int* edge count
 * followed by one device int32* per vertex slot (reference util.t:664-692).  Unknowns are updated
 * in place. */
void Opt_ProblemSolve(Opt_State* state, Opt_Plan* plan, void** problemparams);

/* Caller-driven outer loop: Init once, then Step until it returns 0 (finished). */
void Opt_ProblemInit(Opt_State* state, Opt_Plan* plan, void** problemparams);
int  Opt_ProblemStep(Opt_State* state, Opt_Plan* plan, void** problemparams);

/* Cost 1/2 sum r^2 at the current unknowns as of the last Init/Step (host value, no sync). */
double Opt_ProblemCurrentCost(Opt_State* state, Opt_Plan* plan);
