// solver.h -- host-side plan: the generic GN/LM driver and the per-energy kernel interface.
//
// Mirrors the split the reference makes between the generic solver (API/src/solverGPUGaussNewton.t,
// `PlanData` :167-214, init :956-1007, step :1016-1177) and the per-problem function map
// `fmap.{cost,evalJTF,applyJTJ,computeCtC,modelcost,precompute,exclude}` the front end hands it
// (API/src/o.t:2425-2460).  Here the function map is a C++ interface (EnergyOps) implemented by
// hand-written sm_100a kernels, one translation unit per energy.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "common.cuh"

namespace optb {

// Solver parameters: names, types and defaults of the reference (solverGPUGaussNewton.t:26-39,
// :148-163): the nine LM parameters are `float` even in double-precision mode, the counters `int`.
struct SolverParameters {
    float min_relative_decrease = 1e-3f;
    float min_trust_region_radius = 1e-32f;
    float max_trust_region_radius = 1e16f;
    float q_tolerance = 0.0001f;
    float function_tolerance = 0.000001f;
    float trust_region_radius = 1e4f;
    float radius_decrease_factor = 2.0f;
    float min_lm_diagonal = 1e-6f;
    float max_lm_diagonal = 1e32f;
    int residual_reset_period = 10;
    int nIterations = 10;
    int lIterations = 10;
};

// One Unknown(...) image of the spec inside the solver's flat unknown vector.
struct UnknownArray {
    void* user = nullptr;   // caller's device buffer (AoS, `channels` scalars per element), updated in place
    long elements = 0;
    int channels = 1;
    long offset = 0;        // start inside every PCG vector, in scalars; 64-scalar aligned
};

// Halo exchange of one flat vector between slab neighbours.  Segments are this rank's
// boundary rows; peerVec are the neighbours' copies of the same vector (CUDA IPC), flags live in the
// ranks' mailbox blocks.  Used either by the stand-alone k_haloExchange or from inside an energy's
// J^T J kernel (push at kernel start, wait only before tiles that touch halo rows).
constexpr int kMaxHaloSegs = 6;          // unknown arrays x 2 directions (robust_nonrigid_alignment: Offset, Angle, RobustWeights)
struct HaloPlan {
    int nSeg;
    long srcOff[kMaxHaloSegs], dstOff[kMaxHaloSegs], len[kMaxHaloSegs];   // in scalars
    int dir[kMaxHaloSegs];               // 0 = to the upper neighbour, 1 = to the lower neighbour
    void* peerVec[2];                    // neighbour's copy of the vector being exchanged (nullptr: no neighbour)
    unsigned long long* peerFlag[2];     // neighbour's "halo arrived from me" flag (in-kernel halo paths: numbered by the r.z epochs)
    unsigned long long* myFlag[2];       // my flags, written by the upper / lower neighbour
    unsigned long long* peerXFlag[2];    // the same pair for the stand-alone exchange kernel, numbered by its own device-side counter ...
    unsigned long long* myXFlag[2];
    unsigned long long* epoch;           // ... this one (local; every rank runs the same sequence of exchange kernels)
};

struct LaunchCtx {
    cudaStream_t stream;
    ReduceWorkspace ws;
    int numSMs;
    bool pdl;   // launch inner-loop kernels with programmatic dependent launch (common.cuh)
    // LM decision on the device (solver.cu runDeviceIterations): 0 = launch as always; 1 = the kernels return at once when scalars[kLmGate] != 0
    // (the solve ended in an earlier enqueued iteration); 2 = they run only if the iteration just decided was rejected (scalars[kLmReject])
    // and the solve has not ended.  Only energies that answer deviceLmCapable() look at it.
    int gate = 0;
};

// Per-energy kernels ("function map").  T = opt_float (float or double).
template <class T>
struct EnergyOps {
    std::vector<UnknownArray> unknowns;
    long nUnknownsPadded = 0;          // length of every PCG vector
    bool usePreconditioner = true;     // spec's UsePreconditioner(); default false (o.t:214)
    bool usesGraph = false;
    virtual ~EnergyOps() {}
    // positional re-binding of the caller's table; called at Init and at every Step (util.t:664-692)
    virtual void bind(void** problemparams) = 0;
    // PCGInit1 (+_Graph,+_Finish): delta=0, r=-J^T F, pre=guardedInvert(diag J^T J), p=pre*r,
    // scalars[rzSlot] = r.p ; also refreshes whatever per-linearisation data applyJTJ needs.
    virtual void pcgInit(const LaunchCtx& c, T* delta, T* r, T* pre, T* p, int rzSlot) = 0;
    // PCGStep1 (+_Graph): out = (J^T J [+ diag(CtC)]) x, scalars[dotSlot] = x.out  (dotSlot<0: no reduction)
    // halo != nullptr (only when inKernelHalo()): the kernel itself pushes x's boundary rows to the neighbours and
    // waits for theirs before touching halo rows.
    virtual void applyJTJ(const LaunchCtx& c, const T* x, T* out, const T* CtC, int dotSlot, const HaloPlan* halo = nullptr) = 0;
    virtual bool inKernelHalo() const { return false; }
    // optional fusion of PCGStep3 into PCGStep1:  p <- z + beta*p  (beta = rz[new]/rz[old], guarded),
    // then out = (J^T J [+CtC]) p and scalars[dotSlot] = p.out.  Return false if not implemented.
    // p is double-buffered by the caller (pOld read incl. neighbours, pNew written) so the stencil never
    // sees a half-updated direction.  Called with dotSlot == -2 it only reports support (no launch).
    virtual bool step3ApplyJTJ(const LaunchCtx&, const T* /*z*/, const T* /*pOld*/, T* /*pNew*/, T* /*out*/, const T* /*CtC*/,
                               int /*rzNew*/, int /*rzOld*/, int /*dotSlot*/, const HaloPlan* /*haloZ*/ = nullptr) { return false; }
    // optional: the WHOLE inner loop (L iterations of PCGStep1/2/3 from the state pcgInit left) as one persistent
    // cooperative kernel -- for problems whose per-iteration kernels are launch/tail bound rather than HBM bound.
    // p/p2 are the two direction buffers (p holds the initial direction); *pFinal receives the buffer holding p
    // afterwards; the scalars kRz0/kRz1/kPAp are left exactly as the kernel-per-step loop would leave them.
    // Returns false (nothing launched) when not implemented / not applicable; `force` overrides the size heuristic.
    // haloZ != nullptr: multi-GPU (the plan's halo description of z, host copy); p's halo rows have been exchanged by the caller.
    virtual bool pcgLoop(const LaunchCtx&, int /*L*/, bool /*force*/, T* /*delta*/, T* /*r*/, T* /*z*/, T* /*p*/, T* /*p2*/, T* /*Ap*/,
                         const T* /*pre*/, T** /*pFinal*/, const HaloPlan* /*haloZ*/ = nullptr) { return false; }
    // optional, single GPU: PCGInit1 AND the LM head that follows it (PCGSaveSSq at the first nonlinear iteration, PCGComputeCtC,
    // PCGFinalizeDiagonal, Q0 = q, early-out flag = 0; solverGPUGaussNewton.t:1038-1053) in one launch.  The trust-region radius is read
    // from scalars[kRadius].  Returns false (nothing launched) when not implemented.
    virtual bool pcgInitLM(const LaunchCtx&, T* /*delta*/, T* /*r*/, T* /*pre*/, T* /*p*/, T* /*CtC*/, T* /*SSq*/, T* /*b*/, int /*rzSlot*/,
                           T /*minDiag*/, T /*maxDiag*/, bool /*saveSSq*/) { return false; }
    // optional, single GPU: the L iterations of the PCG loop (LM: with the residual reset every resetPeriod-th iteration and the zeta
    // early-out) as one capturable launch with the state on chip; in place on the plan's vectors, scalars (rz pair, p.Ap, q, Q0,
    // early-out flag) left exactly as L iterations of the kernel-per-step loop leave them.  Returns the number of iterations covered
    // (0: not implemented / does not fit; the caller emits the rest kernel by kernel).
    // modelCostDone != nullptr (LM, the call covers the whole loop): the kernel may also leave computeModelCost (:666-678) of the delta
    // it found in scalars[kModelCost] and says so there; the caller then skips modelCost().  updateDone != nullptr: likewise for
    // savePreviousUnknowns (LM, into prevX, laid out like the plan's vectors) + PCGLinearUpdate of every unknown array.
    // headArgs != nullptr (only after pcgLoopTakesHead(L, lm) answered true; the caller then launched neither pcgInit nor pcgInitLM): the
    // kernel also runs the LM head of pcgInitLM on its tiles first and leaves pre / CtC / b / SSq in memory as that call would.
    struct LoopHeadArgs { T* SSq; T minDiag, maxDiag; bool saveSSq; };
    virtual bool pcgLoopTakesHead(int /*L*/, bool /*lm*/) { return false; }
    virtual int pcgLoopOnChip(const LaunchCtx&, int /*L*/, bool /*lm*/, int /*resetPeriod*/, T /*qTolerance*/, T* /*delta*/, T* /*r*/, T* /*z*/, T* /*p*/, T* /*Ap*/,
                              const T* /*pre*/, const T* /*b*/, const T* /*CtC*/, bool* /*modelCostDone*/, T* /*prevX*/, bool* /*updateDone*/,
                              const LoopHeadArgs* /*headArgs*/) { return 0; }
    // computeCost (+_Graph): scalars[kCost] = 1/2 sum r^2
    virtual void cost(const LaunchCtx& c) = 0;
    // LM: CtC <- diag(J^T J)/radius (PCGComputeCtC), scalars[kModelCost] = 1/2 sum (r + J delta)^2
    virtual void computeCtC(const LaunchCtx&, T* /*CtC*/, T /*invRadius*/) { unsupported("computeCtC (LM)"); }
    virtual void modelCost(const LaunchCtx&, const T* /*delta*/) { unsupported("modelCost (LM)"); }
    // ComputedArray refresh (kernels.precompute, solverGPUGaussNewton.t:607-614)
    virtual void precompute(const LaunchCtx&) {}
    // optional: with the current toggles a whole LM iteration is pcgLoopOnChip (head, loop, model cost, update) + precompute (which also
    // leaves the cost), and both honour LaunchCtx::gate -- the plan may then keep the accept / reject decision on the device
    virtual bool deviceLmCapable() const { return false; }
    // multi-GPU: can THIS rank run pcgLoop's multi-rank form for its slab (tile shape fits the shared memory, all tiles co-resident,
    // cooperative launch available)?  Asked once per OptB200_PlanConnect, after setSlab(); the plan combines the answers of all ranks
    // so that either every rank takes the collective kernel or none does.  After a "yes" a refusal inside pcgLoop is fatal.
    virtual bool pcgLoopMultiFits() { return false; }
    // multi-GPU: kPeerScratchBytes of zero-initialised device memory of this rank and of its upper / lower neighbour (nullptr: none), mapped
    // through CUDA IPC at OptB200_PlanConnect -- for data the energy's collective kernels exchange as self-validating LL words
    virtual void setPeerScratch(void* /*mine*/, void* /*upper*/, void* /*lower*/) {}
    // ---- multi-GPU slab decomposition.  The plan's dimensions describe the LOCAL domain =
    // owned slab rows + halo rows; [ownedBegin, ownedEnd) are the owned local rows, upperRows/lowerRows the
    // local row counts of the neighbouring ranks (0 = no neighbour).  Returns false if unsupported.
    virtual bool setSlab(int /*ownedBegin*/, int /*ownedEnd*/, int /*upperRows*/, int /*lowerRows*/) { return false; }
    struct HaloSegment { long srcOff, dstOff, len; int dir; };          // scalars in the flat layout; dir 0 = up, 1 = down
    virtual std::vector<HaloSegment> haloSegments() const { return {}; }
    virtual std::vector<std::pair<long, long>> ownedScalarRanges() const { return {}; }
    // energies whose plan holds GLOBAL-size (replicated) vectors: the windows (owned + halo) the streaming kernels must touch
    virtual std::vector<std::pair<long, long>> processScalarRanges() const { return {}; }
    // called at every Opt_ProblemInit: graph indices / masks may have changed since the last solve
    virtual void newSolve() {}
    // true when, in the current state, none of the per-iteration entry points (pcgInit, precompute, cost, modelCost, computeCtC, applyJTJ,
    // step3ApplyJTJ, pcgLoop) allocates, synchronises or reads anything back: a whole nonlinear iteration can then be stream-captured
    virtual bool capturable() const { return false; }
    // Fingerprint of everything the inner-loop kernels (applyJTJ / step3ApplyJTJ) receive BY VALUE or as launch geometry:
    // weights, caller pointers, plan-owned buffers that bind()/pcgInit() may reallocate, residual counts.  The plan compares it
    // (after pcgInit, when all of that is final for the step) with the fingerprint its captured PCG-loop graph was built from and
    // re-captures on a mismatch -- the reference re-reads problemparams at every Init/Step (util.t:664-692).
    virtual unsigned long long captureKey() const = 0;
    static unsigned long long hashBytes(const void* data, size_t n, unsigned long long h = 1469598103934665603ull) {   // FNV-1a
        const unsigned char* b = (const unsigned char*)data;
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
        return h;
    }
    template <class V> static unsigned long long hashValue(const V& v, unsigned long long h = 1469598103934665603ull) { return hashBytes(&v, sizeof(V), h); }
    virtual const char* name() const = 0;
    static void unsupported(const char* what) {
        fprintf(stderr, "opt_b200: %s is not implemented for this energy\n", what);
        exit(2);
    }
    long layoutUnknowns() {   // assigns 64-scalar-aligned offsets
        long o = 0;
        for (auto& u : unknowns) { u.offset = o; o += (u.elements * u.channels + 63) / 64 * 64; }
        nUnknownsPadded = o;
        return o;
    }
};

struct KernelTiming { std::string name; float ms = 0; int count = 0; };

struct PlanBase {
    virtual ~PlanBase() {}
    virtual void setParameter(const char* name, const void* value) = 0;
    virtual void init(void** params) = 0;
    virtual int step(void** params) = 0;
    // Opt_ProblemSolve calls this between Init and its Step loop: a plan may run nonlinear iterations here without returning to the
    // caller in between (LM with the decision on the device); whatever is left is done by the Step loop
    virtual void runDeviceIterations(void** /*params*/) {}
    virtual double currentCost() const = 0;
    // ---- extension entry points (include/OptB200.h)
    virtual long numUnknowns() const = 0;
    virtual int unknownLayout(int maxArrays, long* offsets, long* elements, int* channels) const = 0;
    virtual void evalJTF(void** params, double* cost, void* negJtF, void* preInv, double* rDotP) = 0;
    virtual void applyJTJ(void** params, const void* x, void* out, double* xDotOut) = 0;
    virtual int readVector(const char* name, void* hostOut) = 0;
    virtual double readScalar(const char* name) = 0;
    virtual void setMaxLinearIterations(int k) = 0;
    virtual void setStream(cudaStream_t s) = 0;
    virtual int launchesLastStep() const = 0;
    virtual float pcgLoopMsLastStep() const = 0;
    virtual int profileKernels(void** params, int reps, int maxOut, const char** names, float* avgMs) = 0;
    virtual int commHandles(void* out, int outBytes) = 0;
    virtual int connect(int rank, int world, const void* allHandles, int ownedBegin, int ownedEnd, int upperRows, int lowerRows) = 0;
};

namespace fe { struct CompiledSpec; }
// factory implemented in solver.cu
PlanBase* makePlan(const std::string& energy, bool lm, bool doublePrecision, const unsigned* dims,
                   int verbosity, int collectTiming);

// plan over a spec produced by the .t front end (frontend.cpp) and compiled with NVRTC (dynamic_ops.cu)
PlanBase* makePlanDynamic(const fe::CompiledSpec& cs, const std::string& name, bool lm, bool doublePrecision, const unsigned* dims,
                          int verbosity, int collectTiming, std::string& error);
template <class T> EnergyOps<T>* makeDynamicEnergy(const fe::CompiledSpec& cs, const unsigned* dims, const std::string& name, std::string& error);
bool compileSpecToCubin(const fe::CompiledSpec& cs, bool doublePrecision, std::vector<char>& cubin, std::string& log);

// per-energy factories (energy_*.cu); return nullptr when the energy/precision is not built
template <class T> EnergyOps<T>* makeImageWarping(const unsigned* dims);
template <class T> EnergyOps<T>* makeArapMesh(const unsigned* dims);
template <class T> EnergyOps<T>* makeSfs(const unsigned* dims);   // energy_sfs.cu
template <class T> EnergyOps<T>* makeRna(const unsigned* dims);   // energy_rna.cu
template <class T> EnergyOps<T>* makeGenericEnergy(const std::string& energy, const unsigned* dims);

}  // namespace optb
