// api.cpp -- the C boundary: Opt_* (include/Opt.h) and OptB200_* (include/OptB200.h).
//
// Replaces the reference's wrapper layer (API/src/createwrapper.t:124-220, which owns a lua_State
// and forwards each Opt_* to a Terra function) and the plan lifecycle of API/src/o.t:126-141,
// :861-882, :2521-2558, with plain C++ over the CUDA plans in solver.cu.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "solver.h"
#include "spec.h"
#include "frontend.h"

extern "C" {
#include "../../include/OptB200.h"
}

struct ProblemDef { std::string filename, solverkind; bool alive = true; };

struct Opt_State {
    Opt_InitializationParameters params;
    std::vector<ProblemDef> problems;
};
struct Opt_Plan { optb::PlanBase* impl; };

static const char* kVersion = "opt_b200 0.1.0 sm_100a";

extern "C" {

const char* OptB200_Version(void) { return kVersion; }

Opt_State* Opt_NewState(Opt_InitializationParameters params) {
    Opt_State* s = new Opt_State();
    s->params = params;
    // reference createwrapper.t:149-153: an invalid block size falls back to 256
    if (s->params.threadsPerBlock <= 0 || s->params.threadsPerBlock % 32 != 0) s->params.threadsPerBlock = 256;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        // No CPU fallback exists: every solve runs on the CUDA device.  Fail loudly.
        fprintf(stderr, "opt_b200: no usable CUDA device (%s); this library has no CPU path\n", cudaGetErrorString(e));
        delete s;
        return nullptr;
    }
    if (s->params.verbosityLevel > 1) {
        cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
        printf("opt_b200: device %d %s, %d SMs, cc %d.%d\n", dev, prop.name, prop.multiProcessorCount, prop.major, prop.minor);
    }
    return s;
}

// As in the reference (o.t:136-141, :2521-2525) the returned handle is just the problem's index + 1.
Opt_Problem* Opt_ProblemDefine(Opt_State* state, const char* filename, const char* solverkind) {
    if (!state || !filename || !solverkind) return nullptr;
    ProblemDef d; d.filename = filename; d.solverkind = solverkind;
    state->problems.push_back(d);
    return reinterpret_cast<Opt_Problem*>(state->problems.size());
}
void Opt_ProblemDelete(Opt_State* state, Opt_Problem* problem) {
    size_t id = reinterpret_cast<size_t>(problem);
    if (state && id >= 1 && id <= state->problems.size()) state->problems[id - 1].alive = false;
}

Opt_Plan* Opt_ProblemPlan(Opt_State* state, Opt_Problem* problem, unsigned int* dimensions) {
    size_t id = reinterpret_cast<size_t>(problem);
    if (!state || id < 1 || id > state->problems.size() || !state->problems[id - 1].alive) {
        fprintf(stderr, "opt_b200: Opt_ProblemPlan: invalid problem handle\n");
        return nullptr;
    }
    const ProblemDef& d = state->problems[id - 1];
    // reference o.t:122: only these two solver kinds exist; LM-ness is the substring "LM" (o.t:315)
    if (d.solverkind != "gaussNewtonGPU" && d.solverkind != "LMGPU") {
        fprintf(stderr, "opt_b200: unknown solver kind '%s' (expected gaussNewtonGPU or LMGPU)\n", d.solverkind.c_str());
        return nullptr;
    }
    bool lm = d.solverkind.find("LM") != std::string::npos;
    // Two tiers.  (a) the spec's canonical token stream is in the registry: its hand-written / statically compiled
    // kernel set is used.  (b) anything else (or everything, with OPTB_FRONTEND=1): the .t file is executed by the
    // front end, its residual code is generated, compiled with NVRTC and run on the generic engine -- the
    // counterpart of the reference's Lua/Terra JIT at this point (o.t:861-882).
    optb::SpecMatch m = optb::matchSpecFile(d.filename);
    optb::PlanBase* impl = nullptr;
    if (m.ok && !getenv("OPTB_FRONTEND")) {
        impl = optb::makePlan(m.energy, lm, state->params.doublePrecision != 0, dimensions, state->params.verbosityLevel, state->params.collectPerKernelTimingInfo);
        if (!impl) { fprintf(stderr, "opt_b200: energy '%s' is registered but has no kernel set in this build\n", m.energy.c_str()); return nullptr; }
    } else {
        optb::fe::CompiledSpec cs = optb::fe::compileSpecFile(d.filename);
        if (!cs.ok) { fprintf(stderr, "opt_b200: %s\n", cs.error.c_str()); return nullptr; }   // like the reference: message + NULL plan (o.t:862-880)
        std::string err;
        impl = optb::makePlanDynamic(cs, d.filename, lm, state->params.doublePrecision != 0, dimensions, state->params.verbosityLevel, state->params.collectPerKernelTimingInfo, err);
        if (!impl) { fprintf(stderr, "opt_b200: %s: %s\n", d.filename.c_str(), err.c_str()); return nullptr; }
        if (state->params.verbosityLevel > 0) printf("compiled '%s' through the front end: %zu scalar residual groups\n", d.filename.c_str(), cs.groups.size());
    }
    if (state->params.verbosityLevel > 0) printf("problem plan complete\n");
    return new Opt_Plan{ impl };
}
void Opt_PlanFree(Opt_State*, Opt_Plan* plan) { if (plan) { delete plan->impl; delete plan; } }

void Opt_SetSolverParameter(Opt_State*, Opt_Plan* plan, const char* name, void* value) { plan->impl->setParameter(name, value); }
void Opt_ProblemInit(Opt_State*, Opt_Plan* plan, void** problemparams) { plan->impl->init(problemparams); }
int Opt_ProblemStep(Opt_State*, Opt_Plan* plan, void** problemparams) { return plan->impl->step(problemparams); }
void Opt_ProblemSolve(Opt_State* s, Opt_Plan* plan, void** problemparams) {   // o.t:2548-2551
    Opt_ProblemInit(s, plan, problemparams);
    plan->impl->runDeviceIterations(problemparams);   // no-op unless the plan keeps the LM decision on the device (OPTB_DEVICE_LM=1)
    while (Opt_ProblemStep(s, plan, problemparams)) {}
}
double Opt_ProblemCurrentCost(Opt_State*, Opt_Plan* plan) { return plan->impl->currentCost(); }

// ---- extensions
long OptB200_PlanNumUnknowns(Opt_Plan* plan) { return plan->impl->numUnknowns(); }
int OptB200_PlanUnknownLayout(Opt_Plan* plan, int maxArrays, long* offsets, long* elements, int* channels) {
    return plan->impl->unknownLayout(maxArrays, offsets, elements, channels);
}
void OptB200_EvalJTF(Opt_Plan* plan, void** pp, double* cost, void* negJtF, void* preInv, double* rDotP) { plan->impl->evalJTF(pp, cost, negJtF, preInv, rDotP); }
void OptB200_ApplyJTJ(Opt_Plan* plan, void** pp, const void* x, void* out, double* xDotOut) { plan->impl->applyJTJ(pp, x, out, xDotOut); }
void OptB200_SetMaxLinearIterations(Opt_Plan* plan, int k) { plan->impl->setMaxLinearIterations(k); }
int OptB200_ReadVector(Opt_Plan* plan, const char* name, void* hostOut) { return plan->impl->readVector(name, hostOut); }
double OptB200_ReadScalar(Opt_Plan* plan, const char* name) { return plan->impl->readScalar(name); }
void OptB200_PlanSetStream(Opt_Plan* plan, void* stream) { plan->impl->setStream((cudaStream_t)stream); }
int OptB200_LaunchesLastStep(Opt_Plan* plan) { return plan->impl->launchesLastStep(); }
float OptB200_PcgLoopMsLastStep(Opt_Plan* plan) { return plan->impl->pcgLoopMsLastStep(); }
int OptB200_ProfileKernels(Opt_Plan* plan, void** pp, int reps, int maxOut, const char** names, float* avgMs) { return plan->impl->profileKernels(pp, reps, maxOut, names, avgMs); }
int OptB200_PlanCommHandles(Opt_Plan* plan, void* out, int outBytes) { return plan->impl->commHandles(out, outBytes); }
int OptB200_PlanConnect(Opt_Plan* plan, int rank, int world, const void* allHandles, int ownedBegin, int ownedEnd, int upperRows, int lowerRows) {
    return plan->impl->connect(rank, world, allHandles, ownedBegin, ownedEnd, upperRows, lowerRows);
}
int OptB200_FrontendCompile(const char* filename, int doublePrecision, char* msg, int msgLen) {
    optb::fe::CompiledSpec cs = optb::fe::compileSpecFile(filename);
    if (!cs.ok) { snprintf(msg, msgLen, "%s", cs.error.c_str()); return 1; }
    std::vector<char> cubin; std::string log;
    if (!optb::compileSpecToCubin(cs, doublePrecision != 0, cubin, log)) { snprintf(msg, msgLen, "NVRTC: %s", log.c_str()); return 2; }
    snprintf(msg, msgLen, "ok: %zu residual groups, %zu bytes of generated CUDA, %zu bytes of sm_100a cubin", cs.groups.size(), cs.cudaSource.size(), cubin.size());
    return 0;
}
int OptB200_SpecEnergyName(const char* filename, char* out, int outLen) {
    optb::SpecMatch m = optb::matchSpecFile(filename);
    snprintf(out, outLen, "%s", m.ok ? m.energy.c_str() : m.message.c_str());
    return m.ok ? 0 : 1;
}

}  // extern "C"
