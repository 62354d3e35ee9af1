// dynamic_ops.cu -- kernel set for ARBITRARY energy specifications: the .t front end (frontend.cpp) generates
// the residual code, NVRTC compiles it for sm_100a together with generic_device.cuh at Opt_ProblemPlan time
// (the reference JIT-compiles Terra through LLVM->PTX at the same point, API/src/o.t:861-882,
// API/src/util.t:845-904), and the generic cached-Jacobian engine does the rest.
//
// libnvrtc and the three driver entry points used to load/launch the module are resolved at run time
// (dlopen / cudaGetDriverEntryPoint), so libOpt.so carries no link-time dependency on libcuda or libnvrtc
// and still loads on machines without a GPU (where only the front end itself can be exercised).
#include <dlfcn.h>
#include <cuda.h>
#include <nvrtc.h>
#include "frontend.h"
#include "generic_engine.cuh"

namespace optb {

static const char kGenericDeviceSrc[] =
#include "generic_device_src.inc"
    ;

namespace {
struct Nvrtc {
    void* lib = nullptr;
    nvrtcResult (*createProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*compileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*getProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*getProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*getCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*getCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*destroyProgram)(nvrtcProgram*) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        for (const char* n : { "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12" }) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = "cannot load libnvrtc (needed to compile energy specifications that are not built in)"; return false; }
#define NV(sym, name) sym = (decltype(sym))dlsym(lib, name); if (!sym) { err = std::string("libnvrtc lacks ") + name; return false; }
        NV(createProgram, "nvrtcCreateProgram") NV(compileProgram, "nvrtcCompileProgram") NV(getProgramLogSize, "nvrtcGetProgramLogSize")
        NV(getProgramLog, "nvrtcGetProgramLog") NV(getCUBINSize, "nvrtcGetCUBINSize") NV(getCUBIN, "nvrtcGetCUBIN") NV(destroyProgram, "nvrtcDestroyProgram")
#undef NV
        return true;
    }
};
Nvrtc g_nvrtc;

template <class F> bool driverFn(const char* name, F& out) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn) return false;
    out = (F)fn; return true;
}
}  // namespace

// Compiles generated spec source to a cubin for sm_100a.  Host-only: works without a GPU (used by the CPU tests).
bool compileSpecToCubin(const fe::CompiledSpec& cs, bool doublePrecision, std::vector<char>& cubin, std::string& log) {
    if (!g_nvrtc.load(log)) return false;
    std::string src = kGenericDeviceSrc;
    src += "\nusing namespace optb;\n";
    src += cs.cudaSource;
    src += std::string("\ntypedef ") + (doublePrecision ? "double" : "float") + " TT;\n";
    src += R"SRC(
extern "C" __global__ void __launch_bounds__(128) ge_fill(GenParams<TT> P, GeView V, int g, TT* F, TT* val, int* col, int* entRow, unsigned char* rowCost) {
    geFillBody<GenSpec<TT>, TT>(P, V, g, F, val, col, entRow, rowCost);
}
extern "C" __global__ void __launch_bounds__(128) ge_cost(GenParams<TT> P, GeView V, int g, double* partials) { geCostBody<GenSpec<TT>, TT>(P, V, g, partials); }
extern "C" __global__ void __launch_bounds__(256) ge_active(GenParams<TT> P, int a, long elements, int channels, long offset, unsigned char* active) {
    geActiveBody<GenSpec<TT>, TT>(P, a, elements, channels, offset, active);
}
)SRC";
    nvrtcProgram prog;
    if (g_nvrtc.createProgram(&prog, src.c_str(), "opt_b200_generated_spec.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) { log = "nvrtcCreateProgram failed"; return false; }
    const char* opts[] = { "--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo" };
    nvrtcResult rc = g_nvrtc.compileProgram(prog, 3, opts);
    size_t ls = 0; g_nvrtc.getProgramLogSize(prog, &ls);
    if (ls > 1) { log.resize(ls); g_nvrtc.getProgramLog(prog, &log[0]); }
    if (rc != NVRTC_SUCCESS) { g_nvrtc.destroyProgram(&prog); if (getenv("OPTB_DUMP_GENERATED")) fprintf(stderr, "%s\n", src.c_str()); return false; }
    size_t cs_ = 0; g_nvrtc.getCUBINSize(prog, &cs_);
    cubin.resize(cs_); g_nvrtc.getCUBIN(prog, cubin.data());
    g_nvrtc.destroyProgram(&prog);
    return cs_ > 0;
}

template <class T>
struct DynamicOps : GenericOpsBase<T> {
    fe::CompiledSpec cs;
    std::vector<unsigned> dims;
    GenParams<T> P{};
    std::string label;
    CUmodule mod = nullptr; CUfunction fFill = nullptr, fCost = nullptr, fActive = nullptr;
    CUresult (*pModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*pModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*pLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*pModuleUnload)(CUmodule) = nullptr;
    std::vector<int> slotBase;   // global slot numbering: first slot of each graph
    bool good = false; std::string error;

    long cardinality(const std::vector<int>& d) const { long n = 1; for (int k : d) n *= (long)dims[k]; return n; }
    DynamicOps(const fe::CompiledSpec& spec, const unsigned* dimensions, const std::string& name) : cs(spec), label(name) {
        dims.assign(dimensions, dimensions + (cs.maxDimIndex + 1));
        // reference o.t:651-656: in double-precision mode every unknown must be an opt_float (= double) array
        for (auto& im : cs.images) if (im.unknown && sizeof(T) == 8 && im.type == fe::ScalarType::Float) { error = "unknown '" + im.name + "' is declared float but the state is double precision (declare it opt_float)"; return; }
        int nu = 0; for (auto& im : cs.images) if (im.unknown) nu = std::max(nu, im.unknownId + 1);
        this->unknowns.resize(nu);
        for (auto& im : cs.images) if (im.unknown) { this->unknowns[im.unknownId].elements = cardinality(im.dims); this->unknowns[im.unknownId].channels = im.channels; }
        this->usePreconditioner = cs.usePreconditioner; this->usesGraph = !cs.graphs.empty();
        int b = 0; for (auto& g : cs.graphs) { slotBase.push_back(b); b += (int)g.slots.size(); }
        if (b > kGenMaxGraphSlots) { error = "too many graph vertex slots"; return; }
        this->setup();
        std::vector<char> cubin; std::string log;
        if (!compileSpecToCubin(cs, sizeof(T) == 8, cubin, log)) { error = "NVRTC compilation of the generated spec failed:\n" + log; return; }
        if (!driverFn("cuModuleLoadData", pModuleLoadData) || !driverFn("cuModuleGetFunction", pModuleGetFunction) || !driverFn("cuLaunchKernel", pLaunchKernel) ||
            !driverFn("cuModuleUnload", pModuleUnload)) { error = "CUDA driver entry points unavailable"; return; }
        OPT_CUDA_CHECK(cudaFree(0));   // make sure the primary context is current for the driver calls
        if (pModuleLoadData(&mod, cubin.data()) != CUDA_SUCCESS) { error = "cuModuleLoadData failed for the generated spec"; return; }
        if (pModuleGetFunction(&fFill, mod, "ge_fill") != CUDA_SUCCESS || pModuleGetFunction(&fCost, mod, "ge_cost") != CUDA_SUCCESS ||
            pModuleGetFunction(&fActive, mod, "ge_active") != CUDA_SUCCESS) { error = "generated module lacks its kernels"; return; }
        good = true;
    }
    ~DynamicOps() override { if (mod && pModuleUnload) pModuleUnload(mod); }
    const char* name() const override { return label.c_str(); }
    static double readScalar(const void* p, fe::ScalarType t) {
        switch (t) {
            case fe::ScalarType::Float: return (double)*(const float*)p;
            case fe::ScalarType::Double: return *(const double*)p;
            case fe::ScalarType::OptFloat: return (double)*(const T*)p;
            case fe::ScalarType::Int32: return (double)*(const int*)p;
            default: return (double)*(const unsigned char*)p;
        }
    }
    void bindSpec(void** q, GeLayout& N) override {
        for (size_t i = 0; i < cs.images.size(); ++i) {
            const fe::ImageDecl& im = cs.images[i];
            P.img[i] = q[im.paramIndex];
            if (im.unknown) {
                this->unknowns[im.unknownId].user = q[im.paramIndex];
                P.U[im.unknownId] = GeArray<T>{ (T*)q[im.paramIndex], this->unknowns[im.unknownId].elements, im.channels, this->unknowns[im.unknownId].offset };
            }
        }
        for (size_t k = 0; k < dims.size() && k < (size_t)kGenMaxDims; ++k) P.dim[k] = dims[k];
        for (size_t k = 0; k < cs.params.size(); ++k) P.par[k] = readScalar(q[cs.params[k].paramIndex], cs.params[k].type);
        std::vector<long> gcount(cs.graphs.size(), 0);
        for (size_t g = 0; g < cs.graphs.size(); ++g) {
            gcount[g] = *(const int*)q[cs.graphs[g].countIndex];
            for (size_t s = 0; s < cs.graphs[g].slots.size(); ++s) P.gidx[slotBase[g] + s] = (const int*)q[cs.graphs[g].slots[s].paramIndex];
        }
        N.nGroups = (int)cs.groups.size();
        for (int g = 0; g < N.nGroups; ++g) {
            const fe::GroupDecl& gd = cs.groups[g];
            N.group[g] = GeGroup{ gd.centered ? cardinality(gd.dims) : gcount[gd.graph], 1, gd.ndeps, gd.centered ? 1 : 0, gd.centerUnknown };
        }
    }
    void launch(CUfunction f, int gr, int block, cudaStream_t st, void** args) {
        CUresult rc = pLaunchKernel(f, (unsigned)gr, 1, 1, (unsigned)block, 1, 1, 0, (CUstream)st, args, nullptr);
        if (rc != CUDA_SUCCESS) { fprintf(stderr, "Cuda reported error %d launching a generated kernel\n", (int)rc); exit((int)rc); }
    }
    void launchFill(const LaunchCtx& c, int g, int gr) override {
        GeView V = geView(this->L, g);
        void* args[] = { &P, &V, &g, &this->F, &this->val, &this->col, &this->entRow, &this->rowCost };
        launch(fFill, gr, 128, c.stream, args);
    }
    void launchCost(const LaunchCtx& c, int g, int gr, double* partials) override {
        GeView V = geView(this->L, g);
        void* args[] = { &P, &V, &g, &partials };
        launch(fCost, gr, 128, c.stream, args);
    }
    void launchActive(const LaunchCtx& c, int a, int gr) override {
        long el = this->unknowns[a].elements, off = this->unknowns[a].offset; int ch = this->unknowns[a].channels;
        void* args[] = { &P, &a, &el, &ch, &off, &this->active };
        launch(fActive, gr, 256, c.stream, args);
    }
};

template <class T>
EnergyOps<T>* makeDynamicEnergy(const fe::CompiledSpec& cs, const unsigned* dims, const std::string& name, std::string& error) {
    auto* d = new DynamicOps<T>(cs, dims, name);
    if (!d->good) { error = d->error; delete d; return nullptr; }
    return d;
}
template EnergyOps<float>* makeDynamicEnergy<float>(const fe::CompiledSpec&, const unsigned*, const std::string&, std::string&);
template EnergyOps<double>* makeDynamicEnergy<double>(const fe::CompiledSpec&, const unsigned*, const std::string&, std::string&);

}  // namespace optb
