// Host-side execution of the code the .t front end GENERATES (opt_b200/csrc/frontend.cpp) through the generic engine's own
// device bodies (geActiveBody / geFillBody of generic_device.cuh), compiled for the CPU behind a few shims: with
// blockDim = gridDim = 1 the kernels' grid-stride loops visit every instance sequentially.  Prints the cost and writes
// J^T F and diag(J^T J) so that tests/test_frontend_host_parity.py can compare them with the CPU oracle -- a parity check of
// interpreter + code generator + dual-number Jacobians that needs no GPU.
//   build: g++ -std=c++17 -DGEN_SOURCE='"<generated>.inc"' spec_eval_host.cpp ../../opt_b200/csrc/frontend.cpp
//   run:   spec_eval_host <spec.t> <dataDir> <out.bin> dim0 [dim1 ...]     (dataDir/p<k>.bin = raw bytes of problemparams[k])
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
struct Dim3Shim { unsigned x, y, z; };
static Dim3Shim blockIdx = { 0, 0, 0 }, threadIdx = { 0, 0, 0 }, blockDim = { 1, 1, 1 }, gridDim = { 1, 1, 1 };
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
static inline void __syncthreads() {}
#include "../../opt_b200/csrc/generic_device.cuh"
#include "../../opt_b200/csrc/frontend.h"
namespace optb {
#include GEN_SOURCE
}
using namespace optb;
#ifndef HOST_T
#define HOST_T double
#endif
typedef HOST_T T;   // opt_float of the run (-DHOST_T=float for specs that declare float unknowns)

static std::vector<char> readFile(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<char> b((size_t)n + 8, 0);
    if (n > 0 && fread(b.data(), 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read %s\n", path.c_str()); exit(2); }
    fclose(f);
    return b;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: spec_eval_host spec.t dataDir out.bin dim0 [dim1 ...]\n"); return 2; }
    fe::CompiledSpec cs = fe::compileSpecFile(argv[1]);
    if (!cs.ok) { fprintf(stderr, "front end: %s\n", cs.error.c_str()); return 1; }
    const std::string dir = argv[2];
    std::vector<long> dims; for (int i = 4; i < argc; ++i) dims.push_back(atol(argv[i]));
    if ((int)dims.size() < cs.maxDimIndex + 1) { fprintf(stderr, "need %d dimensions\n", cs.maxDimIndex + 1); return 2; }
    auto card = [&](const std::vector<int>& d) { long n = 1; for (int k : d) n *= dims[k]; return n; };
    std::vector<std::vector<char>> blobs;   // keeps every parameter alive
    auto param = [&](int k) -> void* { blobs.push_back(readFile(dir + "/p" + std::to_string(k) + ".bin")); return blobs.back().data(); };
    blobs.reserve(256);

    GenParams<T> P{};
    struct U { long elements; int channels; long offset; };
    std::vector<U> unknowns;
    for (auto& im : cs.images) if (im.unknown) { if ((int)unknowns.size() <= im.unknownId) unknowns.resize(im.unknownId + 1); unknowns[im.unknownId] = U{ card(im.dims), im.channels, 0 }; }
    long nPadded = 0;
    for (auto& u : unknowns) { u.offset = nPadded; nPadded += (u.elements * u.channels + 63) / 64 * 64; }   // EnergyOps::layoutUnknowns
    for (size_t i = 0; i < cs.images.size(); ++i) {
        const fe::ImageDecl& im = cs.images[i];
        void* q = param(im.paramIndex);
        P.img[i] = q;
        if (im.unknown) {
            const bool okType = im.type == fe::ScalarType::OptFloat || (sizeof(T) == 8 ? im.type == fe::ScalarType::Double : im.type == fe::ScalarType::Float);
            if (!okType) { fprintf(stderr, "unknown '%s' does not have the precision this harness was built for\n", im.name.c_str()); return 3; }
            P.U[im.unknownId] = GeArray<T>{ (T*)q, unknowns[im.unknownId].elements, im.channels, unknowns[im.unknownId].offset };
        }
    }
    for (size_t k = 0; k < dims.size() && k < (size_t)kGenMaxDims; ++k) P.dim[k] = dims[k];
    for (size_t k = 0; k < cs.params.size(); ++k) {
        const void* q = param(cs.params[k].paramIndex);
        switch (cs.params[k].type) {     // DynamicOps::readScalar
            case fe::ScalarType::Float: P.par[k] = (double)*(const float*)q; break;
            case fe::ScalarType::Double: P.par[k] = *(const double*)q; break;
            case fe::ScalarType::OptFloat: P.par[k] = (double)*(const T*)q; break;
            case fe::ScalarType::Int32: P.par[k] = (double)*(const int*)q; break;
            default: P.par[k] = (double)*(const unsigned char*)q;
        }
    }
    std::vector<long> gcount(cs.graphs.size(), 0);
    int slot = 0;
    for (size_t g = 0; g < cs.graphs.size(); ++g) {
        gcount[g] = *(const int*)param(cs.graphs[g].countIndex);
        for (auto& s : cs.graphs[g].slots) P.gidx[slot++] = (const int*)param(s.paramIndex);
    }
    // layout of the row cache (one residual row per group instance: the front end emits scalar residual groups)
    const int nG = (int)cs.groups.size();
    std::vector<GeView> views(nG);
    long nRows = 0, nEnt = 0;
    for (int g = 0; g < nG; ++g) {
        const fe::GroupDecl& gd = cs.groups[g];
        const long count = gd.centered ? card(gd.dims) : gcount[gd.graph];
        views[g] = GeView{ GeGroup{ count, 1, gd.ndeps, gd.centered ? 1 : 0, gd.centerUnknown }, nRows, nEnt };
        nRows += count; nEnt += count * gd.ndeps;
    }
    std::vector<T> F(nRows + 1), val(nEnt + 1);
    std::vector<int> col(nEnt + 1, -1), entRow(nEnt + 1, 0);
    std::vector<unsigned char> rowCost(nRows + 1, 0), active(nPadded + 1, 0);
    for (size_t a = 0; a < unknowns.size(); ++a) geActiveBody<GenSpec<T>, T>(P, (int)a, unknowns[a].elements, unknowns[a].channels, unknowns[a].offset, active.data());
    for (int g = 0; g < nG; ++g) if (views[g].G.count > 0) geFillBody<GenSpec<T>, T>(P, views[g], g, F.data(), val.data(), col.data(), entRow.data(), rowCost.data());
    double cost = 0;
    for (long r = 0; r < nRows; ++r) if (rowCost[r]) cost += 0.5 * F[r] * F[r];
    std::vector<double> jtf(nPadded, 0.0), diag(nPadded, 0.0);
    for (long e = 0; e < nEnt; ++e) if (col[e] >= 0 && active[col[e]]) { jtf[col[e]] += val[e] * F[entRow[e]]; diag[col[e]] += val[e] * val[e]; }
    // output: cost as text; J^T F and diag(J^T J) per unknown array in spec order WITHOUT the alignment padding (the oracle's layout)
    FILE* out = fopen(argv[3], "wb");
    if (!out) { fprintf(stderr, "cannot write %s\n", argv[3]); return 2; }
    for (int pass = 0; pass < 2; ++pass)
        for (auto& u : unknowns) fwrite((pass ? diag.data() : jtf.data()) + u.offset, sizeof(double), (size_t)(u.elements * u.channels), out);
    fclose(out);
    printf("COST %.17g ROWS %ld ENTRIES %ld UNKNOWNS %ld\n", cost, nRows, nEnt, nPadded);
    return 0;
}
