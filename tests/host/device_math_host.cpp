// Host-side check of the device dual-number math in opt_b200/csrc/generic_device.cuh: the header is compiled for the CPU
// behind a few shims (the kernels' bodies are templates and are never instantiated here), and every single-argument
// function's derivative slot is compared with a central finite difference.  Built and run by tests/test_golden_and_capi.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
struct Dim3Shim { unsigned x, y, z; };
static Dim3Shim blockIdx, threadIdx, blockDim, gridDim;
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
static inline void __syncthreads() {}
#include "../../opt_b200/csrc/generic_device.cuh"

using namespace optb;

template <class F, class G>
static int check(const char* name, F dual, G plain, double x) {
    GDual<double> a(x); a.d[0] = 1.0; a.d[3] = -2.0;     // two derivative slots with different seeds
    GDual<double> r = dual(a);
    const double h = 1e-6, fd = (plain(x + h) - plain(x - h)) / (2 * h);
    int bad = 0;
    if (std::fabs(r.v - plain(x)) > 1e-14 * std::fmax(1.0, std::fabs(r.v))) bad = 1;
    if (std::fabs(r.d[0] - fd) > 1e-7 * std::fmax(1.0, std::fabs(fd))) bad = 1;
    if (std::fabs(r.d[3] + 2.0 * fd) > 2e-7 * std::fmax(1.0, std::fabs(fd))) bad = 1;
    for (int i = 0; i < kGeMaxDeps; ++i) if (i != 0 && i != 3 && r.d[i] != 0.0) bad = 1;
    printf("%-6s x=%5.2f value %+.12f d/dx %+.9f fd %+.9f %s\n", name, x, r.v, r.d[0], fd, bad ? "MISMATCH" : "ok");
    return bad;
}

#define CHECK1(NAME, FN, STD, X) bad += check(NAME, [](const GDual<double>& a) { return FN(a); }, [](double x) { return (double)STD(x); }, X)

int main() {
    int bad = 0;
    CHECK1("sqrt", geSqrt, std::sqrt, 1.7);   CHECK1("sin", geSin, std::sin, 0.9);       CHECK1("cos", geCos, std::cos, 0.9);
    CHECK1("exp", geExp, std::exp, 0.4);      CHECK1("log", geLog, std::log, 1.3);       CHECK1("abs", geAbs, std::fabs, -0.8);
    CHECK1("tan", geTan, std::tan, 0.4);      CHECK1("tanh", geTanh, std::tanh, 0.6);    CHECK1("sinh", geSinh, std::sinh, 0.6);
    CHECK1("cosh", geCosh, std::cosh, 0.6);   CHECK1("asin", geAsin, std::asin, 0.3);    CHECK1("acos", geAcos, std::acos, 0.3);
    CHECK1("atan", geAtan, std::atan, 0.7);   CHECK1("asinh", geAsinh, std::asinh, 0.7); CHECK1("acosh", geAcosh, std::acosh, 1.8);
    CHECK1("log10", geLog10, std::log10, 1.6);
    // two-argument: pow(dual, const), pow(const, dual), pow(dual, dual); products and quotients of duals
    {
        GDual<double> a(1.4); a.d[0] = 1.0; GDual<double> b(0.7); b.d[1] = 1.0;
        GDual<double> r = gePow(a, b);
        const double h = 1e-6;
        const double fa = (std::pow(1.4 + h, 0.7) - std::pow(1.4 - h, 0.7)) / (2 * h), fb = (std::pow(1.4, 0.7 + h) - std::pow(1.4, 0.7 - h)) / (2 * h);
        int e = std::fabs(r.d[0] - fa) > 1e-7 || std::fabs(r.d[1] - fb) > 1e-7;
        GDual<double> q = a * b / (a + b);
        const double qa = ((1.4 + h) * 0.7 / (1.4 + h + 0.7) - (1.4 - h) * 0.7 / (1.4 - h + 0.7)) / (2 * h);
        e |= std::fabs(q.d[0] - qa) > 1e-7;
        printf("pow/mul/div %s\n", e ? "MISMATCH" : "ok");
        bad += e;
    }
    printf(bad ? "FAILED\n" : "ALL_OK\n");
    return bad ? 1 : 0;
}
