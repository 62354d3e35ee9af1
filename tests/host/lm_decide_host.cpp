// lm_decide_host.cpp -- opt_b200/csrc/lm_decide.cuh (the LM accept / reject / trust-region decision the device-side outer loop runs in
// k_lmDecide) compiled for the CPU.  Reads n cases (HOST_T each: newCost, modelCost, prevCost, radius, decreaseFactor, then the four
// limits) and writes per case: verdict, prevCost, radius, decreaseFactor as doubles.  tests/test_lm_decide_host.py compares with an
// independent transcription of solverGPUGaussNewton.t:1119-1157.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../opt_b200/csrc/lm_decide.cuh"
#ifndef HOST_T
#define HOST_T float
#endif
int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const long n = atol(argv[3]);
    std::vector<HOST_T> in(9 * n);
    FILE* f = fopen(argv[1], "rb"); if (!f || fread(in.data(), sizeof(HOST_T), in.size(), f) != in.size()) return 3; fclose(f);
    std::vector<double> out(4 * n);
    for (long i = 0; i < n; ++i) {
        const HOST_T* c = &in[9 * i];
        optb::LmState<HOST_T> s{ c[2], c[3], c[4] };
        const optb::LmLimits<HOST_T> lim{ c[5], c[6], c[7], c[8] };
        out[4 * i] = (double)optb::lmDecide<HOST_T>(c[0], c[1], s, lim);
        out[4 * i + 1] = (double)s.prevCost; out[4 * i + 2] = (double)s.radius; out[4 * i + 3] = (double)s.decreaseFactor;
    }
    f = fopen(argv[2], "wb"); if (!f) return 4; fwrite(out.data(), sizeof(double), out.size(), f); fclose(f);
    printf("OK %ld\n", n);
    return 0;
}
