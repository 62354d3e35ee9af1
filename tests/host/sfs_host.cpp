// Host-side run of the per-pixel math of the hand-written shape_from_shading kernel set (opt_b200/csrc/sfs_device.cuh):
// the same functions the GPU kernels of energy_sfs.cu call, looped over the image on the CPU.  Writes cost, J^T F,
// diag(J^T J), J^T J x and the model cost for x, which tests/test_sfs_host_parity.py compares with the CPU oracle.
//   run: sfs_host <dataDir> <out.bin> W H      (dataDir/p<k>.bin = raw bytes of problemparams[k], dataDir/x.bin = the vector x)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../opt_b200/csrc/sfs_device.cuh"
using namespace optb;
#ifndef HOST_T
#define HOST_T double
#endif
typedef HOST_T T;

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
    if (argc != 5) { fprintf(stderr, "usage: sfs_host dataDir out.bin W H\n"); return 2; }
    const std::string dir = argv[1];
    std::vector<std::vector<char>> blobs(22);
    for (int k = 0; k <= 20; ++k) blobs[k] = readFile(dir + "/p" + std::to_string(k) + ".bin");
    blobs[21] = readFile(dir + "/x.bin");
    auto fl = [&](int k) { return (T) * (const float*)blobs[k].data(); };
    SfsData<T> P{};
    P.W = atoi(argv[3]); P.H = atoi(argv[4]);
    const long N = (long)P.W * P.H;
    P.wp = std::sqrt(fl(0)); P.ws = std::sqrt(fl(1)); P.wg = std::sqrt(fl(2));      // same binding as energy_sfs.cu
    P.fx = fl(3); P.fy = fl(4); P.ux = fl(5); P.uy = fl(6);
    for (int i = 1; i <= 9; ++i) P.L[i] = fl(6 + i);
    P.X = (const T*)blobs[16].data(); P.D = (const T*)blobs[17].data(); P.Im = (const T*)blobs[18].data();
    P.eR = (const unsigned char*)blobs[19].data(); P.eC = (const unsigned char*)blobs[20].data();
    std::vector<T> B(N), dB(3 * N), cR(5 * N), cC(5 * N); std::vector<unsigned char> flags(N);
    P.B = B.data(); P.dB = dB.data(); P.cR = cR.data(); P.cC = cC.data(); P.fl = flags.data();
    for (int y = 0; y < P.H; ++y) for (int x = 0; x < P.W; ++x) sfsPrecomputeB(P, x, y);
    for (int y = 0; y < P.H; ++y) for (int x = 0; x < P.W; ++x) sfsBuildRows(P, x, y);
    const T* xv = (const T*)blobs[21].data();
    std::vector<T> xm(N);                                                       // the oracle zeroes excluded entries of x
    for (long i = 0; i < N; ++i) xm[i] = (flags[i] & kSfsIncluded) ? xv[i] : T(0);
    std::vector<double> jtf(N), diag(N), jtjx(N);
    double cost = 0, model = 0;
    for (int y = 0; y < P.H; ++y) for (int x = 0; x < P.W; ++x) {
        const long c = (long)y * P.W + x;
        cost += (double)sfsCostAt<T, false>(P, x, y, nullptr);
        model += (double)sfsCostAt<T, true>(P, x, y, xm.data());
        jtf[c] = (double)sfsGather<T, 1>(P, x, y, nullptr);
        diag[c] = (double)sfsGather<T, 2>(P, x, y, nullptr);
        jtjx[c] = (double)sfsGather<T, 0>(P, x, y, xm.data());
    }
    FILE* out = fopen(argv[2], "wb");
    if (!out) return 2;
    fwrite(jtf.data(), sizeof(double), N, out); fwrite(diag.data(), sizeof(double), N, out); fwrite(jtjx.data(), sizeof(double), N, out);
    fclose(out);
    long nFlat = 0, nInc = 0; for (long i = 0; i < N; ++i) { nFlat += (flags[i] & kSfsFlat) ? 1 : 0; nInc += (flags[i] & kSfsIncluded) ? 1 : 0; }
    printf("COST %.17g MODEL %.17g INCLUDED %ld FLAT %ld\n", cost, model, nInc, nFlat);
    return 0;
}
