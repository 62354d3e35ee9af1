// Host-side run of the per-vertex math of the hand-written robust_nonrigid_alignment kernel set
// (opt_b200/csrc/rna_device.cuh): the same functions the GPU kernels of energy_rna.cu call, looped over the vertices on the
// CPU, with the vertex-CSR lists built by a plain stable counting sort.  Writes J^T F, diag(J^T J), J^T J x (oracle layout:
// Offset, Angle, RobustWeights blocks without padding) and prints cost / model cost for tests/test_rna_host_parity.py.
//   run: rna_host <dataDir> <out.bin> N      (dataDir/p<k>.bin = raw bytes of problemparams[k], dataDir/x.bin = x in the oracle layout)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../opt_b200/csrc/rna_device.cuh"
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
static void csr(int N, int nE, const int* key, const int* val, std::vector<int>& start, std::vector<int>& nbr) {   // stable: keeps the caller's edge order
    start.assign(N + 2, 0); nbr.assign(nE > 0 ? nE : 1, 0);
    for (int e = 0; e < nE; ++e) ++start[key[e] + 1];
    for (int v = 0; v < N; ++v) start[v + 1] += start[v];
    std::vector<int> fill(start.begin(), start.begin() + N);
    for (int e = 0; e < nE; ++e) nbr[fill[key[e]]++] = val[e];
}

int main(int argc, char** argv) {
    if (argc != 4) { fprintf(stderr, "usage: rna_host dataDir out.bin N\n"); return 2; }
    const std::string dir = argv[1];
    std::vector<std::vector<char>> blobs(12);
    for (int k = 0; k <= 10; ++k) blobs[k] = readFile(dir + "/p" + std::to_string(k) + ".bin");
    blobs[11] = readFile(dir + "/x.bin");
    RnaData<T> D{};
    D.N = atoi(argv[3]);
    const long N = D.N;
    D.wf = (T) * (const float*)blobs[0].data(); D.w = (T) * (const float*)blobs[1].data(); D.wc = T(0.1);
    D.Offset = (const T*)blobs[2].data(); D.Angle = (const T*)blobs[3].data(); D.Weights = (const T*)blobs[4].data();
    D.UrShape = (const T*)blobs[5].data(); D.Constraints = (const T*)blobs[6].data(); D.Normals = (const T*)blobs[7].data();
    const int nE = *(const int*)blobs[8].data();
    const int* v0 = (const int*)blobs[9].data(); const int* v1 = (const int*)blobs[10].data();
    auto align64 = [](long n) { return (n + 63) / 64 * 64; };
    D.offA = align64(3 * N); D.offG = D.offA + align64(3 * N);
    const long nPadded = D.offG + align64(N);
    std::vector<int> outStart, outNbr, inStart, inNbr;
    csr(D.N, nE, v0, v1, outStart, outNbr); csr(D.N, nE, v1, v0, inStart, inNbr);
    D.outStart = outStart.data(); D.outNbr = outNbr.data(); D.inStart = inStart.data(); D.inNbr = inNbr.data();
    std::vector<RnaPack<T>> pack(N);
    D.pack = pack.data();
    for (int v = 0; v < D.N; ++v) pack[v] = rnaMakePack(D, v);
    // x: oracle layout -> flat padded layout
    const T* xo = (const T*)blobs[11].data();
    std::vector<T> x(nPadded, T(0));
    for (long i = 0; i < 3 * N; ++i) { x[i] = xo[i]; x[D.offA + i] = xo[3 * N + i]; }
    for (long i = 0; i < N; ++i) x[D.offG + i] = xo[6 * N + i];
    std::vector<double> jtf(7 * N), diag(7 * N), jtjx(7 * N);
    double cost = 0, model = 0;
    for (int v = 0; v < D.N; ++v) {
        cost += (double)rnaCostAt<T, false>(D, v, nullptr);
        model += (double)rnaCostAt<T, true>(D, v, x.data());
        const RnaOut<T> a = rnaGather<T, 1>(D, v, nullptr), b = rnaGather<T, 2>(D, v, nullptr), c = rnaGather<T, 0>(D, v, x.data());
        auto put = [&](std::vector<double>& o, const RnaOut<T>& q) {
            o[3 * v] = q.O.x; o[3 * v + 1] = q.O.y; o[3 * v + 2] = q.O.z;
            o[3 * N + 3 * v] = q.A.x; o[3 * N + 3 * v + 1] = q.A.y; o[3 * N + 3 * v + 2] = q.A.z;
            o[6 * N + v] = q.g;
        };
        put(jtf, a); put(diag, b); put(jtjx, c);
    }
    FILE* out = fopen(argv[2], "wb");
    if (!out) return 2;
    fwrite(jtf.data(), sizeof(double), 7 * N, out); fwrite(diag.data(), sizeof(double), 7 * N, out); fwrite(jtjx.data(), sizeof(double), 7 * N, out);
    fclose(out);
    printf("COST %.17g MODEL %.17g EDGES %d\n", cost, model, nE);
    return 0;
}
