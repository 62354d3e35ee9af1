// Host-side run of the TILE algorithms of the shape_from_shading kernel set (opt_b200/csrc/sfs_tile.cuh): the same functions
// k_sfsInitTiled and k_sfsLoopResident (energy_sfs.cu) call, driven "CTA by CTA, thread by thread" on the CPU with the barriers
// and all-reduces of the kernels replaced by loop boundaries.  tests/test_sfs_host_parity.py compares the outputs with the CPU oracle
// (LM, k PCG steps) and with the per-pixel functions of sfs_device.cuh.
//   run: sfs_tile_host <dataDir> <out.bin> W H tw th NT k lm headInLoop
// out.bin (doubles, N each): r0, CtC, pre, p0 (after the fused PCGInit1 + CtC + FinalizeDiagonal), then delta, r, p, z after k PCG steps.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../opt_b200/csrc/sfs_tile.cuh"
using namespace optb;
#ifndef HOST_T
#define HOST_T double
#endif
typedef HOST_T T;
constexpr int KMAX = 3;

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
    if (argc != 11) { fprintf(stderr, "usage: sfs_tile_host dataDir out.bin W H tw th NT k lm headInLoop\n"); return 2; }
    const std::string dir = argv[1];
    std::vector<std::vector<char>> blobs(21);
    for (int k = 0; k <= 20; ++k) blobs[k] = readFile(dir + "/p" + std::to_string(k) + ".bin");
    auto fl = [&](int k) { return (T) * (const float*)blobs[k].data(); };
    SfsData<T> P{};
    P.W = atoi(argv[3]); P.H = atoi(argv[4]);
    const int tw = atoi(argv[5]), th = atoi(argv[6]), NT = atoi(argv[7]), K = atoi(argv[8]), lm = atoi(argv[9]), headInLoop = atoi(argv[10]);
    const long N = (long)P.W * P.H;
    P.wp = std::sqrt(fl(0)); P.ws = std::sqrt(fl(1)); P.wg = std::sqrt(fl(2));      // same binding as energy_sfs.cu
    if (sizeof(T) == 4) { P.wp = sqrtf(*(const float*)blobs[0].data()); P.ws = sqrtf(*(const float*)blobs[1].data()); P.wg = sqrtf(*(const float*)blobs[2].data()); }
    P.fx = fl(3); P.fy = fl(4); P.ux = fl(5); P.uy = fl(6);
    for (int i = 1; i <= 9; ++i) P.L[i] = fl(6 + i);
    P.X = (const T*)blobs[16].data(); P.D = (const T*)blobs[17].data(); P.Im = (const T*)blobs[18].data();
    P.eR = (const unsigned char*)blobs[19].data(); P.eC = (const unsigned char*)blobs[20].data();
    std::vector<T> B(N), dB(3 * N), cR(5 * N), cC(5 * N); std::vector<unsigned char> flags(N);
    P.B = B.data(); P.dB = dB.data(); P.cR = cR.data(); P.cC = cC.data(); P.fl = flags.data();
    for (int y = 0; y < P.H; ++y) for (int x = 0; x < P.W; ++x) sfsPrecomputeB(P, x, y);
    for (int y = 0; y < P.H; ++y) for (int x = 0; x < P.W; ++x) sfsBuildRows(P, x, y);

    // ------------------------------------------------------------------ precompute + cost in one tiled pass against the two per-pixel passes above
    double preErr = 0.0, costTiled = 0.0, costRef = 0.0;
    {
        std::vector<T> B2(N, T(-5)), cR2(5 * N, T(-5)), cC2(5 * N, T(-5)); std::vector<unsigned char> fl2(N, 255);
        SfsData<T> Q = P; Q.B = B2.data(); Q.cR = cR2.data(); Q.cC = cC2.data(); Q.fl = fl2.data(); Q.dB = nullptr;
        const int itw = 32, ith = 8, block = 256;
        const int tX = (P.W + itw - 1) / itw, tY = (P.H + ith - 1) / ith;
        std::vector<unsigned char> planes(sfsPrePlaneBytes<T>(itw, ith) + 64);
        for (int tile = 0; tile < tX * tY; ++tile) {
            SfsPreView<T> V = sfsCarvePre<T>(planes.data(), (tile % tX) * itw, (tile / tX) * ith, itw, ith);
            for (int tid = 0; tid < block; ++tid) for (int c = tid; c < V.bcells(); c += block) sfsPreCell(Q, V, c);
            for (int tid = 0; tid < block; ++tid) {
                const int tx = tid % itw, ty = tid / itw;
                if (V.x0 + tx < P.W && V.y0 + ty < P.H) costTiled += (double)sfsPrePixel(Q, V, ty, tx);
            }
        }
        for (long i = 0; i < N; ++i) {
            preErr = std::fmax(preErr, std::fabs((double)B2[i] - (double)B[i]));
            if (fl2[i] != flags[i]) preErr = 1e30;
            for (int k = 0; k < 5; ++k) { preErr = std::fmax(preErr, std::fabs((double)cR2[k * N + i] - (double)cR[k * N + i])); preErr = std::fmax(preErr, std::fabs((double)cC2[k * N + i] - (double)cC[k * N + i])); }
        }
        for (int y = 0; y < P.H; ++y) for (int x = 0; x < P.W; ++x) costRef += (double)sfsCostAt<T, false>(P, x, y, nullptr);
    }

    // ------------------------------------------------------------------ fused PCGInit1 (+ CtC + FinalizeDiagonal), 32 x 8 tiles of 256 "threads"
    std::vector<T> delta(N, T(7)), r(N, T(7)), pre(N, T(7)), p(N, T(7)), CtC(N, T(7)), SSq(N, T(7)), b(N, T(7)), z(N, T(0)), Ap(N, T(0));
    const T radius = T(1e4f), minDiag = T(1e-6f), maxDiag = T(1e32f), qTol = T(0.0001f);   // solver defaults (solverGPUGaussNewton.t:148-163)
    double rz0 = 0.0;
    double maxJtf = 0.0, maxDiagErr = 0.0;
    {
        const int itw = 32, ith = 8, block = 256;
        const int tilesX = (P.W + itw - 1) / itw, tilesY = (P.H + ith - 1) / ith;
        std::vector<unsigned char> planes(sfsTilePlaneBytes<T>(itw, ith) + 64);
        SfsInitOut<T> O{ delta.data(), r.data(), pre.data(), p.data(), CtC.data(), SSq.data(), b.data() };
        for (int tile = 0; tile < tilesX * tilesY; ++tile) {
            SfsTileView<T> V = sfsCarveTile<T>(planes.data(), (tile % tilesX) * itw, (tile / tilesX) * ith, itw, ith);
            for (int tid = 0; tid < block; ++tid) {
                for (int i = tid; i < V.PW() + V.PH(); i += block) sfsTileFactors(P, V, i);
                for (int c = tid; c < V.cells(); c += block) sfsTileStageCell(P, V, c);
            }
            for (int tid = 0; tid < block; ++tid) for (int c = tid; c < V.cells(); c += block) sfsTileRowResiduals(P, V, c);
            for (int tid = 0; tid < block; ++tid) {
                const int tx = tid % itw, ty = tid / itw, px = V.x0 + tx, py = V.y0 + ty;
                if (px >= P.W || py >= P.H) continue;
                const long u = (long)py * P.W + px;
                const bool inc = (V.sfl[(ty + 1) * V.RW() + tx + 1] & kSfsIncluded) != 0;
                T jtf = T(0), dg = T(0);
                if (inc) { jtf = sfsTileGather(P, V, ty, tx, P.wp * sfsFP(P, u)); dg = sfsTileGatherDiag(P, V, ty, tx); }
                const double e1 = std::fabs((double)jtf - (double)sfsGather<T, 1>(P, px, py, nullptr)), e2 = std::fabs((double)dg - (double)sfsGather<T, 2>(P, px, py, nullptr));
                if (e1 > maxJtf) maxJtf = e1;
                if (e2 > maxDiagErr) maxDiagErr = e2;
                rz0 += lm ? (double)sfsInitPixel<T, true>(inc, jtf, dg, u, O, 0, 1, radius, minDiag, maxDiag)
                          : (double)sfsInitPixel<T, false>(inc, jtf, dg, u, O, 0, 1, radius, minDiag, maxDiag);
            }
        }
    }
    FILE* out = fopen(argv[2], "wb");
    if (!out) return 2;
    auto dump = [&](const std::vector<T>& v) { std::vector<double> d(v.begin(), v.end()); fwrite(d.data(), sizeof(double), N, out); };
    dump(r); if (lm) dump(CtC); else dump(pre); dump(pre); dump(p);

    // ------------------------------------------------------------------ whole PCG loop, tiles of tw x th, NT "threads" per CTA
    const int tilesX = (P.W + tw - 1) / tw, tilesY = (P.H + th - 1) / th, nCta = tilesX * tilesY;
    if (tw * th > KMAX * NT) { fprintf(stderr, "tile too large for KMAX * NT\n"); return 2; }
    std::vector<std::vector<unsigned char>> planes(nCta, std::vector<unsigned char>(sfsTilePlaneBytes<T>(tw, th, true) + 64));
    std::vector<SfsTileView<T>> V(nCta);
    std::vector<std::vector<SfsLoopRegs<T, KMAX>>> R(nCta, std::vector<SfsLoopRegs<T, KMAX>>(NT));
    std::vector<unsigned long long> zLL(2 * N, 0ull), dLL(2 * N, 0ull);   // tile-edge words (value, tag)
    unsigned long long kRed = 0;                                         // all-reduces completed so far (the kernel continues the plan's sequence)
    bool tailPhase = false;
    std::vector<T> SSqLoop(N, T(7));
    double rzHead = 0.0;
    for (int cta = 0; cta < nCta; ++cta) {
        V[cta] = sfsCarveTile<T>(planes[cta].data(), (cta % tilesX) * tw, (cta / tilesX) * th, tw, th, true);
        for (int tid = 0; tid < NT; ++tid) {
            for (int i = tid; i < V[cta].PW() + V[cta].PH(); i += NT) sfsTileFactors(P, V[cta], i);
            for (int c = tid; c < V[cta].cells(); c += NT) sfsTileStageCell(P, V[cta], c);
            if (!headInLoop) sfsLoopLoad<T, KMAX>(P, V[cta], R[cta][tid], tid, NT, delta.data(), r.data(), b.data(), CtC.data(), pre.data(), p.data(), 0, lm);
            else for (int c = tid; c < V[cta].PW() * V[cta].PH(); c += NT) { V[cta].sp[c] = T(0); V[cta].sd[c] = T(0); }
        }
    }
    if (headInLoop) {   // the LM head inside the loop kernel: residual rows, head per pixel (direction published on the tile edges), ring, r . p
        for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) for (int c = tid; c < V[cta].cells(); c += NT) sfsTileRowResiduals(P, V[cta], c);
        for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid)
            rzHead += (double)sfsLoopHead<T, KMAX>(P, V[cta], R[cta][tid], tid, NT, 0, 1, radius, minDiag, maxDiag, SSqLoop.data(), zLL.data(), sfsEdgeTag(kRed, kSfsEdgeFirstDirection));
        for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) sfsLoopFetchRing<T>(P, V[cta], tid, NT, V[cta].sp, zLL.data(), sfsEdgeTag(kRed, kSfsEdgeFirstDirection));
        ++kRed;
    }
    double dRz = headInLoop ? rzHead : rz0, dRzPrev = 0.0, dPAp = 0.0, dQ = 0.0;
    T Q0 = T(0);
    int done = 0, broke = -1;
    const int resetPeriod = 10;   // residual_reset_period, solver default
    for (int l = 0; l < K; ++l) {
        if (l > 0) {
            T beta = T(0);
            { const T rn = (T)dRz, ro = (T)dRzPrev; if (ro > T(0)) beta = rn / ro; }
            for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) sfsLoopUpdateDirection<T, KMAX>(P, V[cta], R[cta][tid], tid, NT, beta, zLL.data(), sfsEdgeTag(kRed - 1, 1));
        }
        for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) for (int c = tid; c < V[cta].cells(); c += NT) sfsTileRowsTimesVec(P, V[cta], c, V[cta].sp);
        double s = 0.0;
        for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) s += (double)sfsLoopApply<T, KMAX>(P, V[cta], R[cta][tid], tid, NT);
        dPAp = s; ++kRed;
        T alpha = T(0);
        { const T num = (T)dRz, den = (T)dPAp; if (den > T(0)) alpha = num / den; }
        double s0 = 0.0, s1 = 0.0;
        // (the kernel's z stores of step l land before any ring read of step l + 1: separated here by the loop structure)
        if (lm && ((l + 1) % resetPeriod) == 0) {
            for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) sfsLoopResetFirstHalf<T, KMAX>(P, V[cta], R[cta][tid], tid, NT, alpha, dLL.data(), sfsEdgeTag(kRed, 2));
            for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) sfsLoopFetchRing<T>(P, V[cta], tid, NT, V[cta].sd, dLL.data(), sfsEdgeTag(kRed, tailPhase ? 3 : 2));
            for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) for (int c = tid; c < V[cta].cells(); c += NT) sfsTileRowsTimesVec(P, V[cta], c, V[cta].sd);
            for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) {
                T a0 = T(0), a1 = T(0);
                sfsLoopResetSecondHalf<T, KMAX>(P, V[cta], R[cta][tid], tid, NT, 0, zLL.data(), sfsEdgeTag(kRed, 1), a0, a1);
                s0 += (double)a0; s1 += (double)a1;
            }
        } else
        for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) {
            T a0 = T(0), a1 = T(0);
            sfsLoopStep2<T, KMAX>(P, V[cta], R[cta][tid], tid, NT, alpha, 0, zLL.data(), sfsEdgeTag(kRed, 1), a0, a1);
            s0 += (double)a0; s1 += (double)a1;
        }
        dRzPrev = dRz; dRz = s0; dQ = s1; ++kRed;
        done = l + 1;
        if (lm) {
            const T Q1 = (T)dQ;
            const T zeta = T(l + 1) * (Q1 - Q0) / Q1;
            Q0 = Q1;
            if (zeta < qTol) { broke = l; break; }
        }
    }
    double model = 0.0;
    tailPhase = true;
    if (lm) {   // computeModelCost inside the loop kernel
        for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) sfsLoopPublishDelta<T, KMAX>(P, V[cta], R[cta][tid], tid, NT, dLL.data(), sfsEdgeTag(kRed, 3));
        for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) sfsLoopFetchRing<T>(P, V[cta], tid, NT, V[cta].sd, dLL.data(), sfsEdgeTag(kRed, tailPhase ? 3 : 2));
        for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) for (int c = tid; c < V[cta].cells(); c += NT) sfsTileRowsTimesVec(P, V[cta], c, V[cta].sd);
        for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid) model += (double)sfsLoopModelCost<T, KMAX>(P, V[cta], R[cta][tid], tid, NT);
    }
    T beta = T(0);
    { const T rn = (T)dRz, ro = (T)dRzPrev; if (ro > T(0)) beta = rn / ro; }
    // poison the vectors first: the epilogue must rewrite every in-image entry
    std::vector<T> z2(N, T(-777)), delta2(N, T(-777)), r2(N, T(-777)), p2(N, T(-777)), Ap2(N, T(-777));
    std::vector<T> pre2(N, T(-777)), ctc2(N, T(-777)), b2(N, T(-777));
    for (int cta = 0; cta < nCta; ++cta) for (int tid = 0; tid < NT; ++tid)
        sfsLoopStore<T, KMAX>(P, V[cta], R[cta][tid], tid, NT, done > 0, beta, delta2.data(), r2.data(), z2.data(), p2.data(), Ap2.data(), (T*)nullptr, (T*)nullptr,
                              headInLoop ? pre2.data() : nullptr, ctc2.data(), b2.data());
    double headErr = 0.0;   // head inside the loop against the stand-alone fused head (same functions, same order: identical)
    if (headInLoop) {
        for (long i = 0; i < N; ++i) {
            headErr = std::fmax(headErr, std::fabs((double)pre2[i] - (double)pre[i])); headErr = std::fmax(headErr, std::fabs((double)ctc2[i] - (double)CtC[i]));
            headErr = std::fmax(headErr, std::fabs((double)b2[i] - (double)b[i])); headErr = std::fmax(headErr, std::fabs((double)SSqLoop[i] - (double)SSq[i]));
        }
        headErr = std::fmax(headErr, std::fabs(rzHead - rz0));
    }
    dump(delta2); dump(r2); dump(p2); dump(z2);
    fclose(out);
    double modelRef = 0.0;   // the per-pixel form (sfs_device.cuh), itself checked against the oracle by test_sfs_per_pixel_math_matches_oracle
    for (int y = 0; y < P.H; ++y) for (int x = 0; x < P.W; ++x) modelRef += (double)sfsCostAt<T, true>(P, x, y, delta2.data());
    printf("RZ0 %.17g RZ %.17g PAP %.17g Q %.17g DONE %d BROKE %d JTFERR %.3g DIAGERR %.3g CTAS %d MODEL %.17g MODELREF %.17g PREERR %.3g COST %.17g COSTREF %.17g HEADERR %.3g\n", rz0, dRz, dPAp, dQ, done, broke, maxJtf, maxDiagErr, nCta,
           model, modelRef, preErr, costTiled, costRef, headErr);
    return 0;
}
