// common.cuh -- shared device/host helpers for libOpt.so (sm_100a).
//
// Replaces the reference's GPU runtime utilities (API/src/util.t): error checking
// (`cd`/`checkedLaunch`, util.t:739-768), launch geometry (util.t:782-843) and the reduction
// primitive (`warpReduce` + one global atomicAdd per warp, util.t:526-623 /
// solverGPUGaussNewton.t:312-317).  The reduction here is different by design: every CTA writes ONE
// double partial, the last CTA to finish sums the partials in index order, so results are
// deterministic run-to-run and no same-address float atomics are issued.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define OPT_CUDA_CHECK(call)                                                                       \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            fprintf(stderr, "Cuda reported error %d: %s\nIn call: %s\nIn file: %s:%d\n", (int)_e,  \
                    cudaGetErrorString(_e), #call, __FILE__, __LINE__);                            \
            exit((int)_e); /* same contract as the reference: print and exit(code), util.t:739-752 */ \
        }                                                                                          \
    } while (0)

namespace optb {

constexpr int kBlock = 256;          // threads per CTA for all solver kernels
constexpr int kMaxPartials = 4096;   // upper bound on CTAs that contribute to one reduction

// Device-resident scalars of one plan (the reference keeps 6 separate 1-element allocations,
// solverGPUGaussNewton.t:1274-1280).  Kept in double; kernels cast to opt_float on use.
enum Scalar : int {
    kRz0 = 0,       // r.z ping  (scanAlphaNumerator / scanBetaNumerator alternate by PCG-iteration parity)
    kRz1 = 1,       // r.z pong
    kPAp = 2,       // p.Ap      (scanAlphaDenominator)
    kCost = 3,      // scratch   (cost)
    kModelCost = 4, // modelCost
    kQ = 5,         // q (LM)
    kQPrev = 6,     // previous q (LM zeta test on device)
    kFlags = 7,     // device-side control word (LM early-out), stored as a double 0/1
    kRadius = 10,   // LM trust-region radius of the current nonlinear iteration (host decides, kernels read it here: the iteration graph stays the same)
    // LM decision on the device (OPTB_DEVICE_LM=1, lm_decide.cuh / k_lmDecide in solver.cu): the nonlinear state the host keeps otherwise.
    // All zero unless that mode runs, so a kernel that tests kLmGate / kLmReject behaves as before everywhere else.
    kLmPrevCost = 11, // cost of the last accepted iterate
    kLmFactor = 12,   // radius_decrease_factor
    kLmGate = 13,     // != 0: the solve has ended (LmVerdict that ended it); gated kernels of the remaining enqueued iterations return at once
    kLmReject = 14,   // 1: the iteration just decided was rejected (revert the unknowns, redo precompute)
    kLmIters = 15,    // nonlinear iterations completed on the device since the state was uploaded
    kNumScalars = 16
};

constexpr int kMaxRanks = 8;

// Wait-loop watchdog.  A waiter that never sees its word is a bug (a lost CTA / rank), and a hang would cost the whole job, so the
// loops give up with __trap() -- but by WALL CLOCK, not by poll count: on a time-sliced GPU (MPS off, two processes sharing a device --
// the multi-rank tests do exactly that) or under a debugger a peer can be descheduled for many milliseconds, which a poll-count limit
// (2^24 polls ~ 1 s) turned into a spurious abort.  %globaltimer is read once per 4096 polls; the limit is 30 s.
__device__ __forceinline__ unsigned long long globalTimerNs() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
struct SpinGuard {
    unsigned long long spins = 0, t0 = 0;
    __device__ __forceinline__ void tick() {
        if ((++spins & 4095ull) == 0ull) {
            const unsigned long long now = globalTimerNs();
            if (t0 == 0ull) t0 = now; else if (now - t0 > 30000000000ull) __trap();
        }
    }
};


// Cross-GPU view of the scalar mailboxes ("a single allreduce of the PCG scalars").
// Every rank owns one mailbox block (cudaMalloc + CUDA IPC); peerWords[r] is rank r's block mapped into
// this process.  A reduction's last CTA stores its local sum into slot [scalar][epoch&1][myRank] of EVERY
// rank's block over NVLink; consumers spin on their OWN block until all `world` entries carry the epoch
// and add them in rank order, so every rank obtains the bitwise-identical sum with no host involvement
// and no NCCL launch.  Two parities suffice: a rank cannot produce epoch e+2 of a scalar before every rank
// has consumed epoch e (it needs every rank's epoch e+1 contribution first).
// Wire format ("LL", as in NCCL's low-latency protocol): a double travels as two 8-byte words, each
// carrying 32 payload bits and the 32-bit epoch, so an 8-byte store is self-validating -- no fence and no
// separate flag write, one NVLink one-way latency per reduction.
struct CommView {
    int world, rank;
    unsigned long long* const* peerWords;           // device table [world]: rank r's mailbox [kNumScalars][2 parities][kMaxRanks senders][2 words]
                                                    // (a table in global memory, not an array in the kernel parameters: run-time indexing of a
                                                    // by-value parameter array forces a per-thread local-memory copy of the whole struct)
    unsigned long long* epoch;                      // local [kNumScalars]: epochs produced so far
};

struct ReduceWorkspace {
    double* partials;        // [kNumScalars][kMaxPartials]
    unsigned int* counters;  // [kNumScalars]  arrival counters, self-resetting
    double* scalars;         // [kNumScalars]   (single GPU: the reduced values)
    CommView comm;           // world == 1: unused
};

__device__ __forceinline__ unsigned long long ldAcquireSys(const unsigned long long* p) {
    unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void stReleaseSys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ldRelaxedSysF64(const double* p) {
    double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v;
}

__device__ __forceinline__ unsigned long long ldVolatileU64(const unsigned long long* p) {
    unsigned long long v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void stVolatileU64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ size_t mailboxWord(int slot, unsigned long long e, int sender) {
    return (((size_t)slot * 2 + (size_t)(e & 1ull)) * kMaxRanks + sender) * 2;
}
// last CTA of a reduction: publish the local sum `u` of `slot` to every rank (including this one)
__device__ __forceinline__ void publishScalar(const ReduceWorkspace& ws, int slot, double u) {
    const unsigned long long e = ws.comm.epoch[slot] + 1ull;
    ws.comm.epoch[slot] = e;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(u);
    const unsigned long long w0 = ((bits & 0xffffffffull) << 32) | (e & 0xffffffffull);
    const unsigned long long w1 = ((bits >> 32) << 32) | (e & 0xffffffffull);
    const size_t at = mailboxWord(slot, e, ws.comm.rank);
    for (int r = 0; r < ws.comm.world; ++r) { stVolatileU64(ws.comm.peerWords[r] + at, w0); stVolatileU64(ws.comm.peerWords[r] + at + 1, w1); }
}
// Same, issued by one full warp whose lanes all hold `u`: lane (2*rank + word) stores one word, so the 2*world NVLink
// stores leave together.
__device__ __forceinline__ void publishScalarWarp(const ReduceWorkspace& ws, int slot, double u) {
    const int lane = threadIdx.x & 31;
    const unsigned long long e = ws.comm.epoch[slot] + 1ull;
    __syncwarp();
    if (lane == 0) ws.comm.epoch[slot] = e;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(u);
    if (lane < 2 * ws.comm.world) {
        const unsigned long long w = (((lane & 1) ? (bits >> 32) : (bits & 0xffffffffull)) << 32) | (e & 0xffffffffull);
        stVolatileU64(ws.comm.peerWords[lane >> 1] + mailboxWord(slot, e, ws.comm.rank) + (lane & 1), w);
    }
}
// Value of reduction `slot` as seen by every rank (single thread).  All 2*world words of the (slot, parity) row are
// contiguous; they are requested together and re-polled together, so the wait costs one L2 round trip per poll, not
// one per rank (polling rank after rank cost 8 x 0.7 us per scalar on 8 GPUs).
__device__ __forceinline__ double loadScalarGlobal(const ReduceWorkspace& ws, int slot) {
    if (ws.comm.world <= 1) return ws.scalars[slot];
    const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(ws.comm.epoch + slot);
    if (e == 0ull) return 0.0;     // never produced
    const unsigned long long want = e & 0xffffffffull;
    const unsigned long long* row = ws.comm.peerWords[ws.comm.rank] + mailboxWord(slot, e, 0);
    const int nWords = 2 * ws.comm.world;
    unsigned long long w[2 * kMaxRanks];
    bool ok; SpinGuard guard;
    do {
        ok = true; guard.tick();
#pragma unroll
        for (int i = 0; i < 2 * kMaxRanks; ++i) { w[i] = want; if (i < nWords) w[i] = ldVolatileU64(row + i); }
#pragma unroll
        for (int i = 0; i < 2 * kMaxRanks; ++i) ok = ok && ((w[i] & 0xffffffffull) == want);
    } while (!ok);
    double sum = 0.0;
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)   // fixed rank order => bitwise identical on every rank
        if (r < ws.comm.world) sum += __longlong_as_double((long long)((w[2 * r] >> 32) | ((w[2 * r + 1] >> 32) << 32)));
    return sum;
}
// Sequence number of the halo exchange that follows the most recent r.z production (PCGInit1 / PCGStep2):
// identical on all ranks, strictly increasing from one exchange to the next, stable while a kernel runs.
__device__ __forceinline__ unsigned long long haloEpochNow(const ReduceWorkspace& ws) {
    return *reinterpret_cast<volatile unsigned long long*>(ws.comm.epoch + kRz0) + *reinterpret_cast<volatile unsigned long long*>(ws.comm.epoch + kRz1);
}
// Block-wide: returns the globally reduced values of two slots to every thread.  Warp 0 polls all 2 x 2*world mailbox
// words at once (lane = slot half | sender | word) and every lane adds the payloads in rank order.
__device__ __forceinline__ void loadScalars2(const ReduceWorkspace& ws, int slotA, int slotB, double& a, double& b) {
    if (ws.comm.world <= 1) { a = ws.scalars[slotA]; b = ws.scalars[slotB]; return; }
    __shared__ double s_ab[2];
    static_assert(kMaxRanks == 8, "lane mapping below assumes 16 words per slot");
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x, idx = lane & 15, sender = idx >> 1, word = idx & 1;
        const int slot = (lane & 16) ? slotB : slotA;
        const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(ws.comm.epoch + slot);
        unsigned long long w = 0ull;
        if (e != 0ull && sender < ws.comm.world) {
            const unsigned long long want = e & 0xffffffffull;
            const unsigned long long* src = ws.comm.peerWords[ws.comm.rank] + mailboxWord(slot, e, sender) + word;
            SpinGuard guard;
            do { w = ldVolatileU64(src); guard.tick(); } while ((w & 0xffffffffull) != want);
        }
        __syncwarp();
        const unsigned payload = (unsigned)(w >> 32);   // 0 for idle lanes and for a slot that was never produced
        double sum = 0.0;
        for (int r = 0; r < ws.comm.world; ++r) {       // fixed rank order => bitwise identical on every rank
            const unsigned lo = __shfl_sync(0xffffffffu, payload, (lane & 16) + 2 * r), hi = __shfl_sync(0xffffffffu, payload, (lane & 16) + 2 * r + 1);
            sum += __longlong_as_double((long long)((unsigned long long)lo | ((unsigned long long)hi << 32)));
        }
        if (lane == 0) s_ab[0] = sum;
        if (lane == 16) s_ab[1] = sum;
    }
    __syncthreads();
    a = s_ab[0]; b = s_ab[1];
    __syncthreads();
}

__device__ __forceinline__ double warpSum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Thread-strided sum of n per-CTA partials (thread t adds elements t, t+NT, t+2NT, ... in that order, so the result
// does not depend on the unrolling).  The loads bypass L1 (the partials were written by other SMs) and are issued four
// at a time: one L2 round trip per four elements instead of one per element in the critical tail of every reduction.
template <int NT>
__device__ __forceinline__ double orderedStridedSum(const double* part, unsigned n) {
    double t = 0.0;
    unsigned i = threadIdx.x;
    for (; i + 3u * NT < n; i += 4u * NT) {
        const double a0 = __ldcg(part + i), a1 = __ldcg(part + i + NT), a2 = __ldcg(part + i + 2 * NT), a3 = __ldcg(part + i + 3 * NT);
        t += a0; t += a1; t += a2; t += a3;
    }
    for (; i < n; i += NT) t += __ldcg(part + i);
    return t;
}

// Block-wide sum of `v` (every thread calls), then grid-wide deterministic finish:
// partial -> partials[slot][blockIdx]; last CTA sums gridDim.x partials in order -> scalars[slot].
// Must be called by all threads of all CTAs of the launch exactly once per slot.
template <int NT>
__device__ __forceinline__ void gridReduceTo(double v, const ReduceWorkspace& ws, int slot, unsigned long long* const* releaseFlags = nullptr) {
    // releaseFlags != nullptr (multi-GPU, the kernel stored halo data into neighbours' memory over NVLink before calling this): every
    // CTA's arrival is preceded by a SYSTEM-scope fence, and the last CTA -- which has then observed every CTA's arrival -- releases
    // releaseFlags[0..1] (the neighbours' "halo arrived" words, nullptr = no neighbour) with the halo sequence number of this step.
    __shared__ double s_warp[NT / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warpSum(v);
    if (lane == 0) s_warp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = lane < NT / 32 ? s_warp[lane] : 0.0;
        t = warpSum(t);
        if (lane == 0) {
            double* part = ws.partials + (size_t)slot * kMaxPartials;
            part[blockIdx.x] = t;
            if (releaseFlags) __threadfence_system(); else __threadfence();
            unsigned int prev = atomicAdd(ws.counters + slot, 1u);
            s_last = (prev == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        const double* part = ws.partials + (size_t)slot * kMaxPartials;
        const double t0 = orderedStridedSum<NT>(part, gridDim.x);
        double t = t0;
        // fixed summation tree: thread-strided partial sums, then the same shuffle tree every run
        t = warpSum(t);
        if (lane == 0) s_warp[wid] = t;
        __syncthreads();
        if (wid == 0) {
            double u = lane < NT / 32 ? s_warp[lane] : 0.0;
            u = warpSum(u);   // every lane holds the total
            if (lane == 0) {
                ws.counters[slot] = 0u;
                if (ws.comm.world <= 1) ws.scalars[slot] = u;
            }
            if (ws.comm.world > 1) publishScalarWarp(ws, slot, u);
            if (releaseFlags) {
                __syncwarp();
                if (lane == 0) {
                    __threadfence_system();
                    const unsigned long long e = haloEpochNow(ws);
                    for (int d = 0; d < 2; ++d) if (releaseFlags[d]) stReleaseSys(releaseFlags[d], e);
                }
            }
        }
    }
    __syncthreads();   // s_warp / s_last reusable by a second call in the same kernel
}

// Two sums with ONE arrival round (PCGStep2 of LM produces r.z and q together): same per-slot partial layout and summation order as
// two gridReduceTo calls, so the values are bitwise the same; returns true in every thread of the CTA that finished the reduction, whose
// thread 0 also receives the two LOCAL sums (single GPU: the reduced values).
template <int NT>
__device__ __forceinline__ bool gridReduceTo2(double v0, double v1, const ReduceWorkspace& ws, int slot0, int slot1, unsigned long long* const* releaseFlags,
                                              double* out0, double* out1) {
    __shared__ double s_warp2[2][NT / 32];
    __shared__ bool s_last2;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v0 = warpSum(v0); v1 = warpSum(v1);
    if (lane == 0) { s_warp2[0][wid] = v0; s_warp2[1][wid] = v1; }
    __syncthreads();
    if (wid == 0) {
        double t0 = lane < NT / 32 ? s_warp2[0][lane] : 0.0, t1 = lane < NT / 32 ? s_warp2[1][lane] : 0.0;
        t0 = warpSum(t0); t1 = warpSum(t1);
        if (lane == 0) {
            ws.partials[(size_t)slot0 * kMaxPartials + blockIdx.x] = t0;
            ws.partials[(size_t)slot1 * kMaxPartials + blockIdx.x] = t1;
            if (releaseFlags) __threadfence_system(); else __threadfence();
            unsigned int prev = atomicAdd(ws.counters + slot0, 1u);
            s_last2 = (prev == gridDim.x - 1);
        }
    }
    __syncthreads();
    const bool last = s_last2;
    if (last) {
        __threadfence();
        double t0 = orderedStridedSum<NT>(ws.partials + (size_t)slot0 * kMaxPartials, gridDim.x);
        double t1 = orderedStridedSum<NT>(ws.partials + (size_t)slot1 * kMaxPartials, gridDim.x);
        t0 = warpSum(t0); t1 = warpSum(t1);
        if (lane == 0) { s_warp2[0][wid] = t0; s_warp2[1][wid] = t1; }
        __syncthreads();
        if (wid == 0) {
            double u0 = lane < NT / 32 ? s_warp2[0][lane] : 0.0, u1 = lane < NT / 32 ? s_warp2[1][lane] : 0.0;
            u0 = warpSum(u0); u1 = warpSum(u1);
            if (lane == 0) {
                ws.counters[slot0] = 0u;
                if (ws.comm.world <= 1) { ws.scalars[slot0] = u0; ws.scalars[slot1] = u1; }
                if (out0) *out0 = u0; if (out1) *out1 = u1;
            }
            if (ws.comm.world > 1) { publishScalarWarp(ws, slot0, u0); publishScalarWarp(ws, slot1, u1); }
            if (releaseFlags) {
                __syncwarp();
                if (lane == 0) {
                    __threadfence_system();
                    const unsigned long long e = haloEpochNow(ws);
                    for (int d = 0; d < 2; ++d) if (releaseFlags[d]) stReleaseSys(releaseFlags[d], e);
                }
            }
        }
    }
    __syncthreads();
    return last;
}

// ------------------------------------------------------------------ persistent-kernel reductions
// State of the grid-wide all-reduce used by the persistent PCG-loop kernels (cooperative launch: all CTAs are
// co-resident).  One object per plan; `k` is a sequence number that only ever grows (the host hands every launch a
// fresh base), so a stale flag can never be mistaken for a new one.
struct LoopSync { unsigned long long flag[2]; double value[2]; unsigned int counter; unsigned int pad; };

// Every thread of every CTA calls this with the same k.  Returns the sum of `v` over the whole grid to every thread,
// and is a grid-wide barrier: global stores issued (by any thread of any CTA) before the call are visible to every
// thread after it, provided the readers do not go through a non-coherent path (use __ldcg / plain pointers, not
// const __restrict__, for data that other CTAs rewrite during the kernel).
// Deterministic like gridReduceTo: one double per CTA, summed by the last arriver with a fixed tree.
template <int NT>
__device__ __forceinline__ double gridAllReduce(double v, double* partials, LoopSync* S, unsigned long long k) {
    __shared__ double s_warp[NT / 32];
    __shared__ double s_val;
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warpSum(v);
    if (lane == 0) s_warp[wid] = v;
    __syncthreads();                        // also: every thread's earlier global stores precede thread 0's fence below
    if (wid == 0) {
        double t = lane < NT / 32 ? s_warp[lane] : 0.0;
        t = warpSum(t);
        if (lane == 0) {
            partials[blockIdx.x] = t;
            __threadfence();
            const unsigned int prev = atomicAdd(&S->counter, 1u);
            s_last = (prev == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double t = orderedStridedSum<NT>(partials, gridDim.x);
        t = warpSum(t);
        if (lane == 0) s_warp[wid] = t;
        __syncthreads();
        if (wid == 0) {
            double u = lane < NT / 32 ? s_warp[lane] : 0.0;
            u = warpSum(u);
            if (lane == 0) {
                S->counter = 0u;
                *reinterpret_cast<volatile double*>(&S->value[k & 1ull]) = u;
                __threadfence();
                stVolatileU64(&S->flag[k & 1ull], k);
            }
        }
    }
    if (threadIdx.x == 0) {
        SpinGuard guard;
        while (ldVolatileU64(&S->flag[k & 1ull]) != k) guard.tick();   // a lost CTA is a bug: fail, do not hang
        __threadfence();
        s_val = *reinterpret_cast<volatile double*>(&S->value[k & 1ull]);
    }
    __syncthreads();
    const double out = s_val;
    __syncthreads();   // s_val / s_warp reusable by the next call
    return out;
}

// Low-latency variant for small cooperative grids (<= kMaxLLCtas CTAs, e.g. one CTA per SM).  The serial chain of
// gridAllReduce is  partial -> fence -> atomic | last CTA: fence -> loads -> value -> fence -> flag | waiter: flag ->
// fence -> value.  Here partials and the result travel as self-validating 8-byte words (payload32 | sequence32, as in
// the NVLink mailboxes above), so none of those fences and no second dependent load is needed:
//   arrive:   two word stores + one atomicAdd on a counter that is never reset (arrival number gridDim * k marks the
//             last CTA of reduction k; k = 1, 2, 3, ... over the life of the plan, same grid size every launch)
//   last CTA: polls the gridDim partial slots (normally already valid), adds them in slot order, stores two result words
//   everyone: one thread polls the two result words.
// (Letting every CTA poll every slot instead was measured slower: 16 k pollers on 2 KB of L2.)
// `words` = [2 parities][kMaxLLCtas + 1][2]; slot kMaxLLCtas carries the result.  Two parities suffice: reduction k+2
// cannot start before every CTA has read the result of k.  releaseData: the CTA's earlier global stores must be visible
// to the other CTAs after the call (one __threadfence on each side); pass false for a pure reduction.
constexpr int kMaxLLCtas = 256;
template <int NT>
__device__ __forceinline__ double gridAllReduceLL(double v, unsigned long long* words, unsigned long long* arrivals, unsigned long long k, bool releaseData) {
    static_assert(NT >= kMaxLLCtas && NT % 32 == 0 && NT / 32 <= 32, "one polling thread per CTA slot");
    __shared__ double s_warp[NT / 32];
    __shared__ double s_part[kMaxLLCtas];
    __shared__ double s_val;
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned long long seq = k & 0xffffffffull;
    unsigned long long* row = words + (size_t)(k & 1ull) * (2 * (kMaxLLCtas + 1));
    v = warpSum(v);
    if (lane == 0) s_warp[wid] = v;
    __syncthreads();                        // also: every thread's earlier global stores precede the fence below
    if (wid == 0) {
        double t = lane < NT / 32 ? s_warp[lane] : 0.0;
        t = warpSum(t);                     // every lane holds the CTA's partial
        const unsigned long long bits = (unsigned long long)__double_as_longlong(t);
        if (lane < 2) stVolatileU64(row + 2 * blockIdx.x + lane, ((lane ? (bits >> 32) : (bits & 0xffffffffull)) << 32) | seq);
        if (lane == 0) {
            if (releaseData) __threadfence();
            const unsigned long long prev = atomicAdd(arrivals, 1ull);
            s_last = (prev + 1ull == (unsigned long long)gridDim.x * k);
        }
    }
    __syncthreads();
    if (s_last) {
        if (threadIdx.x < gridDim.x) {
            const unsigned long long* src = row + 2 * threadIdx.x;
            unsigned long long w0, w1; SpinGuard guard;
            do {
                w0 = ldVolatileU64(src); w1 = ldVolatileU64(src + 1);
                guard.tick();
            } while ((w0 & 0xffffffffull) != seq || (w1 & 0xffffffffull) != seq);
            s_part[threadIdx.x] = __longlong_as_double((long long)((w0 >> 32) | ((w1 >> 32) << 32)));
        }
        __syncthreads();
        if (wid == 0) {
            double u = 0.0;
            for (unsigned i = lane; i < gridDim.x; i += 32) u += s_part[i];   // fixed order
            u = warpSum(u);
            const unsigned long long bits = (unsigned long long)__double_as_longlong(u);
            if (releaseData) __threadfence();   // cumulativity: the arrivals observed above (and the stores they release) precede the result
            if (lane < 2) stVolatileU64(row + 2 * kMaxLLCtas + lane, ((lane ? (bits >> 32) : (bits & 0xffffffffull)) << 32) | seq);
        }
    }
    if (threadIdx.x == 0) {
        const unsigned long long* src = row + 2 * kMaxLLCtas;
        unsigned long long w0, w1; SpinGuard guard;
        do {
            w0 = ldVolatileU64(src); w1 = ldVolatileU64(src + 1);
            guard.tick();   // a lost CTA is a bug: fail, do not hang
        } while ((w0 & 0xffffffffull) != seq || (w1 & 0xffffffffull) != seq);
        if (releaseData) __threadfence();   // acquire side: loads issued after the barrier see the other CTAs' stores
        s_val = __longlong_as_double((long long)((w0 >> 32) | ((w1 >> 32) << 32)));
    }
    __syncthreads();
    const double out = s_val;
    __syncthreads();   // shared scratch reusable by the next call
    return out;
}

// Multi-GPU form of gridAllReduceLL (one cooperative kernel per rank, all ranks launch it collectively): the local part is
// identical; the last CTA of each rank then publishes the RANK's partial to every rank's NVLink mailbox (slot `slot`,
// sequence k, sender = this rank -- the wire format of publishScalar) instead of a local result slot, and warp 0 of every CTA
// polls the `world` senders of its own mailbox (one lane per word, as loadScalars2) and adds them in rank order, so every
// CTA of every rank obtains the bitwise-identical sum.  releaseData makes this CTA's earlier LOCAL global stores visible to
// the other CTAs of this rank after the call (device-scope fences, as in gridAllReduceLL).  Data for OTHER ranks must not rely
// on this call: it travels as self-validating LL words (llStore / llLoad below) -- a system-scope fence per CTA on each side
// of the reduction cost ~10 us per PCG step when it was done that way (measured on 2 B200s: 26.9 us per step).
template <int NT>
__device__ __forceinline__ double gridAllReduceLLMulti(double v, unsigned long long* words, unsigned long long* arrivals, unsigned long long k, bool releaseData,
                                                       const CommView& comm, int slot) {
    static_assert(NT >= kMaxLLCtas && NT % 32 == 0 && NT / 32 <= 32, "one polling thread per CTA slot");
    static_assert(kMaxRanks == 8, "lane mapping below assumes at most 16 mailbox words per slot");
    __shared__ double s_warp[NT / 32];
    __shared__ double s_part[kMaxLLCtas];
    __shared__ double s_val;
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned long long seq = k & 0xffffffffull;
    unsigned long long* row = words + (size_t)(k & 1ull) * (2 * (kMaxLLCtas + 1));
    v = warpSum(v);
    if (lane == 0) s_warp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = lane < NT / 32 ? s_warp[lane] : 0.0;
        t = warpSum(t);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(t);
        if (lane < 2) stVolatileU64(row + 2 * blockIdx.x + lane, ((lane ? (bits >> 32) : (bits & 0xffffffffull)) << 32) | seq);
        if (lane == 0) {
            if (releaseData) __threadfence();
            const unsigned long long prev = atomicAdd(arrivals, 1ull);
            s_last = (prev + 1ull == (unsigned long long)gridDim.x * k);
        }
    }
    __syncthreads();
    if (s_last) {
        if (threadIdx.x < gridDim.x) {
            const unsigned long long* src = row + 2 * threadIdx.x;
            unsigned long long w0, w1; SpinGuard guard;
            do {
                w0 = ldVolatileU64(src); w1 = ldVolatileU64(src + 1);
                guard.tick();
            } while ((w0 & 0xffffffffull) != seq || (w1 & 0xffffffffull) != seq);
            s_part[threadIdx.x] = __longlong_as_double((long long)((w0 >> 32) | ((w1 >> 32) << 32)));
        }
        __syncthreads();
        if (wid == 0) {
            double u = 0.0;
            for (unsigned i = lane; i < gridDim.x; i += 32) u += s_part[i];   // fixed order
            u = warpSum(u);
            const unsigned long long bits = (unsigned long long)__double_as_longlong(u);
            if (releaseData) __threadfence();
            if (lane < 2 * comm.world)   // lane = 2 * destination rank + word
                stVolatileU64(comm.peerWords[lane >> 1] + mailboxWord(slot, k, comm.rank) + (lane & 1), (((lane & 1) ? (bits >> 32) : (bits & 0xffffffffull)) << 32) | seq);
        }
    }
    if (wid == 0) {
        unsigned long long w = 0ull;
        if (lane < 2 * comm.world) {     // lane = 2 * sender + word
            const unsigned long long* src = comm.peerWords[comm.rank] + mailboxWord(slot, k, lane >> 1) + (lane & 1);
            SpinGuard guard;
            do { w = ldVolatileU64(src); guard.tick(); } while ((w & 0xffffffffull) != seq);   // a lost rank is a bug: fail, do not hang
        }
        __syncwarp();
        const unsigned payload = (unsigned)(w >> 32);
        double sum = 0.0;
        for (int r = 0; r < comm.world; ++r) {   // fixed rank order => bitwise identical on every rank
            const unsigned lo = __shfl_sync(0xffffffffu, payload, 2 * r), hi = __shfl_sync(0xffffffffu, payload, 2 * r + 1);
            sum += __longlong_as_double((long long)((unsigned long long)lo | ((unsigned long long)hi << 32)));
        }
        if (lane == 0) { if (releaseData) __threadfence(); s_val = sum; }
    }
    __syncthreads();
    const double out = s_val;
    __syncthreads();
    return out;
}

// Two values in ONE round, no data release: the all-reduce of the single-reduction PCG loop kernel (gamma = r.z and delta = w.z travel
// together; tile edges move as LL data words, so the reduction carries no ordering duty and contains no fence at all).
// `words` = [2 parities][kMaxLLCtas + 1][4]: per CTA slot the lo/hi words of both values, slot kMaxLLCtas = the result (single GPU);
// MULTI: the rank partials go to mailbox slots `slot` and `slot + 1` of every rank (4 * world words = one warp at world = 8).
// RELEASE: the CTA's earlier global stores are visible to the other CTAs of THIS GPU after the call (the three device-scope fences of
// gridAllReduceLL's releaseData; single GPU use -- k_sfsLoopResident publishes its tile-edge z that way).
template <int NT, bool MULTI, bool RELEASE = false>
__device__ __forceinline__ void gridAllReduce2LL(double v0, double v1, unsigned long long* words, unsigned long long* arrivals, unsigned long long k,
                                                 const CommView& comm, int slot, double& o0, double& o1) {
    static_assert(NT >= kMaxLLCtas && NT % 32 == 0 && NT / 32 <= 32, "one polling thread per CTA slot");
    static_assert(kMaxRanks == 8, "lane mapping below assumes at most 32 mailbox words for the two values");
    __shared__ double s_w[2][NT / 32];
    __shared__ double s_p[2][kMaxLLCtas];
    __shared__ double s_v[2];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned long long seq = k & 0xffffffffull;
    unsigned long long* row = words + (size_t)(k & 1ull) * (4 * (kMaxLLCtas + 1));
    auto word = [&](double a, double b, int l) -> unsigned long long {   // word l of the pair: 0,1 = lo/hi of a; 2,3 = lo/hi of b
        const unsigned long long bits = (unsigned long long)__double_as_longlong((l & 2) ? b : a);
        return (((l & 1) ? (bits >> 32) : (bits & 0xffffffffull)) << 32) | seq;
    };
    auto unword = [](unsigned long long lo, unsigned long long hi) -> double { return __longlong_as_double((long long)((lo >> 32) | ((hi >> 32) << 32))); };
    v0 = warpSum(v0); v1 = warpSum(v1);
    if (lane == 0) { s_w[0][wid] = v0; s_w[1][wid] = v1; }
    __syncthreads();
    if (wid == 0) {
        double t0 = lane < NT / 32 ? s_w[0][lane] : 0.0, t1 = lane < NT / 32 ? s_w[1][lane] : 0.0;
        t0 = warpSum(t0); t1 = warpSum(t1);
        if (lane < 4) stVolatileU64(row + 4 * blockIdx.x + lane, word(t0, t1, lane));
        if (lane == 0) { if (RELEASE) __threadfence(); const unsigned long long prev = atomicAdd(arrivals, 1ull); s_last = (prev + 1ull == (unsigned long long)gridDim.x * k); }
    }
    __syncthreads();
    if (s_last) {
        if (threadIdx.x < gridDim.x) {
            const unsigned long long* src = row + 4 * threadIdx.x;
            unsigned long long w[4]; SpinGuard guard; bool ok;
            do {
                ok = true;
#pragma unroll
                for (int i = 0; i < 4; ++i) { w[i] = ldVolatileU64(src + i); ok = ok && ((w[i] & 0xffffffffull) == seq); }
                guard.tick();
            } while (!ok);
            s_p[0][threadIdx.x] = unword(w[0], w[1]); s_p[1][threadIdx.x] = unword(w[2], w[3]);
        }
        __syncthreads();
        if (wid == 0) {
            double u0 = 0.0, u1 = 0.0;
            for (unsigned i = lane; i < gridDim.x; i += 32) { u0 += s_p[0][i]; u1 += s_p[1][i]; }   // fixed order
            u0 = warpSum(u0); u1 = warpSum(u1);
            if (RELEASE) __threadfence();   // cumulativity: the arrivals observed above (and the stores they release) precede the result
            if (MULTI) { if (lane < 4 * comm.world) stVolatileU64(comm.peerWords[lane >> 2] + mailboxWord(slot + ((lane >> 1) & 1), k, comm.rank) + (lane & 1), word(u0, u1, lane & 3)); }
            else if (lane < 4) stVolatileU64(row + 4 * kMaxLLCtas + lane, word(u0, u1, lane));
        }
    }
    if (wid == 0) {
        unsigned long long w = 0ull;
        const int nWords = MULTI ? 4 * comm.world : 4;
        if (lane < nWords) {
            const unsigned long long* src = MULTI ? comm.peerWords[comm.rank] + mailboxWord(slot + ((lane >> 1) & 1), k, lane >> 2) + (lane & 1) : row + 4 * kMaxLLCtas + lane;
            SpinGuard guard;
            do { w = ldVolatileU64(src); guard.tick(); } while ((w & 0xffffffffull) != seq);   // a lost CTA / rank is a bug: fail, do not hang
        }
        __syncwarp();
        const unsigned payload = (unsigned)(w >> 32);
        double sum0 = 0.0, sum1 = 0.0;
        const int senders = MULTI ? comm.world : 1;
        for (int r = 0; r < senders; ++r) {   // fixed rank order => bitwise identical on every rank
            const unsigned a = __shfl_sync(0xffffffffu, payload, 4 * r), b = __shfl_sync(0xffffffffu, payload, 4 * r + 1);
            const unsigned c = __shfl_sync(0xffffffffu, payload, 4 * r + 2), d = __shfl_sync(0xffffffffu, payload, 4 * r + 3);
            sum0 += __longlong_as_double((long long)((unsigned long long)a | ((unsigned long long)b << 32)));
            sum1 += __longlong_as_double((long long)((unsigned long long)c | ((unsigned long long)d << 32)));
        }
        if (lane == 0) { if (RELEASE) __threadfence(); s_v[0] = sum0; s_v[1] = sum1; }   // acquire side: loads issued after the call see the other CTAs' stores
    }
    __syncthreads();
    o0 = s_v[0]; o1 = s_v[1];
    __syncthreads();
}

// Self-validating data words for peer memory (NCCL-LL style): an opt_float travels as one (float) or two (double) 8-byte words, each
// payload32 | sequence32, written with a single 8-byte store.  The reader polls the word itself until it carries the expected
// sequence -- no flag, no fence on either side; the sequence only grows, so a stale word is never mistaken for a new one.
__device__ __forceinline__ void llStore(unsigned long long* slot, float v, unsigned seq) {
    stVolatileU64(slot, ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)seq);
}
__device__ __forceinline__ void llStore(unsigned long long* slot, double v, unsigned seq) {   // slot[0], slot[1]
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    stVolatileU64(slot, ((bits & 0xffffffffull) << 32) | (unsigned long long)seq);
    stVolatileU64(slot + 1, ((bits >> 32) << 32) | (unsigned long long)seq);
}
__device__ __forceinline__ void llLoad(const unsigned long long* slot, unsigned seq, float& out) {
    unsigned long long w; SpinGuard guard;
    do { w = ldVolatileU64(slot); guard.tick(); } while ((unsigned)(w & 0xffffffffull) != seq);   // a lost rank is a bug: fail, do not hang
    out = __uint_as_float((unsigned)(w >> 32));
}
__device__ __forceinline__ void llLoad(const unsigned long long* slot, unsigned seq, double& out) {
    unsigned long long w0, w1; SpinGuard guard;
    do { w0 = ldVolatileU64(slot); w1 = ldVolatileU64(slot + 1); guard.tick(); } while ((unsigned)(w0 & 0xffffffffull) != seq || (unsigned)(w1 & 0xffffffffull) != seq);
    out = __longlong_as_double((long long)((w0 >> 32) | ((w1 >> 32) << 32)));
}
template <class T> struct LLWords;
template <> struct LLWords<float> { static constexpr int N = 1; };
template <> struct LLWords<double> { static constexpr int N = 2; };
constexpr size_t kPeerScratchBytes = 1u << 20;   // per-plan scratch behind the mailbox block, reachable by the peers (LL halo rows of the loop kernel)

// Programmatic dependent launch (PDL): a kernel launched with launchPdl may begin while its predecessor on the stream
// is still draining; it must call pdlWait() before touching anything the predecessor wrote.  Kernels call
// pdlTrigger() first so that THEIR successor can be scheduled early.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdlTrigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdlWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <class... KArgs, class... Args>
inline void launchPdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    OPT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

inline int divUp(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace optb
