// lm_decide.cuh -- the Levenberg-Marquardt accept / reject / trust-region decision of one nonlinear iteration
// (reference API/src/solverGPUGaussNewton.t:1119-1157) as ONE function of scalars, usable on the device: with it the decision needs no
// host round trip, and Opt_ProblemSolve can enqueue all remaining LM iterations back to back (solver.cu, OPTB_DEVICE_LM=1).
// Same operations in the same precision as Plan<T>::decide (solver.cu): T = opt_float for the costs and the radius, double for the
// cubic of the step quality, as the reference's Terra code computes them.  Host/device code: tests/host/lm_decide_host.cpp runs it on
// the CPU against an independent transcription of the reference lines.
#pragma once
#include <cmath>

#ifndef OPTB_HD
#ifdef __CUDACC__
#define OPTB_HD __host__ __device__ __forceinline__
#else
#define OPTB_HD inline
#endif
#endif

namespace optb {

enum LmVerdict : int {
    kLmAccepted = 0,            // step kept; radius grown / shrunk by the step quality; prevCost = newCost           (:1131-1146)
    kLmRejected = 1,            // step reverted; radius /= decrease factor, factor doubled; precompute to be redone     (:1147-1156)
    kLmFunctionTolerance = 2,   // accepted step changed the cost by <= prevCost * function_tolerance: solve ends        (:1126-1130)
    kLmRadiusTooSmall = 3       // rejected step took the radius to <= min_trust_region_radius: revert, then solve ends  (:1152-1155)
};

template <class T> struct LmState { T prevCost, radius, decreaseFactor; };
template <class T> struct LmLimits { T min_relative_decrease, function_tolerance, max_trust_region_radius, min_trust_region_radius; };

template <class T>
OPTB_HD int lmDecide(T newCost, T modelCost, LmState<T>& s, const LmLimits<T>& lim) {
    const T model_cost_change = s.prevCost - modelCost;
    const T cost_change = s.prevCost - newCost;
    const T relative_decrease = cost_change / model_cost_change;
    if (cost_change >= T(0) && relative_decrease > lim.min_relative_decrease) {
        const T absolute_function_tolerance = s.prevCost * lim.function_tolerance;
        if (cost_change <= absolute_function_tolerance) return kLmFunctionTolerance;
        const double min_factor = 1.0 / 3.0;
        const double tmp_factor = 1.0 - pow(2.0 * (double)relative_decrease - 1.0, 3.0);
        s.radius = (T)((double)s.radius / fmax(min_factor, tmp_factor));
        s.radius = (T)fmin((double)s.radius, (double)lim.max_trust_region_radius);
        s.decreaseFactor = T(2.0);
        s.prevCost = newCost;
        return kLmAccepted;
    }
    s.radius = s.radius / s.decreaseFactor;
    s.decreaseFactor = T(2.0) * s.decreaseFactor;
    if (s.radius <= lim.min_trust_region_radius) return kLmRadiusTooSmall;
    return kLmRejected;
}

}  // namespace optb
