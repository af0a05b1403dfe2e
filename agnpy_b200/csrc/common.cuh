// Shared device functions for the agnpy_b200 emission-integral kernels (sm_100a, FP64).
//
// The whole library is compiled with -fmad=false: every a*b+c written with plain operators
// is rounded twice exactly like numpy does, so mask operands (gamma >= gamma_min,
// gamma_min <= gamma/delta_D <= gamma_max, s < 1, ...) are bit-identical to the reference's.
// FMAs are used only where written explicitly (fma()) inside the hot accumulation loops.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/agnpy_b200.h"

// Bounds / invariant checks of the kernels' index arithmetic.  Compiled in only with -DAGNPY_B200_CHECKS
// (`python agnpy_b200/build.py --checked` -> libagnpy_b200_checked.so, selected at run time by
// AGNPY_B200_CHECKED=1); a failed check prints its location and traps, which the host sees as a CUDA
// error.  compute-sanitizer is not available on the measurement pool: the GPU test-suite run against the
// checked build (profiles/r2_checked_build.txt) stands in for memcheck on the shared-memory indexing.
#ifdef AGNPY_B200_CHECKS
#include <stdio.h>
#define AGB_CHECK(cond)                                                                          \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            printf("AGB_CHECK failed: %s  (%s:%d, block %d,%d,%d thread %d)\n", #cond, __FILE__, \
                   __LINE__, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);                   \
            __trap();                                                                            \
        }                                                                                        \
    } while (0)
#else
#define AGB_CHECK(cond) ((void)0)
#endif

namespace agb {

// ---- constants: astropy>=7 (CODATA 2018) in CGS; literals are the repr() of the doubles the
// ---- reference computes (agnpy/utils/conversion.py:1-9, agnpy/synchrotron/synchrotron.py:12)
constexpr double kC = 29979245800.0;
constexpr double kH = 6.62607015e-27;
constexpr double kMe = 9.1093837015e-28;
constexpr double kMp = 1.67262192369e-24;
constexpr double kMpc2 = 0.0015032776159851257;  // m_p * c**2
constexpr double kE = 4.803204712570263e-10;
constexpr double kSigmaT = 6.6524587321e-25;
constexpr double kG = 6.6743e-08;
constexpr double kMsun = 1.988409870698051e+33;
constexpr double kMec2 = 8.187105776823886e-07;   // m_e * c**2
constexpr double kLambdaC = 2.426310238683092e-10; // h / (m_e c)
constexpr double kPi = 3.141592653589793;
constexpr double kTiny = 2.2250738585072014e-308;
constexpr double kFmax = 1.7976931348623157e+308;

// ---- small helpers -------------------------------------------------------------------------
__device__ __forceinline__ double sq(double x) { return x * x; }

// utils/math.py:49-52  log(clip(x, tiny, fmax)); NaN propagates like np.clip
__device__ __forceinline__ double clipped_log(double x) {
    double v = x;
    if (v < kTiny) v = kTiny;
    if (v > kFmax) v = kFmax;
    return log(v);
}

// nu -> dimensionless energy in a frame (z, delta_D): conversion.py:30-35
__device__ __forceinline__ double nu_to_eps(double nu, double z, double delta_D) {
    double e0 = kH * nu / kMec2;
    return (1.0 + z) * e0 / delta_D;
}

// spectra.py:210-225: value of the (possibly snapped) k-th grid point
__device__ __forceinline__ double snapped(double g, int k, int n, double g0, double gN,
                                          double gmin, double gmax) {
    // g0 / gN are the unsnapped first / last grid values
    if (k == 0 && gmin > g0 && (gmin - g0) / g0 < 1e-10) return gmin;
    if (k == n - 1 && gN > gmax && (gN - gmax) / gmax < 1e-10) return gmax;
    return g;
}

// ---- electron distributions: spectra.py static evaluate() of the five analytic shapes --------
// par layout = the reference's `parameters` order (see include/agnpy_b200.h)
__device__ __forceinline__ void ne_bounds(int kind, const double* p, double& gmin, double& gmax) {
    switch (kind) {
        case AGNPY_NE_POWERLAW: gmin = p[2]; gmax = p[3]; break;
        case AGNPY_NE_BROKEN_POWERLAW: gmin = p[4]; gmax = p[5]; break;
        case AGNPY_NE_LOG_PARABOLA: gmin = p[4]; gmax = p[5]; break;
        case AGNPY_NE_EXP_CUTOFF_POWERLAW: gmin = p[3]; gmax = p[4]; break;
        default: gmin = p[5]; gmax = p[6]; break;  // EXP_CUTOFF_BROKEN_POWERLAW
    }
}

__device__ inline double ne_eval(int kind, const double* p, double g) {
    double gmin, gmax;
    ne_bounds(kind, p, gmin, gmax);
    if (!((gmin <= g) && (g <= gmax))) return 0.0;
    switch (kind) {
        case AGNPY_NE_POWERLAW:  // spectra.py:272-277
            return p[0] * pow(g, -p[1]);
        case AGNPY_NE_BROKEN_POWERLAW: {  // :366-374
            double idx = (g <= p[3]) ? p[1] : p[2];
            return p[0] * pow(g / p[3], -idx);
        }
        case AGNPY_NE_LOG_PARABOLA: {  // :470-479
            double ratio = g / p[3];
            double idx = -p[1] - p[2] * log10(ratio);
            return p[0] * pow(ratio, idx);
        }
        case AGNPY_NE_EXP_CUTOFF_POWERLAW:  // :569-576
            return p[0] * pow(g, -p[1]) * exp(-g / p[2]);
        default: {  // EXP_CUTOFF_BROKEN_POWERLAW :673-681  (k,p1,p2,gamma_c,gamma_b,..)
            double idx = (g <= p[4]) ? p[1] : p[2];
            return p[0] * pow(g / p[4], -idx) * exp(-g / p[3]);
        }
    }
}

// gamma^2 d/dgamma(n_e/gamma^2): spectra.py:282-290, 387-396, 492-499, 583-591, 695-710
__device__ inline double ne_ssa(int kind, const double* p, double g) {
    double gmin, gmax;
    ne_bounds(kind, p, gmin, gmax);
    if (!((gmin <= g) && (g <= gmax))) return 0.0;
    switch (kind) {
        case AGNPY_NE_POWERLAW:
            return p[0] * (-(p[1] + 2.0) * pow(g, -p[1] - 1.0));
        case AGNPY_NE_BROKEN_POWERLAW: {
            double idx = (g <= p[3]) ? p[1] : p[2];
            return p[0] * -(idx + 2.0) / g * pow(g / p[3], -idx);
        }
        case AGNPY_NE_LOG_PARABOLA: {
            double pref = -(p[1] + 2.0 * p[2] * log10(g / p[3]) + 2.0) / g;
            return pref * ne_eval(kind, p, g);
        }
        case AGNPY_NE_EXP_CUTOFF_POWERLAW: {
            double pref = -(p[1] + 2.0) / g + (-1.0 / p[2]);
            return pref * ne_eval(kind, p, g);
        }
        default: {
            double idx = (g <= p[4]) ? p[1] : p[2];
            double pref = -(idx + 2.0) / g + (-1.0 / p[3]);
            return pref * ne_eval(kind, p, g);
        }
    }
}

// ---- integration rules -------------------------------------------------------------------
// numpy.trapz as a fixed weighted sum: w_0=(x1-x0)/2, w_k=(x_{k+1}-x_{k-1})/2, w_{n-1}=...
__device__ __forceinline__ double trapz_weight(const double* x, int k, int n) {
    if (n < 2) return 0.0;
    double lo = x[k > 0 ? k - 1 : 0];
    double hi = x[k < n - 1 ? k + 1 : n - 1];
    return (hi - lo) * 0.5;
}

// one interval of utils/math.py:55-119 (trapz_loglog); ly_* are clipped logs of y_*.
// m is the slope of the power law through both end points, so (x_up/x_lo)^m is y_up/y_lo by
// construction and  y_lo/(m+1) (x_up (x_up/x_lo)^m - x_lo) = (x_up y_up - x_lo y_lo)/(m+1):
// no pow / exp per interval.  (The reference evaluates the power numerically; the two agree to
// ~|m ln(x_up/x_lo)| ulp.)  The identity needs the logs of the values themselves, so a y
// that the reference clips (below the smallest normal double) keeps the literal formula.
__device__ __forceinline__ double loglog_interval(double x_lo, double x_up, double y_lo,
                                                  double y_up, double ly_lo, double ly_up) {
    if (y_lo == 0.0 || y_up == 0.0 || x_lo == x_up) return 0.0;  // :112-117
    double m = (ly_lo - ly_up) / (clipped_log(x_lo) - clipped_log(x_up));  // :104
    double m1 = m + 1.0;
    if (fabs(m1) > 1e-10) {  // :106-110
        if (fabs(y_lo) >= kTiny && fabs(y_up) >= kTiny && y_lo > 0.0 && y_up > 0.0)
            return (x_up * y_up - x_lo * y_lo) / m1;
        return y_lo / m1 * (x_up * pow(x_up / x_lo, m) - x_lo);
    }
    return x_lo * y_lo * clipped_log(x_up / x_lo);
}

// same interval with the per-interval constants precomputed: inv_dlx = 1/(ln x_lo - ln x_up),
// lr = ln(x_up/x_lo)
__device__ __forceinline__ double loglog_interval_pre(double x_lo, double x_up, double inv_dlx,
                                                      double lr, double y_lo, double y_up,
                                                      double ly_lo, double ly_up) {
    if (y_lo == 0.0 || y_up == 0.0 || x_lo == x_up) return 0.0;
    double m = (ly_lo - ly_up) * inv_dlx;
    double m1 = m + 1.0;
    if (fabs(m1) > 1e-10) {
        if (y_lo >= kTiny && y_up >= kTiny) return (x_up * y_up - x_lo * y_lo) / m1;
        return y_lo / m1 * (x_up * exp(m * lr) - x_lo);
    }
    return x_lo * y_lo * lr;
}

// The same interval without logarithms.  With z = x y the exact integral of the power law through
// both end points is  ln(x_up/x_lo) * L(z_lo, z_up),  L(a, b) = (b - a)/(ln b - ln a)  the logarithmic
// mean -- this is (x_up y_up - x_lo y_lo)/(m + 1) with m + 1 = ln(z_up/z_lo)/ln(x_up/x_lo) -- and
//     L(a, b) = (a + b)/2 * u/atanh(u),   u = (b - a)/(b + a),
// where u/atanh(u) = 1 - w/3 - 4w^2/45 - 44w^3/945 - ... (w = u^2) converges fast when neighbouring
// samples are within a factor ~2 of each other (|u| <= 0.36: the degree-16 series is exact to
// 4e-18; on the reference's grids of 25-40 points per decade that is every interval except those
// across a cut-off).  One reciprocal and 16 FMAs instead of a logarithm per point and a division
// per interval; outside |u| <= 0.36, and for samples the reference would clip (below the smallest
// normal double), the caller falls back to loglog_interval_pre.  The m = -1 switch of utils/math.py
// :106-110 needs no special case here: L is smooth through z_up = z_lo.
__device__ __forceinline__ double rcp_fast(double x) {  // 1/x to ~1 ulp for normal positive x
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));  // ~20 good bits; two Newton steps: > 53
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double rsqrt_fast(double v) {  // v^(-1/2) to ~1 ulp for normal positive v
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));  // ~20 good bits; two Newton steps: > 53
    double hv = 0.5 * v;
    double e = fma(-hv * y, y, 0.5);   // (1 - v y^2) / 2
    y = fma(y, e, y);
    e = fma(-hv * y, y, 0.5);
    return fma(y, e, y);
}
constexpr double kLogMeanW = 0.1296;  // |u| <= 0.36: neighbouring samples within a factor 2.1
__device__ __forceinline__ double logmean_series(double w) {  // u/atanh(u), w = u^2 <= kLogMeanW
    // coefficients of 1 / sum_n w^n/(2n+1), rounded from their exact rational values
    double t = -0.004352550122988381;
    t = fma(t, w, -0.004746430692874202);
    t = fma(t, w, -0.005208475658374312);
    t = fma(t, w, -0.005756954446288888);
    t = fma(t, w, -0.006417063031397296);
    t = fma(t, w, -0.007224464737393906);
    t = fma(t, w, -0.008231206567350501);
    t = fma(t, w, -0.009516073194527899);
    t = fma(t, w, -0.011203745637718733);
    t = fma(t, w, -0.013502765051265932);
    t = fma(t, w, -0.016787551856334924);
    t = fma(t, w, -0.021796804019026242);
    t = fma(t, w, -0.030194003527336862);
    t = fma(t, w, -0.04656084656084656);
    t = fma(t, w, -4.0 / 45.0);
    t = fma(t, w, -1.0 / 3.0);
    return fma(t, w, 1.0);
}
// the same series cut after w^8: exact to 6e-17 for w <= 0.02 (neighbouring samples within a factor
// 1.33 -- every interval of a smooth spectrum on the reference's 25-40 points-per-decade grids)
constexpr double kLogMeanWSmall = 0.02;
__device__ __forceinline__ double logmean_series8(double w) {
    double t = -0.011203745637718733;
    t = fma(t, w, -0.013502765051265932);
    t = fma(t, w, -0.016787551856334924);
    t = fma(t, w, -0.021796804019026242);
    t = fma(t, w, -0.030194003527336862);
    t = fma(t, w, -0.04656084656084656);
    t = fma(t, w, -4.0 / 45.0);
    t = fma(t, w, -1.0 / 3.0);
    return fma(t, w, 1.0);
}
// interval [x_lo, x_up] from z = x y at both ends (both > 0) and lr = ln(x_up/x_lo); `ok` is false
// when the series does not apply (the caller then uses the logarithmic form).  The short series is
// taken when every active lane of the warp qualifies (one vote, no divergence).
__device__ __forceinline__ double loglog_interval_z(double z_lo, double z_up, double lr, bool& ok) {
    const double s = z_up + z_lo, d = z_up - z_lo;
    const double u = d * rcp_fast(s);
    const double w = u * u;
    ok = w <= kLogMeanW;
    const bool all_small = __all_sync(__activemask(), w <= kLogMeanWSmall);
    const double f = all_small ? logmean_series8(w) : logmean_series(w);
    return lr * ((0.5 * s) * f);
}

// ---- reductions ------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum (fixed tree); result valid in thread 0. `red` needs >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    return v;
}

// synchrotron.py:16-24 R(x) (Aharonian et al. 2010 D7); x^(1/3) via cbrt
__device__ __forceinline__ double R_synch(double x) {
    double x13 = cbrt(x);
    double x23 = x13 * x13;
    double x43 = x23 * x23;
    double t1 = 1.808 * x13 / sqrt(1.0 + 3.4 * x23);
    double t2n = 1.0 + 2.21 * x23 + 0.347 * x43;
    double t2d = 1.0 + 1.353 * x23 + 0.217 * x43;
    return t1 * t2n / t2d * exp(-x);
}

// synchrotron.py:64-68 attenuation 3u/tau, u = 1/2 + e^-t/t - (1-e^-t)/t^2.  The reference
// formula cancels catastrophically for small tau (abs. error ~ ulp/tau^2); it is evaluated
// here through its series u = sum_{n>=1} (-1)^(n+1) t^n (n+1)/(n+2)! for tau < 0.5 -- unless the
// strict-reference switch (agnpy_b200_strict_reference) asks for the literal formula.
__device__ __forceinline__ double ssa_attenuation(double tau, bool strict = false) {
    if (tau < 1e-3) return 1.0;
    if (tau < 0.5 && !strict) {
        double term = tau / 2.0;  // t^n/(n+1)!  for n=1
        double u = 0.0, sgn = 1.0;
        for (int n = 1; n <= 24; ++n) {
            // coefficient (n+1)/(n+2)! = [1/(n+1)!] * (n+1)/(n+2)
            u += sgn * term * (double)(n + 1) / (double)(n + 2);
            term = term * tau / (double)(n + 2);
            sgn = -sgn;
        }
        return 3.0 * u / tau;
    }
    double e = exp(-tau);
    double u = 0.5 + e / tau - (1.0 - e) / (tau * tau);  // np.power(tau, 2) == tau * tau
    return 3.0 * u / tau;
}
// bit 1 of the `ssa` / `blob_frame` kernel arguments carries the strict-reference switch
constexpr int kStrictBit = 2;

// launch bookkeeping shared by all translation units (defined in runtime.cu)
void note_launch(int n = 1);
void* workspace(size_t bytes, cudaStream_t stream);
int fail(int code, const char* what);
int check_last(const char* what);
bool strict_reference();  // agnpy_b200_strict_reference(): literal reference formulas where we deviate
// CUDA-event bracket around the dominant kernel of a call (no-ops unless enabled)
void timing_begin(cudaStream_t stream);
void timing_end(cudaStream_t stream);

}  // namespace agb
