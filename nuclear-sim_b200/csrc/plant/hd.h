// Host/device portability layer for the single-source plant physics.
//
// Every function under csrc/plant/ is plain scalar FP64 written once and compiled
//   * by nvcc for sm_100a  (the product: one thread advances one plant), and
//   * by g++ for the host  (test infrastructure only; the product never links that build).
// Arithmetic must match CPython/numpy scalar semantics bit for bit, so:
//   - no FMA contraction (nvcc -fmad=false, g++ -ffp-contract=off),
//   - py_max/py_min/np_clip reproduce Python's and numpy's NaN/ordering behaviour,
//   - sums are written in the reference's left-to-right order.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NPS_HD __host__ __device__ __forceinline__
#define NPS_HD_NOINLINE __host__ __device__ __noinline__
// Helpers called from many sites (saturation temperatures: 24 call sites; water chemistry: 3; oil quality: 2) are
// compiled ONCE on the device instead of inlined everywhere: the step kernel is ~0.5 MB of straight-line SASS and
// 15 % of its stall samples were instruction-fetch misses; sharing these bodies measured +3.5 % (profiles/
// r01_tuning_variants.txt).  NPS_INLINE_HELPERS restores full inlining.
#if !defined(NPS_INLINE_HELPERS)
#define NPS_HD_SHARED __host__ __device__ __noinline__
#else
#define NPS_HD_SHARED __host__ __device__ __forceinline__
#endif
#else
#define NPS_HD inline
#define NPS_HD_NOINLINE
#define NPS_HD_SHARED inline
#endif

// Loops over repeated plant units (4 pumps, 3 SGs, 14 stages, 4 bearings) stay rolled on the device unless
// NPS_UNROLL_UNITS is defined: the step kernel is ~36 K SASS instructions and instruction-cache bound otherwise.
#if defined(__CUDA_ARCH__) && !defined(NPS_UNROLL_UNITS)
#define NPS_UNIT_LOOP _Pragma("unroll 1")
#else
#define NPS_UNIT_LOOP
#endif

// Software prefetch of plant state (device only).  A thread's PlantState lives in local memory and streams
// through L1/L2 from HBM every substep (10 KB per plant x 448 plants per SM does not fit on chip), so loads issued
// at the point of use expose the full DRAM latency.  NPS_PREFETCH(obj) issues prefetch.L1 for the fields of obj that
// are live on entry to a step (live_gen.inc) a few thousand instructions ahead of their first use; local memory
// interleaves 32-bit words across the warp, so each double needs both of its words prefetched.
// NPS_PREFETCH: short distance (about one turbine stage, < 3 K instructions) into L1.  NPS_PREFETCH_FAR: one whole
// pump / steam generator / subsystem ahead; an L1 line does not survive that long under 14 warps of streaming
// traffic per SM, so the far variant targets L2 (NPS_PF_FAR_LEVEL: 0 off, 1 L1, 2 L2).
#ifndef NPS_PF_NEAR_LEVEL
#define NPS_PF_NEAR_LEVEL 1
#endif
#ifndef NPS_PF_FAR_LEVEL
#define NPS_PF_FAR_LEVEL 0   /* measured on B200 (profiles/r01_tuning_variants.txt (2)): far prefetch costs 12-25 %, near is neutral */
#endif
#if defined(__CUDA_ARCH__) && !defined(NPS_NO_PREFETCH)
#define NPS_PF2_L1(a, off) do { asm volatile("prefetch.L1 [%0];" ::"l"((a) + (off))); \
                                asm volatile("prefetch.L1 [%0];" ::"l"((a) + (off) + 4)); } while (0)
#define NPS_PF2_L2(a, off) do { asm volatile("prefetch.L2 [%0];" ::"l"((a) + (off))); \
                                asm volatile("prefetch.L2 [%0];" ::"l"((a) + (off) + 4)); } while (0)
#define NPS_PF2(a, off) do { if (LEVEL == 1) NPS_PF2_L1(a, off); else if (LEVEL == 2) NPS_PF2_L2(a, off); } while (0)
#define NPS_PREFETCH(obj) nps_prefetch_live<NPS_PF_NEAR_LEVEL>(obj)
#define NPS_PREFETCH_FAR(obj) nps_prefetch_live<NPS_PF_FAR_LEVEL>(obj)
// NPS_PREFETCH_SELF: at the start of a unit's own processing.  Level 1/2 = prefetch.L1/.L2; level 3 = "burst touch":
// real loads of every live-in field issued back to back (one exposed DRAM latency instead of one per field).
#ifndef NPS_PF_SELF_LEVEL
#define NPS_PF_SELF_LEVEL 0
#endif
#if NPS_PF_SELF_LEVEL == 3
#define NPS_PREFETCH_SELF(obj) do { double t__ = nps_touch_live(obj), s__; asm volatile("mov.f64 %0, %1;" : "=d"(s__) : "d"(t__)); } while (0)
#else
#define NPS_PREFETCH_SELF(obj) nps_prefetch_live<NPS_PF_SELF_LEVEL>(obj)
#endif
#else
#define NPS_PREFETCH_SELF(obj) ((void)0)
#define NPS_PF2(a, off) ((void)0)
#define NPS_PREFETCH(obj) ((void)0)
#define NPS_PREFETCH_FAR(obj) ((void)0)
#endif

// NPS_TOUCH(field): a volatile read whose value is dropped.  Put in a group at the top of a function it starts the
// DRAM round trips of every state field the function is about to read, all at once; the function's own reads a few
// hundred instructions later then find the lines in L1 (or merge with the fill in flight) instead of paying one round
// trip each.  Works at function granularity only: 16 KB of L1 per warp does not keep a line for thousands of
// instructions (profiles/r01_tuning_variants.txt (2), (9)).  Host builds: nothing.
#if defined(__CUDA_ARCH__) && !defined(NPS_NO_TOUCH)
#define NPS_TOUCH(x) do { double nps_t__ = *reinterpret_cast<const volatile double*>(&(x)); (void)nps_t__; } while (0)
#else
#define NPS_TOUCH(x) ((void)0)
#endif

namespace nps {

// Python builtin max(a, b): returns a unless b > a.  (max(0, nan) == 0, max(nan, 0) is nan)
NPS_HD double py_max(double a, double b) { return (b > a) ? b : a; }
// Python builtin min(a, b): returns a unless b < a.
NPS_HD double py_min(double a, double b) { return (b < a) ? b : a; }
NPS_HD double py_max3(double a, double b, double c) { return py_max(py_max(a, b), c); }
NPS_HD double py_min3(double a, double b, double c) { return py_min(py_min(a, b), c); }
// numpy.clip(x, lo, hi) on scalars == minimum(maximum(x, lo), hi); NaN propagates.
NPS_HD double np_clip(double x, double lo, double hi) {
    double t = (x < lo) ? lo : x;   // NaN: comparison false -> stays NaN
    return (t > hi) ? hi : t;
}
// Python/numpy scalar `x ** y` on floats is libm pow (even for y == 2: glibc pow(x, 2.0) differs from
// x*x in ~0.08 % of cases, so the host build uses -fno-builtin-pow to keep the libm call).
// On the device the exponents whose result is exactly computable go through correctly-rounded arithmetic
// instead of libdevice's general pow (<= 2 ulp, ~230 SASS instructions per call and 35 % of the step kernel's
// executed instructions in the round-1 profile; 280 calls per plant-step, 125 of them with y == 2): x*x and sqrt(x)
// ARE the correctly rounded pow(x, 2) and pow(x, 0.5), i.e. at least as close to glibc's (<= 0.52 ulp) as libdevice
// pow is.  The exponent is a literal at
// most call sites, so the test folds away; the wear exponents are batch-uniform parameters (uniform branch).
// Every other exponent goes to nps_pow_general: fastpow.h for a positive finite base (all 78 general calls of a plant
// step), libdevice pow for anything else.
}  // namespace nps
#include "fastpow.h"
namespace nps {
NPS_HD_SHARED double nps_pow_general(double x, double y) {
#if defined(__CUDA_ARCH__) && !defined(NPS_LIBDEVICE_POW)
    double r;
#if defined(NPS_POW_TABLES)
    // table-driven form: 0.50 ulp and ~30 fewer instructions per call, but its five table loads miss L1 under the state
    // frame's streaming traffic: measured 2 % SLOWER than the series form (profiles/r02_ab_pow.txt) - tuning builds only
    if (nps_pow_pos_tab(x, y, r)) return r;
#else
    if (nps_pow_pos(x, y, r)) return r;     // positive finite base, moderate result: fastpow.h series form (<= 1.1 ulp)
#endif
#endif
    return pow(x, y);
}
NPS_HD double py_pow(double x, double y) {
#if defined(__CUDA_ARCH__) && !defined(NPS_GENERIC_POW)
    if (y == 2.0) return x * x;
    if (y == 1.0) return x;
#if defined(NPS_POW_ONE_SHORTCUT)
    if (x == 1.0) return 1.0;               // pow(1, y) == 1 for every y (C99 F.9.4.4): the impeller / seal wear laws call it
#endif
    if (x > 0.0 && x < 1.7e308) {
        if (y == 0.5) return sqrt(x);
        // two correctly-rounded operations: <= 1 ulp from the exact power, inside libdevice pow's own 2-ulp bound
        if (y == 3.0) return x * x * x;
        if (y == 1.5) return x * sqrt(x);
        if (y == 0.25) return sqrt(sqrt(x));
    }
#endif
    return nps_pow_general(x, y);
}
// libm calls through one shared body each on the device (see NPS_HD_SHARED): log / log10 / exp expand to 60-100 SASS
// instructions per call site when inlined, and the kernel is instruction-fetch bound (20 % of stall samples).
NPS_HD_SHARED double nps_log(double x) { return log(x); }
NPS_HD_SHARED double nps_log10(double x) { return log10(x); }
NPS_HD_SHARED double nps_exp(double x) { return exp(x); }
NPS_HD double py_abs(double x) { return fabs(x); }
NPS_HD double np_sign(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : ((x == 0.0) ? 0.0 : x)); }
NPS_HD bool   is_true(double flag) { return flag != 0.0; }
NPS_HD double as_flag(bool b) { return b ? 1.0 : 0.0; }

}  // namespace nps
