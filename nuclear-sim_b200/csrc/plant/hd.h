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
#else
#define NPS_HD inline
#define NPS_HD_NOINLINE
#endif

// Loops over repeated plant units (4 pumps, 3 SGs, 14 stages, 4 bearings) stay rolled on the device unless
// NPS_UNROLL_UNITS is defined: the step kernel is ~36 K SASS instructions and instruction-cache bound otherwise.
#if defined(__CUDA_ARCH__) && !defined(NPS_UNROLL_UNITS)
#define NPS_UNIT_LOOP _Pragma("unroll 1")
#else
#define NPS_UNIT_LOOP
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
NPS_HD double py_pow(double x, double y) { return pow(x, y); }
NPS_HD double py_abs(double x) { return fabs(x); }
NPS_HD double np_sign(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : ((x == 0.0) ? 0.0 : x)); }
NPS_HD bool   is_true(double flag) { return flag != 0.0; }
NPS_HD double as_flag(bool b) { return b ? 1.0 : 0.0; }

}  // namespace nps
