// exact_math.cuh -- correctly rounded FP64 reciprocal / division / square root WITHOUT branches.
//
// Why: on B200 the FP64 pipe issues one warp instruction per 2 cycles per SM sub-partition and a
// dependent result is ready ~8 cycles later.  The FP64-heavy force sweeps run 4 warps per
// sub-partition (128 registers), and round-1 ncu shows them waiting on fixed-latency dependencies
// (`stall_wait`), FP64 pipe only 47 % busy: the instruction streams have no ILP.  The hardware
// division and sqrt sequences each end in a conditional call to a slow path, which chops the code
// into basic blocks the scheduler cannot interleave across -- so the two particles a thread owns
// are processed one after the other.  The sequences below are straight-line; special operands are
// not handled inline but REPORTED in a flag word, and the caller redoes such a particle with the
// ordinary IEEE operators.
//
// Correct rounding (so results are bit-identical to `/` and `sqrt`):
//   rcp   seed (rcp.approx.ftz.f64, ~2^-20) -> two Newton steps y += y(1-by) -> y within 1 ulp;
//         one more step y' = RN(y + y RN(1-by)) is Markstein's corrected reciprocal = RN(1/b)
//         unless b's significand is all ones (flagged).
//   div   y = RN(1/b):  q0 = RN(ay); q1 = RN(q0 + RN(a-bq0) y) is faithful;  q = RN(q1 + RN(a-bq1) y)
//         = RN(a/b) by Markstein's theorem (residuals are exact under an FMA).
//   sqrt  seed h = rsqrt.approx/2, g = x*rsqrt;  two coupled Newton steps (r = 1/2 - gh; g += gr;
//         h += hr);  final g' = RN(g + h RN(x - g^2)): the Markstein/Cornea-Harrison-Tang square root.
// (P. Markstein, IBM J. Res. Dev. 34 (1990); Muller et al., Handbook of Floating-Point Arithmetic,
// ch. "Division and square root using FMA".)  tools/exact_math_check.cu compares all three with the
// hardware operators on billions of random and adversarial operands on the GPU.
//
// Operand requirements for the straight-line path (anything else sets the flag): divisor and
// radicand finite, non-zero, exponent within +-400 of 1; numerators zero or within +-400.
#pragma once
#include <cstdint>

namespace rebxb {

__device__ __forceinline__ unsigned exponent_out_of_range(double v) {      // 1 if |v| is not in (2^-400, 2^400)
    const unsigned e = ((unsigned)__double2hiint(v) >> 20) & 0x7ffu;
    return (e - 624u) > 798u ? 1u : 0u;
}
__device__ __forceinline__ unsigned significand_all_ones(double v) {
    return (((unsigned long long)__double_as_longlong(v) & 0x000fffffffffffffULL) == 0x000fffffffffffffULL) ? 1u : 0u;
}

__device__ __forceinline__ double rcp_seed(double b) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    return y;
}
__device__ __forceinline__ double rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

struct ExactRecip { double b, y; };

// RN(1/b); sets bit 0 of `bad` if b is outside the straight-line domain.
__device__ __forceinline__ ExactRecip exact_rcp(double b, unsigned& bad) {
    bad |= exponent_out_of_range(b) | significand_all_ones(b);
    double y = rcp_seed(b);
    double e = fma(-b, y, 1.0);  y = fma(y, e, y);
    e = fma(-b, y, 1.0);         y = fma(y, e, y);
    e = fma(-b, y, 1.0);         y = fma(y, e, y);
    return ExactRecip{b, y};
}

// RN(a/b) given R = exact_rcp(b).
__device__ __forceinline__ double exact_div(double a, const ExactRecip& R, unsigned& bad) {
    bad |= (a != 0.0) ? exponent_out_of_range(a) : 0u;
    const double q0 = a*R.y;
    const double r0 = fma(-R.b, q0, a);
    const double q1 = fma(r0, R.y, q0);
    const double r1 = fma(-R.b, q1, a);
    const double q = fma(r1, R.y, q1);
    return (a == 0.0) ? q0 : q;          // a = +-0: the product already carries the right sign of zero
}

// RN(sqrt(x)).
__device__ __forceinline__ double exact_sqrt(double x, unsigned& bad) {
    bad |= exponent_out_of_range(x) | ((x < 0.0) ? 1u : 0u);
    const double s = rsqrt_seed(x);
    double g = x*s, h = 0.5*s;
    double r = fma(-g, h, 0.5);  g = fma(g, r, g);  h = fma(h, r, h);
    r = fma(-g, h, 0.5);         g = fma(g, r, g);  h = fma(h, r, h);
    const double d = fma(-g, g, x);
    return fma(d, h, g);
}

}  // namespace rebxb
