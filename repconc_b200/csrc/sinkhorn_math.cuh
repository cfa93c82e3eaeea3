// sinkhorn_math.cuh -- branch-free fp64 exp2 for the Sinkhorn passes.
//
// The passes are bound by the FP64 pipe, not by HBM (one exp per table element per iteration), and
// CUDA's exp() compiles to a serial 13-DFMA chain with a range-check branch per call, which leaves
// the pipe idle most of the time at the occupancy this kernel runs at.  This version has no
// branches, so the 8 evaluations a lane makes per table row interleave in one basic block, and it
// works in base 2 (the scaling vectors are kept in log2 units), which makes the argument reduction
// exact and two operations shorter:
//     n = rint(w)  (magic-number add),  f = w - n  in [-1/2, 1/2]  (exact),
//     2^w = 2^n * p(f),  p = degree-11 near-minimax polynomial (tools/gen_exp2_poly.py:
//     max relative error 2.0e-17 in exact arithmetic, i.e. < 0.2 ulp before Horner rounding).
// Underflow: n is clamped at -1022, results below ~2e-308 come out as some value <= 2.3e-308
// instead of the exact denormal -- they are added to sums of magnitude >= 1e-10, invisible in fp64.
// Overflow (w >= 1024) cannot happen inside the iteration (Q <= 1); the one place it can
// (Q0 = exp(-d~/eps) with eps < 1/709) is checked by the caller, which raises RC_FLAG_NONFINITE.
#pragma once

namespace rc {

__device__ __forceinline__ double exp2_fast(double w) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
    const double t = w + MAGIC;
    int n = __double2loint(t);
    const double f = w - (t - MAGIC);
    double p = 4.4558179083360645e-10;
    p = fma(p, f, 7.074194297288521e-09);
    p = fma(p, f, 1.0178057087733941e-07);
    p = fma(p, f, 1.3215432535912375e-06);
    p = fma(p, f, 1.5252733841556773e-05);
    p = fma(p, f, 0.00015403530463724353);
    p = fma(p, f, 0.001333355814640647);
    p = fma(p, f, 0.009618129107587256);
    p = fma(p, f, 0.055504108664821625);
    p = fma(p, f, 0.24022650695910158);
    p = fma(p, f, 0.6931471805599453);
    p = fma(p, f, 1.0);
    n = max(n, -1022);
    n = min(n, 1023);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

constexpr double RC_LOG2E = 1.4426950408889634074;
constexpr double RC_PAD_LOG2 = -4000.0;  // log2 scaling of padded (k >= K) lanes: 2^w clamps to ~0

}  // namespace rc
