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

// N evaluations with the Horner recurrences interleaved coefficient by coefficient (N independent
// dependency chains in flight: the fp64 pipe has ~16 cycles of latency and this kernel runs 4 warps per
// scheduler).  CLAMP_HI = false when the caller guarantees w < 1024 (inside the iteration Q <= 1).
template <int N, bool CLAMP_HI>
__device__ __forceinline__ void exp2_fast_batch(double (&w)[N]) {
    const double MAGIC = 6755399441055744.0;
    int n[N];
    double f[N], p[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double t = w[i] + MAGIC;
        n[i] = __double2loint(t);
        f[i] = w[i] - (t - MAGIC);
        p[i] = fma(4.4558179083360645e-10, f[i], 7.074194297288521e-09);
    }
#define RC_HORNER(C)                                   \
    _Pragma("unroll") for (int i = 0; i < N; ++i) p[i] = fma(p[i], f[i], C);
    RC_HORNER(1.0178057087733941e-07)
    RC_HORNER(1.3215432535912375e-06)
    RC_HORNER(1.5252733841556773e-05)
    RC_HORNER(0.00015403530463724353)
    RC_HORNER(0.001333355814640647)
    RC_HORNER(0.009618129107587256)
    RC_HORNER(0.055504108664821625)
    RC_HORNER(0.24022650695910158)
    RC_HORNER(0.6931471805599453)
    RC_HORNER(1.0)
#undef RC_HORNER
#pragma unroll
    for (int i = 0; i < N; ++i) {
        int e = max(n[i], -1022);
        if (CLAMP_HI) e = min(e, 1023);
        w[i] = __hiloint2double(__double2hiint(p[i]) + (e << 20), __double2loint(p[i]));
    }
}

// Latency-oriented single evaluation (Estrin's scheme: dependency depth 5 instead of 11) for the sparse
// pass, where a lane evaluates ONE significant element per table row and the chain latency, not the
// fp64 pipe, is what a row costs.  Same coefficients; w < 1024 guaranteed by the caller.
// coefficients in the constant bank: DFMA takes a c[][] operand directly, so the sparse pass (which runs at
// 3 CTAs/SM and cannot keep 12 fp64 constants in registers) does not rematerialise them every row
__constant__ double RC_EXP2_C[12] = {
    1.0, 0.6931471805599453, 0.24022650695910158, 0.055504108664821625, 0.009618129107587256,
    0.001333355814640647, 0.00015403530463724353, 1.5252733841556773e-05, 1.3215432535912375e-06,
    1.0178057087733941e-07, 7.074194297288521e-09, 4.4558179083360645e-10};

__device__ __forceinline__ double exp2_fast_estrin(double w) {
    const double MAGIC = 6755399441055744.0;
    const double t = w + MAGIC;
    const int n = max(__double2loint(t), -1022);
    const double f = w - (t - MAGIC);
    const double f2 = f * f;
    const double p01 = fma(RC_EXP2_C[1], f, RC_EXP2_C[0]);
    const double p23 = fma(RC_EXP2_C[3], f, RC_EXP2_C[2]);
    const double p45 = fma(RC_EXP2_C[5], f, RC_EXP2_C[4]);
    const double p67 = fma(RC_EXP2_C[7], f, RC_EXP2_C[6]);
    const double p89 = fma(RC_EXP2_C[9], f, RC_EXP2_C[8]);
    const double pab = fma(RC_EXP2_C[11], f, RC_EXP2_C[10]);
    const double f4 = f2 * f2;
    const double q0 = fma(p23, f2, p01);
    const double q1 = fma(p67, f2, p45);
    const double q2 = fma(pab, f2, p89);
    const double f8 = f4 * f4;
    const double r0 = fma(q1, f4, q0);
    const double p = fma(q2, f8, r0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

constexpr double RC_LOG2E = 1.4426950408889634074;
constexpr double RC_PAD_LOG2 = -4000.0;  // log2 scaling of padded (k >= K) lanes: 2^w clamps to ~0

}  // namespace rc
