// common.cuh -- shared helpers of librepconc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/repconc_b200.h"

#define RC_API extern "C" __attribute__((visibility("default")))

namespace rc {

void set_error(const char* fmt, ...);
void count_launch();

inline int check_cuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return RC_E_CUDA;
    }
    return RC_OK;
}

#define RC_CHECK_LAUNCH(name)                                              \
    do {                                                                   \
        ::rc::count_launch();                                              \
        int _rc = ::rc::check_cuda(cudaGetLastError(), name);              \
        if (_rc != RC_OK) return _rc;                                      \
    } while (0)

#define RC_REQUIRE(cond, ...)                  \
    do {                                       \
        if (!(cond)) {                         \
            ::rc::set_error(__VA_ARGS__);      \
            return RC_E_INVALID;               \
        }                                      \
    } while (0)

#define RC_CUDA(call)                                         \
    do {                                                      \
        int _rc = ::rc::check_cuda((call), #call);            \
        if (_rc != RC_OK) return _rc;                         \
    } while (0)

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;  // B200
    }
    return n;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// true the first time a call site runs on the current device: kernel function attributes (dynamic shared
// memory opt-in) are per device, and one process may drive several (e.g. index replicas on every GPU)
inline bool first_use_on_device(unsigned long long& seen) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return true;
    const unsigned long long bit = 1ull << (dev & 63);
    if (seen & bit) return false;
    seen |= bit;
    return true;
}

// ---------------------------------------------------------------------------------------------
// fp32 summation in the order of ATen's CPU sum kernel (inner contiguous reduction, 8-float
// vectors, 4 independent accumulators, cascade of 4 levels).  The reference's distance table is
// `((x - c)**2).sum(-1)` (modeling_repconc.py:50) on CPU tensors, so this order is what makes the
// fp32 table -- and with it every argmin / centring decision -- bit-identical to the reference.
// `f(j)` returns element j.  Explicit __fadd_rn: never contracted, never reassociated.
// ---------------------------------------------------------------------------------------------

// Compile-time length, n < 512 (no cascade level is ever completed below 512 elements).
template <int N, typename F>
__device__ __forceinline__ float sum_aten_order(F f) {
    static_assert(N >= 1 && N < 512, "use sum_aten_order_rt");
    if constexpr (N >= 8) {
        constexpr int NV = N / 8;        // number of 8-float vectors
        constexpr int NI = NV / 4;       // groups of 4 vectors (one per accumulator)
        float acc[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int l = 0; l < 8; ++l) acc[k][l] = 0.0f;
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int l = 0; l < 8; ++l) acc[k][l] = __fadd_rn(acc[k][l], f((i * 4 + k) * 8 + l));
#pragma unroll
        for (int i = NI * 4; i < NV; ++i)
#pragma unroll
            for (int l = 0; l < 8; ++l) acc[0][l] = __fadd_rn(acc[0][l], f(i * 8 + l));
#pragma unroll
        for (int k = 1; k < 4; ++k)
#pragma unroll
            for (int l = 0; l < 8; ++l) acc[0][l] = __fadd_rn(acc[0][l], acc[k][l]);
        float fin = 0.0f;
#pragma unroll
        for (int j = NV * 8; j < N; ++j) fin = __fadd_rn(fin, f(j));
#pragma unroll
        for (int l = 0; l < 8; ++l) fin = __fadd_rn(fin, acc[0][l]);
        return fin;
    } else {
        constexpr int NI = N / 4;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] = __fadd_rn(acc[k], f(i * 4 + k));
#pragma unroll
        for (int j = NI * 4; j < N; ++j) acc[0] = __fadd_rn(acc[0], f(j));
#pragma unroll
        for (int k = 1; k < 4; ++k) acc[0] = __fadd_rn(acc[0], acc[k]);
        return acc[0];
    }
}

// Run-time length (any n >= 1), including the cascade that starts at 512 elements.
// W = 8 (vector items) when n >= 8, else 1.  Slow path: only used for unusual sub-vector sizes.
template <typename F>
__device__ float sum_aten_order_rt(F f, int n) {
    const int W = n >= 8 ? 8 : 1;
    const int nitems = n / W;
    const int size = nitems / 4;
    float acc[4][4][8];
    for (int j = 0; j < 4; ++j)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < 8; ++l) acc[j][k][l] = 0.0f;
    int clog = 0;
    if (size > 1) {
        int x = size - 1;
        while (x > 0) { x >>= 1; ++clog; }
    }
    const int level_power = (clog / 4) > 4 ? (clog / 4) : 4;
    const int level_step = 1 << level_power;
    const int level_mask = level_step - 1;
    int i = 0;
    while (i + level_step <= size) {
        for (int j = 0; j < level_step; ++j, ++i)
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < W; ++l)
                    acc[0][k][l] = __fadd_rn(acc[0][k][l], f((i * 4 + k) * W + l));
        for (int j = 1; j < 4; ++j) {
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < W; ++l) {
                    acc[j][k][l] = __fadd_rn(acc[j][k][l], acc[j - 1][k][l]);
                    acc[j - 1][k][l] = 0.0f;
                }
            if ((i & (level_mask << (j * level_power))) != 0) break;
        }
    }
    for (; i < size; ++i)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < W; ++l) acc[0][k][l] = __fadd_rn(acc[0][k][l], f((i * 4 + k) * W + l));
    for (int j = 1; j < 4; ++j)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < W; ++l) acc[0][k][l] = __fadd_rn(acc[0][k][l], acc[j][k][l]);
    for (int it = size * 4; it < nitems; ++it)
        for (int l = 0; l < W; ++l) acc[0][0][l] = __fadd_rn(acc[0][0][l], f(it * W + l));
    for (int k = 1; k < 4; ++k)
        for (int l = 0; l < W; ++l) acc[0][0][l] = __fadd_rn(acc[0][0][l], acc[0][k][l]);
    if (W == 1) return acc[0][0][0];
    float fin = 0.0f;
    for (int j = nitems * 8; j < n; ++j) fin = __fadd_rn(fin, f(j));
    for (int l = 0; l < 8; ++l) fin = __fadd_rn(fin, acc[0][0][l]);
    return fin;
}

// (x - c)^2 in the reference's arithmetic: fp32 subtract, fp32 multiply, no contraction.
__device__ __forceinline__ float sqdiff(float x, float c) {
    const float d = __fsub_rn(x, c);
    return __fmul_rn(d, d);
}

// ---------------------------------------------------------------------------------------------
// Packed fp32 (Blackwell FADD2 / FFMA2: two IEEE-rn fp32 operations per instruction).  The distance kernels
// are bound by instruction issue (3 non-contractable ops per dimension); the packed forms halve the count
// without touching the arithmetic: every lane of sub/add.rn.f32x2 is the scalar rn operation, and the
// square is fma.rn(d, d, +0) = rn(d*d) -- one rounding, identical to mul.rn (a square is never -0).
// (mul.rn.f32x2 followed by add.rn.f32x2 is NOT used: ptxas contracts that pair into one FFMA2.)
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(f32x2_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t sub2_rn(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t add2_rn(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fma2_rn(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2_t sq2_rn(f32x2_t d) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(r) : "l"(d), "l"(0ull));
    return r;
}

// sum_j (x[j] - c[j])^2 in the order of sum_aten_order<N> for N a multiple of 8, two elements per instruction.
// x2(p) / c2(p) return elements (2p, 2p+1) packed.  Accumulators that are still exactly +0 are assigned
// instead of added to (0 + v == v), everything else is the scalar sequence operation for operation.
template <int N, typename FX, typename FC>
__device__ __forceinline__ float sqdist_aten_order_packed(FX x2, FC c2) {
    static_assert(N % 8 == 0 && N >= 8 && N < 512, "packed path: N = 8, 16, 24, ...");
    constexpr int NV = N / 8, NI = NV / 4;
    f32x2_t acc[4][4];
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int e = ((i * 4 + k) * 8) / 2 + p;
                const f32x2_t t = sq2_rn(sub2_rn(x2(e), c2(e)));
                acc[k][p] = i == 0 ? t : add2_rn(acc[k][p], t);
            }
#pragma unroll
    for (int i = NI * 4; i < NV; ++i)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int e = (i * 8) / 2 + p;
            const f32x2_t t = sq2_rn(sub2_rn(x2(e), c2(e)));
            acc[0][p] = (NI == 0 && i == 0) ? t : add2_rn(acc[0][p], t);
        }
    if (NI > 0) {
#pragma unroll
        for (int k = 1; k < 4; ++k)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[0][p] = add2_rn(acc[0][p], acc[k][p]);
    }
    float a[8];
#pragma unroll
    for (int p = 0; p < 4; ++p) unpk2(acc[0][p], a[2 * p], a[2 * p + 1]);
    float fin = a[0];                                   // 0 + a[0]
#pragma unroll
    for (int l = 1; l < 8; ++l) fin = __fadd_rn(fin, a[l]);
    return fin;
}

// warp reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// monotone float <-> uint32 mapping (larger float <=> larger key), NaN-free inputs
__device__ __forceinline__ uint32_t f32_to_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_f32(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// ---------------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA) helpers shared by the Sinkhorn passes and the ADC scan
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// an n-byte global -> shared bulk copy accounted on `bar` (the expect_tx for it must have been posted)
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// Ampere-style asynchronous copies (LDGSTS) for software pipelines whose tiles are too irregular for a bulk copy
__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace rc
