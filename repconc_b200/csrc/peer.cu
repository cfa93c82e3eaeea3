// peer.cu -- one-shot all-reduce (SUM, fp64) of the Sinkhorn row sums over NVLink peer memory.
//
// The reference all-reduces the (M,K) fp64 row sums once per Sinkhorn iteration
// (modeling_repconc.py:156-157): 98 KB at M=48, strictly serialised with the passes, i.e. pure latency.
// NCCL costs ~50 us per call on 8 GPUs; this kernel does the exchange itself on NVSwitch-connected peers:
// every rank publishes its vector in a symmetric buffer (peer-mapped by the caller, e.g. torch
// symmetric memory), raises a sequence flag on every peer with a system-scope release store, waits for the
// W flags addressed to it, and then sums the W published vectors IN RANK ORDER from peer memory -- every
// rank computes the bitwise identical result.  Two slots alternate by sequence parity, so a rank may run
// one exchange ahead of the slowest peer without overwriting data that is still being read.
//
// Buffer layout per rank (caller-allocated, zero-initialised, peer-mapped):
//   [0, 1024)                    uint32 flags[2 slots][8 slices][16]  (a slice = one CTA's share of the vector,
//                                exchanged independently of the other slices)
//   [1024, 1024 + 2*n*8)         double slot[2][n]
#include <stdlib.h>

#include "common.cuh"

namespace rc {

constexpr int PEER_THREADS = 512;
constexpr int PEER_MAX_W = 16;
constexpr int PEER_UNR = 4;         // elements per thread whose W peer loads are in flight together
constexpr int PEER_SLICES = 8;      // independent CTAs, each exchanging its own slice with its own flags
constexpr int PEER_HEADER = 2 * PEER_SLICES * PEER_MAX_W * 4;   // uint32 flags[2 slots][slices][W]

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// data loads: system-scope relaxed (never served from a stale L1 line); deliberately NOT volatile / no memory
// clobber so that the W x UNROLL loads of a thread are all in flight together (the barrier before them is the
// ordering point)
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
    double v;
    asm("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct PeerPtrs {
    unsigned char* base[PEER_MAX_W];
};

template <int W_T>   // 0 = run-time W
__global__ void __launch_bounds__(PEER_THREADS)
peer_allreduce_f64_kernel(PeerPtrs peers, int rank, int W_rt, int64_t n, uint32_t seq, double* __restrict__ inout,
                          int32_t* __restrict__ flags, unsigned long long timeout_ns) {
    const int W = W_T > 0 ? W_T : W_rt;
    const int slot = (int)(seq & 1u);
    const int slice = blockIdx.x;
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)slice * per, hi = min(n, lo + per);
    double* mine = reinterpret_cast<double*>(peers.base[rank] + PEER_HEADER) + (int64_t)slot * n;
    // 1. publish this slice
    for (int64_t i = lo + threadIdx.x; i < hi; i += PEER_THREADS) mine[i] = inout[i];
    __threadfence_system();
    __syncthreads();
    // 2. signal every peer, 3. wait for every peer (time-bounded: a lost peer must not hang the GPU)
    __shared__ int s_timeout;
    if (threadIdx.x == 0) s_timeout = 0;
    __syncthreads();
    if (threadIdx.x < W) {
        const int fidx = (slot * PEER_SLICES + slice) * PEER_MAX_W;
        st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x]) + fidx + rank, seq);
        const uint32_t* my = reinterpret_cast<const uint32_t*>(peers.base[rank]) + fidx + threadIdx.x;
        const unsigned long long t0 = global_ns();
        while ((int32_t)(ld_acquire_sys(my) - seq) < 0) {   // sequence numbers only grow (wrap-safe compare)
            if (global_ns() - t0 > timeout_ns) { s_timeout = 1; break; }
        }
    }
    __syncthreads();
    if (s_timeout) {
        if (threadIdx.x == 0) atomicOr(flags, 16);   // RC_FLAG_PEER_TIMEOUT
        return;
    }
    // 4. sum in rank order straight from peer memory, PEER_UNR elements x W loads in flight per thread
    const int64_t soff = (int64_t)slot * n;
    for (int64_t i = lo + threadIdx.x; i < hi; i += (int64_t)PEER_UNR * PEER_THREADS) {
        double v[PEER_UNR][PEER_MAX_W];
#pragma unroll
        for (int u = 0; u < PEER_UNR; ++u) {
            const int64_t iu = i + (int64_t)u * PEER_THREADS;
#pragma unroll
            for (int p = 0; p < PEER_MAX_W; ++p) {
                if (p < W) {
                    const double* sp = reinterpret_cast<const double*>(peers.base[p] + PEER_HEADER) + soff;
                    v[u][p] = iu < hi ? ld_relaxed_sys_f64(sp + iu) : 0.0;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PEER_UNR; ++u) {
            const int64_t iu = i + (int64_t)u * PEER_THREADS;
            double sum = 0.0;
#pragma unroll
            for (int p = 0; p < PEER_MAX_W; ++p)
                if (p < W) sum += v[u][p];
            if (iu < hi) inout[iu] = sum;
        }
    }
}

}  // namespace rc

using namespace rc;

RC_API size_t rc_peer_allreduce_buffer_bytes(int64_t n) { return PEER_HEADER + 2 * (size_t)n * 8; }

RC_API int rc_peer_allreduce_f64(const uint64_t* peer_buffers_host, int rank, int W, int64_t n, uint32_t seq,
                                 double* inout, int32_t* flags, void* stream) {
    RC_REQUIRE(peer_buffers_host && inout && flags, "rc_peer_allreduce_f64: null pointer");
    RC_REQUIRE(W >= 1 && W <= PEER_MAX_W && rank >= 0 && rank < W && n >= 1 && seq >= 1,
               "rc_peer_allreduce_f64: bad argument (W=%d rank=%d n=%lld seq=%u)", W, rank, (long long)n, seq);
    PeerPtrs pp{};
    for (int p = 0; p < W; ++p) {
        RC_REQUIRE(peer_buffers_host[p] != 0, "rc_peer_allreduce_f64: null peer buffer %d", p);
        pp.base[p] = reinterpret_cast<unsigned char*>(peer_buffers_host[p]);
    }
    int slices = (int)((n + 1023) / 1024);
    if (slices > PEER_SLICES) slices = PEER_SLICES;
    cudaStream_t st = (cudaStream_t)stream;
    // bound of the wait for a peer: env RC_PEER_TIMEOUT_MS (default 30 s -- a rank may sit in a checkpoint, a
    // garbage collection or a debugger); on expiry RC_FLAG_PEER_TIMEOUT is raised instead of hanging the GPU
    static unsigned long long tmo = 0ull;
    if (tmo == 0ull) {
        const char* e = getenv("RC_PEER_TIMEOUT_MS");
        const double ms = (e && e[0]) ? atof(e) : 30000.0;
        tmo = (unsigned long long)((ms > 1.0 ? ms : 1.0) * 1e6);
    }
    switch (W) {
        case 2: peer_allreduce_f64_kernel<2><<<slices, PEER_THREADS, 0, st>>>(pp, rank, W, n, seq, inout, flags, tmo); break;
        case 4: peer_allreduce_f64_kernel<4><<<slices, PEER_THREADS, 0, st>>>(pp, rank, W, n, seq, inout, flags, tmo); break;
        case 8: peer_allreduce_f64_kernel<8><<<slices, PEER_THREADS, 0, st>>>(pp, rank, W, n, seq, inout, flags, tmo); break;
        default: peer_allreduce_f64_kernel<0><<<slices, PEER_THREADS, 0, st>>>(pp, rank, W, n, seq, inout, flags, tmo);
    }
    RC_CHECK_LAUNCH("peer_allreduce_f64_kernel");
    return RC_OK;
}
