// comm.cuh -- one-shot all-gather / all-reduce over NVLink peer memory.
//
// The exchanges of the tensor-parallel decode loop are 16-57 KB: pure latency.  Each rank owns a "symmetric"
// buffer (same layout everywhere) that every peer maps through CUDA IPC.  A collective is ONE kernel per rank and
// ONE NVLink hop, with the low-latency packet protocol: every float travels as an 8-byte {value, seq} packet
// written with a single 64-bit store (atomic on the wire), so the receiver polls the data words themselves --
// no system fence, no separate flag round trip:
//   1. thread i stores {x[i], seq} into slot [my rank][i] of the site region of every peer;
//   2. thread i polls slot [r][i] of ITS OWN region for every other rank r until the packet carries seq;
//   3. all-gather copies the values out, all-reduce sums the per-source slots in rank order (deterministic and
//      bit-identical on every rank; the local contribution is taken from registers).
// seq lives in device memory and is bumped by the kernel itself, so the launch is CUDA-graph replayable and a
// stale packet of an earlier round can never match.  A site (one call position in the layer) has its own region:
// a peer overwrites slot [r] of a site only after it passed a LATER site's wait, which this rank feeds only after
// its kernel for the earlier site has finished reading (stream order), so packets are never overwritten early.
//
// The kernels that matter fuse the exchange with the elementwise work around it (one CTA each):
//   p2p_allreduce_residual_rmsnorm_kernel : h += sum_r partial_r ; out = rmsNorm(h) * w     (after wo / w2)
//   p2p_silu_allgather_kernel             : x2 = silu(x1) * x3 on the local slice, gathered for w2's cutoff
#pragma once
#include "common.cuh"

namespace effort {

constexpr int kP2PMaxRanks = 16;
constexpr int kP2PSites = 16;
constexpr size_t kP2PSiteBytes = 1u << 20;  // 1 MiB per site = 131072 packets over all source ranks
constexpr int kP2PBlocks = 1;     // 16-57 KB messages: one CTA, every thread owns its packets end to end
constexpr int kP2PThreads = 1024;

struct P2PLayout {  // offsets inside a rank's symmetric buffer
    static constexpr size_t seq_off = 0;         // [sites] u32 (local use)
    static constexpr size_t data_off = 4096;     // [sites][kP2PSiteBytes] packets
    static constexpr size_t total = data_off + kP2PSites * kP2PSiteBytes;
};

struct P2PArgs {
    unsigned char* peer[kP2PMaxRanks];  // every rank's symmetric buffer (peer[rank] = own)
    int rank, world, site;
    unsigned* err;  // the context's error flag: a wait that gives up (peer died / never launched) writes 4 here
};

__device__ __forceinline__ void ll_store(unsigned long long* p, float v, unsigned seq) {
    const unsigned long long pkt = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(pkt) : "memory");
}
// bounded: a peer that never sends (crashed process, mismatched launch order) must not hang this GPU.  After ~2 s the
// wait gives up, raises the context's error flag (read with effort_ctx_error_flag: 4) and returns 0.
__device__ __forceinline__ float ll_wait(const unsigned long long* p, unsigned seq, unsigned* err) {
    unsigned long long pkt;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(pkt) : "l"(p) : "memory");
    if ((unsigned)(pkt >> 32) == seq) return __uint_as_float((unsigned)pkt);
    const unsigned long long t0 = gtime_ns();
    for (;;) {
#pragma unroll 1
        for (int k = 0; k < 256; k++) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(pkt) : "l"(p) : "memory");
            if ((unsigned)(pkt >> 32) == seq) return __uint_as_float((unsigned)pkt);
        }
        if (gtime_ns() - t0 > 2000000000ull) {
            if (err) atomicExch(err, 4u);
            return 0.f;
        }
    }
}
__device__ __forceinline__ unsigned long long* ll_slot(const P2PArgs& a, int peer, int src_rank, int count) {
    return reinterpret_cast<unsigned long long*>(a.peer[peer] + P2PLayout::data_off + (size_t)a.site * kP2PSiteBytes) +
           (size_t)src_rank * count;
}

// mode 0: all-gather  (send: count floats;  out: world*count floats, rank-major)
// mode 1: all-reduce  (send: count floats;  out: count floats = sum over ranks, rank order)
template <int MODE>
__global__ void __launch_bounds__(kP2PThreads)
p2p_collective_kernel(const P2PArgs a, const float* __restrict__ send, float* __restrict__ out, int count) {
    pdl_trigger();
    pdl_wait();
    unsigned* seq_ptr = reinterpret_cast<unsigned*>(a.peer[a.rank] + P2PLayout::seq_off) + a.site;
    const unsigned seq = *seq_ptr + 1u;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const float x = send[i];
        for (int p = 0; p < a.world; p++)
            if (p != a.rank) ll_store(ll_slot(a, p, a.rank, count) + i, x, seq);
        if (MODE == 0) out[(size_t)a.rank * count + i] = x;
    }
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < a.world; r++) {
            if (MODE == 0) {
                if (r != a.rank) out[(size_t)r * count + i] = ll_wait(ll_slot(a, a.rank, r, count) + i, seq, a.err);
            } else {
                s += (r == a.rank) ? send[i] : ll_wait(ll_slot(a, a.rank, r, count) + i, seq, a.err);
            }
        }
        if (MODE == 1) out[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) *seq_ptr = seq;
}

// all-reduce of the row-parallel GEMV output fused with what follows it in the layer (runNetwork.swift:170-175,
// 182-183 + the next rmsNorm): h += sum_r partial_r ;  out_norm = rmsNormFast(h) * w.   dim <= 4 * kP2PThreads.
__global__ void __launch_bounds__(kP2PThreads)
p2p_allreduce_residual_rmsnorm_kernel(const P2PArgs a, const float* __restrict__ partial, float* __restrict__ h,
                                      const __half* __restrict__ w, int dim, float eps, float* __restrict__ out_norm) {
    __shared__ float red[32];
    __shared__ float total;
    pdl_trigger();
    pdl_wait();
    unsigned* seq_ptr = reinterpret_cast<unsigned*>(a.peer[a.rank] + P2PLayout::seq_off) + a.site;
    const unsigned seq = *seq_ptr + 1u;
    float mine[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = threadIdx.x + k * blockDim.x;
        mine[k] = 0.f;
        if (i < dim) {
            mine[k] = partial[i];
            for (int p = 0; p < a.world; p++)
                if (p != a.rank) ll_store(ll_slot(a, p, a.rank, dim) + i, mine[k], seq);
        }
    }
    float x[4];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = threadIdx.x + k * blockDim.x;
        x[k] = 0.f;
        if (i < dim) {
            float s = 0.f;
            for (int r = 0; r < a.world; r++) s += (r == a.rank) ? mine[k] : ll_wait(ll_slot(a, a.rank, r, dim) + i, seq, a.err);
            x[k] = h[i] + s;
            h[i] = x[k];
            ss += x[k] * x[k];
        }
    }
    ss = warp_sum_f(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
        t = warp_sum_f(t);
        if (threadIdx.x == 0) { total = t; *seq_ptr = seq; }
    }
    __syncthreads();
    const float denom = sqrtf(total / (float)dim + eps);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = threadIdx.x + k * blockDim.x;
        if (i < dim) out_norm[i] = (x[k] / denom) * __half2float(w[i]);
    }
}

// silu*mul (matrix.metal:25-34) on this rank's hidden slice fused with the all-gather that the row-parallel w2
// needs for its cutoff: x2_local = x3 * x1 / (1 + exp(-x1));  x2_full = concat over ranks.
__global__ void __launch_bounds__(kP2PThreads)
p2p_silu_allgather_kernel(const P2PArgs a, const float* __restrict__ x1, const float* __restrict__ x3, int n_local,
                          float* __restrict__ x2_local, float* __restrict__ x2_full) {
    pdl_trigger();
    pdl_wait();
    unsigned* seq_ptr = reinterpret_cast<unsigned*>(a.peer[a.rank] + P2PLayout::seq_off) + a.site;
    const unsigned seq = *seq_ptr + 1u;
    for (int i = threadIdx.x; i < n_local; i += blockDim.x) {
        const float v = x3[i] * x1[i] / (1.f + expf(-x1[i]));
        for (int p = 0; p < a.world; p++)
            if (p != a.rank) ll_store(ll_slot(a, p, a.rank, n_local) + i, v, seq);
        x2_local[i] = v;
        x2_full[(size_t)a.rank * n_local + i] = v;
    }
    for (int r = 0; r < a.world; r++) {
        if (r == a.rank) continue;
        for (int i = threadIdx.x; i < n_local; i += blockDim.x)
            x2_full[(size_t)r * n_local + i] = ll_wait(ll_slot(a, a.rank, r, n_local) + i, seq, a.err);
    }
    __syncthreads();
    if (threadIdx.x == 0) *seq_ptr = seq;
}

}  // namespace effort
