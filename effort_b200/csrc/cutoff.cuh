// cutoff.cuh -- the effort cutoff (reference kernel findCutoff32, bucketMul.metal:141-247).
//
// The reference bisects [min,max] of the 4096 scored probe products
//     bfloat(|1e5 * v[i] * bfloat(probes[i])|)
// until the number of products above the midpoint equals k = 4096 - q (or one of three give-up
// conditions fires) and returns the NEXT midpoint.  On real data the bisection usually does NOT
// terminate early: bf16 ties make count jump over k and the fp32 interval stalls at two adjacent
// floats (> 1e-5 apart), so the reference spins to its 100-loop cap (one threadgroup, 3 barriers per
// loop).  The result is still a deterministic function of the inputs, and it is reproduced here
// BIT-EXACTLY (same cutoff => same selected bucket rows as the oracle's literal restatement), but
// executed as:
//   phase A  block-wide counting rounds (REDUX popcount + ONE double-buffered barrier per round) while
//            more than 128 products lie inside (minBound, maxBound];
//   phase B  the <=128 inside products are compacted to shared memory and ONE warp finishes the
//            bisection with no block barriers;
//   fixpoint once the midpoint stops moving, later loops cannot change the state, so the remaining
//            iterations up to the reference's 100-loop cap are skipped (loops is reported as 101).
// It is a device function so that the fused bucketMul kernel runs it redundantly in every CTA (no extra
// launch, no global round trip for the scalar).
#pragma once
#include "common.cuh"

namespace effort {

constexpr int kCutoffThreads = 1024;
constexpr int kCutoffMaxPerThread = 8;  // n_probes <= 8192 for the stand-alone kernel
constexpr int kCutoffInsideMax = 128;   // phase-B capacity (4 per lane)

struct __align__(16) CutoffSmem {
    int cnt[2][32];
    float red_min[32];
    float red_max[32];
    float inside[kCutoffInsideMax];
    int n_inside;
    float result;
    int loops;
};

struct BisectState {
    float minBound, maxBound, newBound;
    int minCount, maxCount, loops;
};

// one reference iteration given countAbove; returns true when the loop exits (bucketMul.metal:199-246)
__device__ __forceinline__ bool bisect_step(BisectState& s, int countAbove, int effort, bool& fixpoint) {
    if (countAbove < effort) { s.maxBound = s.newBound; s.maxCount = countAbove; }
    else { s.minBound = s.newBound; s.minCount = countAbove; }
    const float prev = s.newBound;
    s.newBound = (s.maxBound + s.minBound) / 2;
    if (countAbove == effort || (s.maxBound - s.minBound < 0.00001f) || abs(s.maxCount - s.minCount) < 3)
        return true;
    if (s.loops > 100) return true;
    if (s.newBound == prev) {  // state can no longer change: the reference would idle to loops == 101
        fixpoint = true;
        s.loops = 101;
        return true;
    }
    return false;
}

// Block-cooperative bisection.  Every thread of the block (blockDim.x multiple of 32, <= 1024) calls this
// with its PER scored values in registers (vals[k] < 0 marks "no value").  Returns the cutoff in every
// thread.  `loops_out` (may be null) receives the reference's iteration count from thread 0.
template <int PER>
__device__ __forceinline__ float block_bisect_cutoff(const float (&vals)[PER], int n_probes, int q,
                                                     CutoffSmem& sm, int* loops_out,
                                                     unsigned long long* trace = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    float tmin = 999.f, tmax = -999.f;  // bucketMul.metal:155-156
#pragma unroll
    for (int k = 0; k < PER; k++)
        if (vals[k] >= 0.f) { tmin = fminf(tmin, vals[k]); tmax = fmaxf(tmax, vals[k]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    }
    if (lane == 0) { sm.red_min[warp] = tmin; sm.red_max[warp] = tmax; }
    if (tid == 0) sm.n_inside = 0;
    if (tid < 64) (&sm.cnt[0][0])[tid] = 0;  // slots of absent warps must read as 0
    __syncthreads();
    float gmin = (lane < nwarps) ? sm.red_min[lane] : 999.f;
    float gmax = (lane < nwarps) ? sm.red_max[lane] : -999.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
        gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    }
    BisectState s;
    // tgMin/tgMax are bfloat in the reference (:172-181): the 999 sentinel becomes 1000.
    s.minBound = bf16_round(gmin);
    s.maxBound = bf16_round(gmax);
    s.newBound = (s.minBound + s.maxBound) / 2;
    s.loops = 0; s.minCount = 4096; s.maxCount = 0;  // literals of the kernel (:197, :168-169)
    const int effort = n_probes - q;                 // :154
    bool done = false, fixpoint = false;
    int buf = 0;
    if (trace && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); trace[4] = t; }
    // ---- phase A: block-wide rounds (state replicated in every thread, identical by construction) ----
    while (!done && (s.minCount - s.maxCount) > kCutoffInsideMax) {
        s.loops++;
        int c = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) c += (vals[k] > s.newBound) ? 1 : 0;  // negatives never count
        c = warp_sum_i(c);
        if (lane == 0) sm.cnt[buf][warp] = c;
        __syncthreads();
        int countAbove = 0;
#pragma unroll
        for (int w4 = 0; w4 < 8; w4++) {  // 32 warp counts, broadcast 16-byte reads (unused slots are 0)
            const int4 q4 = *reinterpret_cast<const int4*>(&sm.cnt[buf][w4 * 4]);
            countAbove += q4.x + q4.y + q4.z + q4.w;
        }
        buf ^= 1;
        done = bisect_step(s, countAbove, effort, fixpoint);
    }
    if (trace && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); trace[5] = t; trace[11] = (unsigned long long)s.loops; }
    if (!done) {
        // ---- compact the products inside (minBound, maxBound] ----
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (vals[k] > s.minBound && vals[k] <= s.maxBound) {
                const int p = atomicAdd(&sm.n_inside, 1);
                if (p < kCutoffInsideMax) sm.inside[p] = vals[k];
            }
        }
        __syncthreads();
        // ---- phase B: one warp, no block barriers ----
        if (warp == 0) {
            const int n_in = min(sm.n_inside, kCutoffInsideMax);
            float x[kCutoffInsideMax / 32];
#pragma unroll
            for (int k = 0; k < kCutoffInsideMax / 32; k++) {
                const int i = lane + 32 * k;
                x[k] = (i < n_in) ? sm.inside[i] : -1.f;
            }
            // products above maxBound keep counting; maxCount is their exact number (0 if never set)
            const int above_max = s.maxCount;
            while (!done) {
                s.loops++;
                int c = above_max;
#pragma unroll
                for (int k = 0; k < kCutoffInsideMax / 32; k++)
                    c += __popc(__ballot_sync(0xffffffffu, x[k] > s.newBound));
                const int countAbove = c;
                done = bisect_step(s, countAbove, effort, fixpoint);
            }
            if (lane == 0) { sm.result = s.newBound; sm.loops = s.loops; }
        }
        __syncthreads();
        s.newBound = sm.result;
        s.loops = sm.loops;
    }
    if (loops_out && tid == 0) *loops_out = s.loops;
    if (trace && tid == 0) trace[12] = (unsigned long long)s.loops;
    return s.newBound;
}

// Scores this thread's probes: thread t owns probes t, t+NT, ...  (bucketMul.metal:158-163)
template <int PER>
__device__ __forceinline__ void score_probes(const float* __restrict__ v, const __half* __restrict__ probes,
                                             uint32_t exp_no, int n_probes, float (&vals)[PER],
                                             uint64_t keep) {
#pragma unroll
    for (int k = 0; k < PER; k++) {
        int i = threadIdx.x + k * blockDim.x;
        if (i < n_probes) {
            const uint16_t pb16 = ldg_keep_u16(probes + (size_t)exp_no * n_probes + i, keep);
            float p = bf16_round(__half2float(__ushort_as_half(pb16)));
            float x = __fmul_rn(__fmul_rn(kCutoffScale, v[i]), p);
            vals[k] = bf16_round(fabsf(x));
        } else {
            vals[k] = -1.f;
        }
    }
}

// Stand-alone launch (test hook effort_find_cutoff, and the first stage of the unfused path).
__global__ void __launch_bounds__(kCutoffThreads, 1)
find_cutoff_kernel(const float* __restrict__ v, const __half* __restrict__ probes,
                   const uint32_t* __restrict__ exp_no_dev, int n_probes, int q,
                   float* __restrict__ cutoff_out, int* __restrict__ loops_out) {
    __shared__ CutoffSmem sm;
    const uint32_t exp_no = exp_no_dev ? *exp_no_dev : 0u;
    float vals[kCutoffMaxPerThread];
    score_probes<kCutoffMaxPerThread>(v, probes, exp_no, n_probes, vals, l2_policy_evict_last());
    float c = block_bisect_cutoff<kCutoffMaxPerThread>(vals, n_probes, q, sm, loops_out);
    if (threadIdx.x == 0) *cutoff_out = c;
}

}  // namespace effort
