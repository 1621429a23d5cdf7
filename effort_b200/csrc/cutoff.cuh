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
//   phase A  group-wide counting rounds while more than 128 products lie inside (minBound, maxBound];
//            every round performs TWO reference iterations: besides the count at the current midpoint it
//            counts at both midpoints the next iteration could pick ((mid+min)/2 and (max+mid)/2 -- the
//            same fp32 expressions the reference evaluates), so the second step needs no synchronisation;
//   phase B  the <=128 inside products are compacted to shared memory and ONE warp finishes with no group
//            barriers; as soon as everything left inside is one repeated value the count is a step function
//            of the midpoint and the remaining halvings down to two adjacent floats cost a compare each;
//   fixpoint once the midpoint stops moving, later loops cannot change the state, so the remaining
//            iterations up to the reference's 100-loop cap are skipped (loops is reported as 101).
// The pieces are device functions over a "group" of warps (the whole CTA, or the selector warps of the
// fused bucketMul kernel, which keep bisecting while the other warps already stream certain rows).
#pragma once
#include "common.cuh"

namespace effort {

constexpr int kCutoffThreads = 1024;
constexpr int kCutoffMaxPerThread = 8;  // n_probes <= 8192 for the stand-alone kernel
constexpr int kCutoffInsideMax = 128;   // phase-B capacity (4 per lane)

struct __align__(16) CutoffSmem {
    int cnt2[2][2][32];  // [buffer][count at mid | packed counts at the two next mids][warp of the group]
    float red_min[32];
    float red_max[32];
    float inside[kCutoffInsideMax];
    int n_inside;
    float result;
    int loops;
    // direct method (block_cutoff_direct)
    unsigned hist[2048];   // products per 16-code bin of the 15-bit bf16 key
    int sub[5][16];        // per target rank: products per code inside its bin
    int wsum[32];
    int tbin[5], trank[5];
};

struct BisectState {
    float minBound, maxBound, newBound;
    int minCount, maxCount, loops;
};

// barrier over a group of warps: id 0 = the whole CTA
__device__ __forceinline__ void group_bar(int id, int n_threads) {
    if (id == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// one reference iteration given countAbove; returns true when the loop exits (bucketMul.metal:199-246)
__device__ __forceinline__ bool bisect_step(BisectState& s, int countAbove, int effort) {
    if (countAbove < effort) { s.maxBound = s.newBound; s.maxCount = countAbove; }
    else { s.minBound = s.newBound; s.minCount = countAbove; }
    const float prev = s.newBound;
    s.newBound = (s.maxBound + s.minBound) / 2;
    if (countAbove == effort || (s.maxBound - s.minBound < 0.00001f) || abs(s.maxCount - s.minCount) < 3)
        return true;
    if (s.loops > 100) return true;
    if (s.newBound == prev) {  // state can no longer change: the reference would idle to loops == 101
        s.loops = 101;
        return true;
    }
    return false;
}

// Initial bracket (bucketMul.metal:155-197).  Whole CTA; every thread passes its PER values (< 0 = no value).
template <int PER>
__device__ __forceinline__ void bisect_init(const float (&vals)[PER], CutoffSmem& sm, BisectState& s) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    float tmin = 999.f, tmax = -999.f;  // :155-156
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
    __syncthreads();
    float gmin = (lane < nwarps) ? sm.red_min[lane] : 999.f;
    float gmax = (lane < nwarps) ? sm.red_max[lane] : -999.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
        gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    }
    // tgMin/tgMax are bfloat in the reference (:172-181): the 999 sentinel becomes 1000.
    s.minBound = bf16_round(gmin);
    s.maxBound = bf16_round(gmax);
    s.newBound = (s.minBound + s.maxBound) / 2;
    s.loops = 0; s.minCount = 4096; s.maxCount = 0;  // literals of the kernel (:197, :168-169)
}

// Phase A over a group of n_warps warps (gwarp = this thread's warp index inside the group; the group's values
// together are all the products).  Runs until the loop exits (returns true), fewer than kCutoffInsideMax products
// remain inside the bracket, or `max_levels` reference iterations have been performed by this call.
// The state is replicated in every thread of the group and stays identical by construction.
template <int PER, int n_warps, int bar_id>
__device__ __forceinline__ bool bisect_rounds(const float (&vals)[PER], BisectState& s, int effort, CutoffSmem& sm,
                                              int gwarp, int max_levels) {
    const int lane = threadIdx.x & 31;
    bool done = false;
    int buf = 0, levels = 0;
    while (!done && (s.minCount - s.maxCount) > kCutoffInsideMax && levels < max_levels) {
        const float t0 = s.newBound;
        const float tL = (t0 + s.minBound) / 2;   // next midpoint if count(t0) <  effort (maxBound := t0)
        const float tR = (s.maxBound + t0) / 2;   // next midpoint if count(t0) >= effort (minBound := t0)
        int c0 = 0, cLR = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {  // negatives never count
            c0 += (vals[k] > t0) ? 1 : 0;
            cLR += ((vals[k] > tL) ? 1 : 0) + ((vals[k] > tR) ? 0x10000 : 0);
        }
        c0 = warp_sum_i(c0);
        cLR = warp_sum_i(cLR);
        if (lane == 0) { sm.cnt2[buf][0][gwarp] = c0; sm.cnt2[buf][1][gwarp] = cLR; }
        group_bar(bar_id, n_warps * 32);
        int n0 = 0, nLR = 0;
        static_assert(n_warps % 4 == 0, "group size must be a multiple of 4 warps");
#pragma unroll
        for (int w4 = 0; w4 < n_warps / 4; w4++) {  // broadcast 16-byte reads
            const int4 a = *reinterpret_cast<const int4*>(&sm.cnt2[buf][0][w4 * 4]);
            const int4 b = *reinterpret_cast<const int4*>(&sm.cnt2[buf][1][w4 * 4]);
            n0 += a.x + a.y + a.z + a.w;
            nLR += b.x + b.y + b.z + b.w;
        }
        buf ^= 1;
        s.loops++;
        levels++;
        const bool went_left = n0 < effort;
        done = bisect_step(s, n0, effort);
        if (!done && (s.minCount - s.maxCount) > kCutoffInsideMax && levels < max_levels) {
            // s.newBound is now exactly tL or tR (same expression, same operands)
            const int n1 = went_left ? (nLR & 0xFFFF) : (nLR >> 16);
            s.loops++;
            levels++;
            done = bisect_step(s, n1, effort);
        }
    }
    return done;
}

// Phase B: the group compacts the products inside (minBound, maxBound] (<= kCutoffInsideMax by the phase-A exit
// condition) and its first warp finishes the bisection alone.  Returns the cutoff in every thread of the group
// (and leaves it in sm.result / sm.loops).  Must be called by all threads of the group; `done` = loop already over.
template <int PER, int n_warps, int bar_id>
__device__ __forceinline__ float bisect_finish(const float (&vals)[PER], BisectState& s, int effort, CutoffSmem& sm,
                                               int gwarp, bool done) {
    const int lane = threadIdx.x & 31;
    if (!done) {
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (vals[k] > s.minBound && vals[k] <= s.maxBound) {
                const int p = atomicAdd(&sm.n_inside, 1);
                if (p < kCutoffInsideMax) sm.inside[p] = vals[k];
            }
        }
        group_bar(bar_id, n_warps * 32);
        if (gwarp == 0) {
            const int n_in = min(sm.n_inside, kCutoffInsideMax);
            float x[kCutoffInsideMax / 32];
#pragma unroll
            for (int k = 0; k < kCutoffInsideMax / 32; k++) {
                const int i = lane + 32 * k;
                x[k] = (i < n_in) ? sm.inside[i] : -1.f;
            }
            // products above maxBound keep counting; maxCount is their exact number (0 if never set)
            const int above_max = s.maxCount;
            int prev_inside = -1, same_streak = 0;
            while (!done) {
                s.loops++;
                int c = above_max;
#pragma unroll
                for (int k = 0; k < kCutoffInsideMax / 32; k++)
                    c += __popc(__ballot_sync(0xffffffffu, x[k] > s.newBound));
                done = bisect_step(s, c, effort);
                if (done) break;
                // products still inside the bracket (an upper bound while minCount is the 4096 literal): when that
                // number stops shrinking the bracket most likely holds one repeated value; verify, then finish
                // with the scalar tail
                const int inside_est = s.minCount - s.maxCount;
                same_streak = (inside_est == prev_inside) ? same_streak + 1 : 0;
                prev_inside = inside_est;
                if (same_streak >= 1) {
                    int inside = 0;
                    unsigned lo = 0xFFFFFFFFu, hi = 0u;
#pragma unroll
                    for (int k = 0; k < kCutoffInsideMax / 32; k++) {
                        const bool in = x[k] > s.minBound && x[k] <= s.maxBound;
                        inside += __popc(__ballot_sync(0xffffffffu, in));
                        if (in) {  // non-negative floats order like their bits
                            lo = min(lo, __float_as_uint(x[k]));
                            hi = max(hi, __float_as_uint(x[k]));
                        }
                    }
                    lo = __reduce_min_sync(0xffffffffu, lo);
                    hi = __reduce_max_sync(0xffffffffu, hi);
                    if (inside > 0 && lo == hi) {  // one repeated value xv: count(b) = cnt_hi if b < xv else cnt_lo
                        const float xv = __uint_as_float(lo);
                        int cnt_lo = above_max;
#pragma unroll
                        for (int k = 0; k < kCutoffInsideMax / 32; k++)
                            cnt_lo += __popc(__ballot_sync(0xffffffffu, x[k] > s.maxBound));
                        const int cnt_hi = cnt_lo + inside;
                        while (!done) {
                            s.loops++;
                            done = bisect_step(s, (s.newBound < xv) ? cnt_hi : cnt_lo, effort);
                        }
                    }
                    same_streak = 0;
                }
            }
            if (lane == 0) { sm.result = s.newBound; sm.loops = s.loops; }
        }
        group_bar(bar_id, n_warps * 32);
        s.newBound = sm.result;
        s.loops = sm.loops;
    } else if (gwarp == 0 && lane == 0) {
        sm.result = s.newBound;
        sm.loops = s.loops;
    }
    return s.newBound;
}

// ---- direct method -----------------------------------------------------------------------------------
// The bisection never looks at the products themselves, only at count(b) = #{products > b}, and then only
//   (1) count(b) < k          <=>  b >= T_k                      (T_j = j-th largest product, k = 4096 - q)
//   (2) count(b) == k         <=>  T_{k+1} <= b < T_k
//   (3) |maxCount - minCount| < 3, which can only hold while both counts are within 3 of k.
// All three are decided exactly by the capped count  c'(b) = (k-3) + sum_{j=k-2..k+2} [T_j > b]
// (= clamp(count(b), k-3, k+2)), with the 4096 / 0 start literals capped the same way.  So the whole loop is
// a scalar recurrence over five order statistics: two histogram passes over the 4096 products (2048 bins of 16
// bf16 codes, then the 16 codes of each target bin) find T_{k-2..k+2} exactly -- ties of any multiplicity
// included -- and ONE warp replays the reference iterations (same bisect_step, same loop count) with five
// compares per iteration instead of a block-wide count.  Bit-identical to the iterative path (tests).
template <int PER>
__device__ __forceinline__ float block_cutoff_direct(const float (&vals)[PER], int n_probes, int q, CutoffSmem& sm,
                                                     int* loops_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, nwarps = NT >> 5;
    const int B = 2048 / NT;  // bins per thread in the scan (NT = 512 or 1024)
    const int k = n_probes - q;
    // S0: zero the histograms; min/max partials (bucketMul.metal:155-181)
    for (int b = tid; b < 2048; b += NT) sm.hist[b] = 0u;
    if (tid < 80) (&sm.sub[0][0])[tid] = 0;
    if (tid < 5) sm.tbin[tid] = -1;  // a rank beyond the number of countable (non-NaN) products has no bin
    float tmin = 999.f, tmax = -999.f;
#pragma unroll
    for (int j = 0; j < PER; j++)
        if (vals[j] >= 0.f) { tmin = fminf(tmin, vals[j]); tmax = fmaxf(tmax, vals[j]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    }
    if (lane == 0) { sm.red_min[warp] = tmin; sm.red_max[warp] = tmax; }
    __syncthreads();
    // S1: coarse histogram
#pragma unroll
    for (int j = 0; j < PER; j++)
        if (vals[j] >= 0.f) atomicAdd(&sm.hist[__float_as_uint(vals[j]) >> 20], 1u);  // (bits>>16)>>4
    __syncthreads();
    // S2: prefix over bins in DESCENDING order; locate the bins of ranks k-2..k+2
    {
        unsigned c[4];
        unsigned ts = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            c[j] = (j < B) ? sm.hist[2047 - (tid * B + j)] : 0u;
            ts += c[j];
        }
        unsigned incl = ts;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sm.wsum[warp] = (int)incl;
        __syncthreads();
        unsigned before = incl - ts;
        for (int w = 0; w < warp; w++) before += (unsigned)sm.wsum[w];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (j < B && c[j]) {
#pragma unroll
                for (int ri = 0; ri < 5; ri++) {
                    const int r = k - 2 + ri;
                    if (r >= 1 && (unsigned)r > before && (unsigned)r <= before + c[j]) {
                        sm.tbin[ri] = 2047 - (tid * B + j);
                        sm.trank[ri] = r - (int)before;
                    }
                }
            }
            before += c[j];
        }
    }
    __syncthreads();
    // S3: fine histogram of each target bin (ranks outside 1..n_probes have no bin: tbin stays unmatched)
    {
        int tb[5];
#pragma unroll
        for (int ri = 0; ri < 5; ri++) {
            const int r = k - 2 + ri;
            tb[ri] = (r >= 1 && r <= n_probes) ? sm.tbin[ri] : -1;
        }
#pragma unroll
        for (int j = 0; j < PER; j++) {
            if (vals[j] >= 0.f) {
                const unsigned key = __float_as_uint(vals[j]) >> 16;
#pragma unroll
                for (int ri = 0; ri < 5; ri++)
                    if ((int)(key >> 4) == tb[ri]) atomicAdd(&sm.sub[ri][key & 15u], 1);
            }
        }
    }
    __syncthreads();
    // S4: one warp resolves T_{k-2..k+2} and replays the reference loop
    if (warp == 0) {
        float T[5];
#pragma unroll
        for (int ri = 0; ri < 5; ri++) {
            const int r = k - 2 + ri;
            float t;
            if (r < 1) t = __int_as_float(0x7F800000);   // count(b) >= r always
            else if (r > n_probes || sm.tbin[ri] < 0) t = -1.f;  // never
            else {
                int need = sm.trank[ri], code = 15, acc = 0;
                for (; code > 0; code--) {
                    acc += sm.sub[ri][code];
                    if (acc >= need) break;
                }
                t = __uint_as_float((((unsigned)sm.tbin[ri] << 4) | (unsigned)code) << 16);
            }
            T[ri] = t;
        }
        float gmin = (lane < nwarps) ? sm.red_min[lane] : 999.f;
        float gmax = (lane < nwarps) ? sm.red_max[lane] : -999.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
            gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        }
        BisectState s;
        s.minBound = bf16_round(gmin);  // 999 -> 1000 through the bfloat tgMin (:172-181)
        s.maxBound = bf16_round(gmax);
        s.newBound = (s.minBound + s.maxBound) / 2;
        s.loops = 0;
        s.minCount = min(4096, k + 2);  // capped start literals (:197, :168-169)
        s.maxCount = max(0, k - 3);
        bool done = false;
        while (!done) {
            s.loops++;
            const float b = s.newBound;
            const int c = (k - 3) + (T[0] > b ? 1 : 0) + (T[1] > b ? 1 : 0) + (T[2] > b ? 1 : 0) +
                          (T[3] > b ? 1 : 0) + (T[4] > b ? 1 : 0);
            done = bisect_step(s, c, k);
        }
        if (lane == 0) { sm.result = s.newBound; sm.loops = s.loops; }
    }
    __syncthreads();
    if (loops_out && tid == 0) *loops_out = sm.loops;
    return sm.result;
}

// Whole-CTA bisection (stand-alone kernel and the non-overlapped fused path).
template <int NWB, int PER>
__device__ __forceinline__ float block_bisect_cutoff(const float (&vals)[PER], int n_probes, int q,
                                                     CutoffSmem& sm, int* loops_out,
                                                     unsigned long long* trace = nullptr) {
    (void)trace;
#ifdef EFFORT_CUTOFF_DIRECT  // measured slower than the iterative path on B200 (shared-memory atomics + a ~30-step
                             // scalar replay cost ~6 us vs ~5.5 us); kept for study, exactness covered by the same tests
    if ((2048 % blockDim.x) == 0 && blockDim.x >= 512) return block_cutoff_direct(vals, n_probes, q, sm, loops_out);
#endif
    BisectState s;
    bisect_init(vals, sm, s);
    const int effort = n_probes - q;  // :154
    const int w = threadIdx.x >> 5;
    const bool done = bisect_rounds<PER, NWB, 0>(vals, s, effort, sm, w, 1 << 30);
    const float c = bisect_finish<PER, NWB, 0>(vals, s, effort, sm, w, done);
    if (loops_out && threadIdx.x == 0) *loops_out = s.loops;
    return c;
}

// Scores this thread's probes: thread t owns probes t, t+NT, ...  (bucketMul.metal:158-163)
template <bool NORM = false, int PER>
__device__ __forceinline__ void score_probes(const float* __restrict__ v, const __half* __restrict__ probes,
                                             uint32_t exp_no, int n_probes, float (&vals)[PER],
                                             uint64_t keep, const __half* __restrict__ norm_w = nullptr,
                                             float denom = 1.f) {
#pragma unroll
    for (int k = 0; k < PER; k++) {
        int i = threadIdx.x + k * blockDim.x;
        if (i < n_probes) {
            const uint16_t pb16 = ldg_keep_u16(probes + (size_t)exp_no * n_probes + i, keep);
            float p = bf16_round(__half2float(__ushort_as_half(pb16)));
            float vi = v[i];
            if constexpr (NORM) vi = (vi / denom) * __half2float(norm_w[i]);  // fused rmsNorm*w on load
            float x = __fmul_rn(__fmul_rn(kCutoffScale, vi), p);
            vals[k] = bf16_round(fabsf(x));
        } else {
            vals[k] = -1.f;
        }
    }
}

// Same in two halves so that the (constant) probes can be fetched before a PDL wait and the activations after.
template <int PER>
__device__ __forceinline__ void load_probes(const __half* __restrict__ probes, uint32_t exp_no, int n_probes,
                                            uint16_t (&pr)[PER], uint64_t keep) {
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int i = threadIdx.x + k * blockDim.x;
        pr[k] = (i < n_probes) ? ldg_keep_u16(probes + (size_t)exp_no * n_probes + i, keep) : (uint16_t)0;
    }
}
template <bool NORM = false, int PER>
__device__ __forceinline__ void score_loaded(const float* __restrict__ v, const uint16_t (&pr)[PER], int n_probes,
                                             float (&vals)[PER], const __half* __restrict__ norm_w = nullptr,
                                             float denom = 1.f) {
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int i = threadIdx.x + k * blockDim.x;
        if (i < n_probes) {
            const float p = bf16_round(__half2float(__ushort_as_half(pr[k])));
            float vi = v[i];
            if constexpr (NORM) vi = (vi / denom) * __half2float(norm_w[i]);
            vals[k] = bf16_round(fabsf(__fmul_rn(__fmul_rn(kCutoffScale, vi), p)));
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
    score_probes(v, probes, exp_no, n_probes, vals, l2_policy_evict_last());
    float c = block_bisect_cutoff<kCutoffThreads / 32>(vals, n_probes, q, sm, loops_out);
    if (threadIdx.x == 0) *cutoff_out = c;
}

}  // namespace effort
