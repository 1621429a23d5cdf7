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
// executed in one of two ways:
//   group path (group_cutoff, further down): what the fused bucketMul kernel and the stand-alone launch run for
//            n_probes <= 4096 -- four warps, packed bf16 counting, rank-sort of the bracket, scalar replay;
//   block-wide path (block_bisect_cutoff): any n_probes <= 8192, all warps of the CTA:
//     phase A  counting rounds while more than 128 products lie inside (minBound, maxBound]; every round performs
//              TWO reference iterations: besides the count at the current midpoint it counts at both midpoints
//              the next iteration could pick ((mid+min)/2 and (max+mid)/2 -- the same fp32 expressions the
//              reference evaluates), so the second step needs no synchronisation;
//     phase B  the <=128 inside products are compacted to shared memory and ONE warp finishes with no block
//              barriers; as soon as everything left inside is one repeated value the count is a step function
//              of the midpoint and the remaining halvings down to two adjacent floats cost a compare each;
//   fixpoint (both) once the midpoint stops moving, later loops cannot change the state, so the remaining
//            iterations up to the reference's 100-loop cap are skipped (loops is reported as 101).
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
    float sorted[kCutoffInsideMax];   // group path: the inside products in descending order
    int n_inside;
    float result;
    int loops;
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

// The same iteration without early-return branches (one select per state variable, one predicate out).
__device__ __forceinline__ bool bisect_step_flat(BisectState& s, int countAbove, int effort) {
    const bool left = countAbove < effort;
    const float b = s.newBound;
    s.maxBound = left ? b : s.maxBound;
    s.maxCount = left ? countAbove : s.maxCount;
    s.minBound = left ? s.minBound : b;
    s.minCount = left ? s.minCount : countAbove;
    s.newBound = (s.maxBound + s.minBound) / 2;
    const bool e1 = (countAbove == effort) | (s.maxBound - s.minBound < 0.00001f) | (abs(s.maxCount - s.minCount) < 3);
    const bool e2 = s.loops > 100;
    const bool e3 = s.newBound == b;  // fixpoint: the reference would idle to loops == 101
    s.loops = (!e1 & !e2 & e3) ? 101 : s.loops;
    return e1 | e2 | e3;
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

// ---- group path (the one the fused kernel and the stand-alone launch use) --------------------------------
// The block-wide path above is latency-bound on a serial chain of predicated adds and on the work every thread
// replicates, so the cutoff runs on FOUR warps (the others are free), 32 products per thread:
//   * the products are bf16 values, so they are kept packed two per register and counted with HSET2.BF16 +
//     HADD2.BF16: for a bf16 x and an fp32 threshold t >= 0,  x > t  <=>  x > trunc_bf16(t)  -- exact;
//   * phase A: the same two-iterations-per-round scheme, 128-thread named barrier;
//   * the <= 128 products left inside the bracket are compacted as keys (value bits | slot: unique, ordered like
//     the values) and rank-sorted by counting (thread t ranks key t), which yields the five order statistics
//     T_{k-2..k+2} of the whole product set (ranks above the bracket are +inf, ranks below it can never exceed
//     a later midpoint);
//   * ONE thread replays the remaining reference iterations with the capped count
//         c'(b) = (k-3) + sum_j [T_j > b] = clamp(count(b), k-3, k+2)
//     (the loop only ever tests count < k, count == k and |maxCount - minCount| < 3; see "direct method" below
//     and tests/test_oracle.py for the proof).  The bracket update needs only [T_k > b], so the loop-carried
//     chain is compare -> select -> add -> mul.
constexpr int kCutWarps = 4;
constexpr int kCutThreads = kCutWarps * 32;
constexpr int kCutPer = 32;                          // products per group thread
constexpr int kCutPairs = kCutPer / 2;
constexpr int kCutGroupMax = kCutThreads * kCutPer;  // 4096 = probesCount (bucketMul.swift:19)

struct GroupProbes { uint4 w[kCutPer / 8]; };  // fp16 codes; element (c, m) <-> probe (c*128 + gt)*8 + m
struct GroupProducts {
    uint32_t pv[kCutPairs];  // bf16x2; 0xBF80 (-1) = no product
    float tmin, tmax;        // this thread's min / max over its products (:155-163)
};

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ unsigned long long cut_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// -1 if a > b else 0 (FSET): lets the five compares of the replay feed a two-level IADD3 tree
__device__ __forceinline__ int fgt_mask(float a, float b) {
    int d;
    asm("set.gt.s32.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ __nv_bfloat162 as_bf162(uint32_t u) { return *reinterpret_cast<__nv_bfloat162*>(&u); }
__device__ __forceinline__ uint32_t bf162_bits(__nv_bfloat162 h) { return *reinterpret_cast<uint32_t*>(&h); }
// fp32 threshold (>= 0) -> truncated bf16 in both halves
__device__ __forceinline__ __nv_bfloat162 thr2(float t) {
    const uint32_t hi = __float_as_uint(t) & 0xFFFF0000u;
    return as_bf162(hi | (hi >> 16));
}
// sum of the two bf16 halves (small exact integers) as int
__device__ __forceinline__ int bf162_count(__nv_bfloat162 a) {
    const uint32_t u = bf162_bits(a);
    return (int)(__uint_as_float(u << 16) + __uint_as_float(u & 0xFFFF0000u));
}

// probes already points at the expert's n_probes codes.  Constant data: may be issued before a PDL wait.
__device__ __forceinline__ void group_load_probes(const __half* __restrict__ probes, int n_probes, int gt,
                                                  GroupProbes& pr, uint64_t keep) {
    const bool vec = ((n_probes & 7) == 0) && aligned16(probes);
#pragma unroll
    for (int c = 0; c < kCutPer / 8; c++) {
        const int i0 = (c * kCutThreads + gt) * 8;
        if (vec && i0 + 8 <= n_probes) {
            pr.w[c] = ldg_keep_u4(reinterpret_cast<const uint4*>(probes + i0), keep);
        } else {
            uint32_t h[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int m = 0; m < 8; m++)
                if (i0 + m < n_probes) h[m >> 1] |= (uint32_t)ldg_keep_u16(probes + i0 + m, keep) << (16 * (m & 1));
            pr.w[c] = make_uint4(h[0], h[1], h[2], h[3]);
        }
    }
}

// products of group thread gt: bfloat(|1e5 * v[i] * bfloat(probes[i])|)  (bucketMul.metal:158-163)
template <bool NORM>
__device__ __forceinline__ void group_score(const float* __restrict__ v, const GroupProbes& pr, int n_probes, int gt,
                                            GroupProducts& gp, const __half* __restrict__ norm_w, float denom) {
    const bool vec = ((n_probes & 7) == 0) && aligned16(v) && (!NORM || aligned16(norm_w));
    // cvt.rn.bf16x2.f32 is the same round-to-nearest-even as bf16_round (NaN payloads aside: a NaN never counts)
    __nv_bfloat162 mn2 = as_bf162(0x7F807F80u), mx2 = as_bf162(0xBF80BF80u);  // +inf / -1: neutral for min / max
#pragma unroll
    for (int c = 0; c < kCutPer / 8; c++) {
        const int i0 = (c * kCutThreads + gt) * 8;
        float vv[8];
        uint32_t nw[4] = {0u, 0u, 0u, 0u};
        if (vec && i0 + 8 <= n_probes) {
            const float4 a = *reinterpret_cast<const float4*>(v + i0), b = *reinterpret_cast<const float4*>(v + i0 + 4);
            vv[0] = a.x; vv[1] = a.y; vv[2] = a.z; vv[3] = a.w; vv[4] = b.x; vv[5] = b.y; vv[6] = b.z; vv[7] = b.w;
            if constexpr (NORM) {
                const uint4 t = *reinterpret_cast<const uint4*>(norm_w + i0);
                nw[0] = t.x; nw[1] = t.y; nw[2] = t.z; nw[3] = t.w;
            }
        } else {
#pragma unroll
            for (int m = 0; m < 8; m++) {
                vv[m] = (i0 + m < n_probes) ? v[i0 + m] : 0.f;
                if constexpr (NORM)
                    if (i0 + m < n_probes)
                        nw[m >> 1] |= (uint32_t)__half_as_ushort(norm_w[i0 + m]) << (16 * (m & 1));
            }
        }
        const uint32_t pw[4] = {pr.w[c].x, pr.w[c].y, pr.w[c].z, pr.w[c].w};
#pragma unroll
        for (int m2 = 0; m2 < 4; m2++) {
            const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&pw[m2]));
            const uint32_t pb = bf162_bits(__floats2bfloat162_rn(pf.x, pf.y));  // bfloat(probe)
            float v0 = vv[2 * m2], v1 = vv[2 * m2 + 1];
            if constexpr (NORM) {
                const float2 wf = __half22float2(*reinterpret_cast<const __half2*>(&nw[m2]));
                v0 = (v0 / denom) * wf.x;
                v1 = (v1 / denom) * wf.y;
            }
            const float x0 = fabsf(__fmul_rn(__fmul_rn(kCutoffScale, v0), __uint_as_float(pb << 16)));
            const float x1 = fabsf(__fmul_rn(__fmul_rn(kCutoffScale, v1), __uint_as_float(pb & 0xFFFF0000u)));
            uint32_t xb = bf162_bits(__floats2bfloat162_rn(x0, x1));
            if (i0 + 2 * m2 >= n_probes) xb = (xb & 0xFFFF0000u) | 0xBF80u;       // no product: -1
            if (i0 + 2 * m2 + 1 >= n_probes) xb = (xb & 0x0000FFFFu) | 0xBF800000u;
            gp.pv[c * 4 + m2] = xb;
            mx2 = __hmax2(mx2, as_bf162(xb));  // maxNum semantics: NaN is ignored, like the reference's compares
            uint32_t xm = xb;                  // for the minimum "no product" must be +inf
            if (i0 + 2 * m2 >= n_probes) xm = (xm & 0xFFFF0000u) | 0x7F80u;
            if (i0 + 2 * m2 + 1 >= n_probes) xm = (xm & 0x0000FFFFu) | 0x7F800000u;
            mn2 = __hmin2(mn2, as_bf162(xm));
        }
    }
    const uint32_t mnb = bf162_bits(mn2), mxb = bf162_bits(mx2);
    gp.tmin = fminf(999.f, fminf(__uint_as_float(mnb << 16), __uint_as_float(mnb & 0xFFFF0000u)));   // :155-163
    gp.tmax = fmaxf(-999.f, fmaxf(__uint_as_float(mxb << 16), __uint_as_float(mxb & 0xFFFF0000u)));
}

// Called by the kCutThreads threads of the group (gt = index inside it).  Leaves the cutoff in sm.result and the
// reference's loop count in sm.loops; the caller publishes them with a barrier that includes the group.
template <int bar_id>
__device__ __forceinline__ void group_cutoff(const GroupProducts& gp, int n_probes, int q, CutoffSmem& sm, int gt,
                                             unsigned long long* trace = nullptr) {
    const int lane = gt & 31, gwarp = gt >> 5;
    const int k = n_probes - q;  // bucketMul.metal:154
    uint32_t* keys = reinterpret_cast<uint32_t*>(sm.inside);
    uint32_t* sorted = reinterpret_cast<uint32_t*>(sm.sorted);
    float tmin = gp.tmin, tmax = gp.tmax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    }
    if (lane == 0) { sm.red_min[gwarp] = tmin; sm.red_max[gwarp] = tmax; }
    keys[gt] = 0u;  // padding: below every key that is ranked
    if (gt == 0) sm.n_inside = 0;
    group_bar(bar_id, kCutThreads);
    float gmin = (lane < kCutWarps) ? sm.red_min[lane] : 999.f;
    float gmax = (lane < kCutWarps) ? sm.red_max[lane] : -999.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
        gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    }
    BisectState s;
    s.minBound = bf16_round(gmin);  // tgMin/tgMax are bfloat (:172-181): the 999 sentinel becomes 1000
    s.maxBound = bf16_round(gmax);
    s.newBound = (s.minBound + s.maxBound) / 2;
    s.loops = 0; s.minCount = 4096; s.maxCount = 0;  // literals of the kernel (:197, :168-169)
    if (trace && gt == 0) trace[4] = cut_gtime();

    // phase A: two reference iterations per round
    bool done = false;
    int buf = 0;
    while (!done && (s.minCount - s.maxCount) > kCutoffInsideMax) {
        const float t0 = s.newBound;
        const float tL = (t0 + s.minBound) / 2;   // next midpoint if count(t0) <  k (maxBound := t0)
        const float tR = (s.maxBound + t0) / 2;   // next midpoint if count(t0) >= k (minBound := t0)
        const __nv_bfloat162 b0 = thr2(t0), bL = thr2(tL), bR = thr2(tR);
        __nv_bfloat162 a0 = as_bf162(0u), aL = as_bf162(0u), aR = as_bf162(0u);
#pragma unroll
        for (int j = 0; j < kCutPairs; j++) {
            const __nv_bfloat162 x = as_bf162(gp.pv[j]);
            a0 = __hadd2(a0, __hgt2(x, b0));
            aL = __hadd2(aL, __hgt2(x, bL));
            aR = __hadd2(aR, __hgt2(x, bR));
        }
        int c0 = bf162_count(a0);
        int cLR = bf162_count(aL) | (bf162_count(aR) << 16);
        c0 = warp_sum_i(c0);
        cLR = warp_sum_i(cLR);
        if (lane == 0) { sm.cnt2[buf][0][gwarp] = c0; sm.cnt2[buf][1][gwarp] = cLR; }
        group_bar(bar_id, kCutThreads);
        const int4 ca = *reinterpret_cast<const int4*>(&sm.cnt2[buf][0][0]);
        const int4 cb = *reinterpret_cast<const int4*>(&sm.cnt2[buf][1][0]);
        const int n0 = (ca.x + ca.y) + (ca.z + ca.w);
        const int nLR = (cb.x + cb.y) + (cb.z + cb.w);
        buf ^= 1;
        s.loops++;
        const bool went_left = n0 < k;
        done = bisect_step_flat(s, n0, k);
        if (!done && (s.minCount - s.maxCount) > kCutoffInsideMax) {
            s.loops++;  // s.newBound is now exactly tL or tR (same expression, same operands)
            done = bisect_step_flat(s, went_left ? (nLR & 0xFFFF) : (nLR >> 16), k);
        }
    }
    if (trace && gt == 0) { trace[5] = cut_gtime(); trace[11] = (unsigned long long)s.loops; }
    if (!done) {
        // <= kCutoffInsideMax products are left inside (minBound, maxBound]: compact their keys (one
        // shared-memory atomic per warp: thread counts, warp scan, warp base)
        const __nv_bfloat162 bm = thr2(s.minBound), bM = thr2(s.maxBound);
        uint32_t in_mask[kCutPairs];
        int mine = 0;
#pragma unroll
        for (int j = 0; j < kCutPairs; j++) {
            const __nv_bfloat162 x = as_bf162(gp.pv[j]);
            in_mask[j] = __hgt2_mask(x, bm) & ~__hgt2_mask(x, bM);  // x > minBound && !(x > maxBound)
            mine += (int)(in_mask[j] & 1u) + (int)(in_mask[j] >> 31);
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int wbase = 0;
        if (lane == 31 && incl) wbase = atomicAdd(&sm.n_inside, incl);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        {
            int pos = wbase + incl - mine;
#pragma unroll
            for (int j = 0; j < kCutPairs; j++) {  // predicated stores, no branches
                const int lo = (int)(in_mask[j] & 1u), hi = (int)(in_mask[j] >> 31);
                if (lo & (pos < kCutoffInsideMax)) keys[pos] = (gp.pv[j] << 16) | (uint32_t)pos;
                pos += lo;
                if (hi & (pos < kCutoffInsideMax)) keys[pos] = (gp.pv[j] & 0xFFFF0000u) | (uint32_t)pos;
                pos += hi;
            }
        }
        group_bar(bar_id, kCutThreads);
        if (trace && gt == 0) trace[14] = cut_gtime();
        const int n_in = min(sm.n_inside, kCutoffInsideMax);
        const uint32_t x = keys[gt];
        int r0 = 0, r1 = 0, r2 = 0, r3 = 0;  // descending rank of key gt among the n_in keys
        const int n4 = (n_in + 3) >> 2;
#pragma unroll 8
        for (int j4 = 0; j4 < n4; j4++) {
            const uint4 y = *reinterpret_cast<const uint4*>(&keys[j4 * 4]);
            r0 += (y.x > x) ? 1 : 0;
            r1 += (y.y > x) ? 1 : 0;
            r2 += (y.z > x) ? 1 : 0;
            r3 += (y.w > x) ? 1 : 0;
        }
        if (gt < n_in) sorted[(r0 + r1) + (r2 + r3)] = x;  // padding (0) is never above a key and is not ranked
        group_bar(bar_id, kCutThreads);
        if (trace && gt == 0) trace[13] = cut_gtime();
        if (gt == 0) {
            const int above = s.maxCount;  // exact number of products > maxBound (0 while maxBound is the maximum)
            float T[5];
#pragma unroll
            for (int ri = 0; ri < 5; ri++) {
                const int idx = (k - 2 + ri) - above - 1;  // 0-based rank inside the bracket
                T[ri] = (idx < 0) ? __int_as_float(0x7F800000)
                                  : (idx >= n_in ? -1.f : __uint_as_float(sorted[idx] & 0xFFFF0000u));
            }
            // scalar replay in the capped domain; same statements as bisect_step, arranged so that the
            // loop-carried chain is  [T_k > b] -> select -> add -> mul
            float minB = s.minBound, maxB = s.maxBound, nb = s.newBound;
            int minC = min(s.minCount, k + 2), maxC = max(s.maxCount, k - 3), loops = s.loops, iters = 0;
            // four iterations per trip: the exit tests (a long predicate chain) of one iteration overlap the
            // bracket updates of the next ones; the first iteration that exits provides the result
            struct Rp { float minB, maxB, nb; int minC, maxC; };
            Rp r = {minB, maxB, nb, minC, maxC};
            auto step = [&](Rp& q, int loops_now, bool& stop, bool& fix) {
                const float bnd = q.nb;
                const bool left = !(T[2] > bnd);  // count(b) < k
                const int c = (k - 3) - ((fgt_mask(T[0], bnd) + fgt_mask(T[1], bnd) + fgt_mask(T[2], bnd)) +
                                         (fgt_mask(T[3], bnd) + fgt_mask(T[4], bnd)));
                q.maxB = left ? bnd : q.maxB;
                q.minB = left ? q.minB : bnd;
                q.maxC = left ? c : q.maxC;
                q.minC = left ? q.minC : c;
                q.nb = (q.maxB + q.minB) / 2;
                const bool e1 = (c == k) | (q.maxB - q.minB < 0.00001f) | (abs(q.maxC - q.minC) < 3);
                const bool e2 = loops_now > 100;
                const bool e3 = q.nb == bnd;  // fixpoint: the reference would idle to its loop cap
                fix = !e1 & !e2 & e3;
                stop = e1 | e2 | e3;
            };
            while (true) {
                bool s1, s2, s3, s4, f1, f2, f3, f4;
                Rp r1 = r;  step(r1, loops + 1, s1, f1);
                Rp r2 = r1; step(r2, loops + 2, s2, f2);
                Rp r3 = r2; step(r3, loops + 3, s3, f3);
                Rp r4 = r3; step(r4, loops + 4, s4, f4);
                if (s1 | s2 | s3 | s4) {
                    const int n = s1 ? 1 : (s2 ? 2 : (s3 ? 3 : 4));
                    const bool fx = s1 ? f1 : (s2 ? f2 : (s3 ? f3 : f4));
                    nb = s1 ? r1.nb : (s2 ? r2.nb : (s3 ? r3.nb : r4.nb));
                    iters += n;
                    loops = fx ? 101 : loops + n;
                    break;
                }
                r = r4;
                loops += 4;
                iters += 4;
            }
            sm.result = nb;
            sm.loops = loops;
            if (trace) { trace[12] = (unsigned long long)loops; trace[15] = (unsigned long long)iters; }
        }
    } else if (gt == 0) {
        sm.result = s.newBound;
        sm.loops = s.loops;
    }
}

// ---- why five order statistics are enough -----------------------------------------------------------------
// The bisection never looks at the products themselves, only at count(b) = #{products > b}, and then only
//   (1) count(b) < k          <=>  b >= T_k                      (T_j = j-th largest product, k = 4096 - q)
//   (2) count(b) == k         <=>  T_{k+1} <= b < T_k
//   (3) |maxCount - minCount| < 3, which can only hold while both counts are within 3 of k.
// All three are decided exactly by the capped count  c'(b) = (k-3) + sum_{j=k-2..k+2} [T_j > b]
// (= clamp(count(b), k-3, k+2)), with the 4096 / 0 start literals capped the same way: this is what group_cutoff's
// scalar replay evaluates (tests/test_oracle.py checks it against the literal loop, from the start and from a
// mid-way state).  A variant that found the T_j with two shared-memory histogram passes instead of phase A was
// measured slower (atomic contention on the few occupied bins) and removed; it is in the history (commit 7979362).

// Whole-CTA bisection (stand-alone kernel when n_probes > 4096).
template <int NWB, int PER>
__device__ __forceinline__ float block_bisect_cutoff(const float (&vals)[PER], int n_probes, int q,
                                                     CutoffSmem& sm, int* loops_out,
                                                     unsigned long long* trace = nullptr) {
    (void)trace;
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

// Stand-alone launch (test hook effort_find_cutoff, and the first stage of the unfused path).
__global__ void __launch_bounds__(kCutoffThreads, 1)
find_cutoff_kernel(const float* __restrict__ v, const __half* __restrict__ probes,
                   const uint32_t* __restrict__ exp_no_dev, int n_probes, int q,
                   float* __restrict__ cutoff_out, int* __restrict__ loops_out) {
    __shared__ CutoffSmem sm;
    const uint32_t exp_no = exp_no_dev ? *exp_no_dev : 0u;
    if (n_probes <= kCutGroupMax) {  // same code path as the fused bucketMul kernel
        if (threadIdx.x < kCutThreads) {
            GroupProbes pr;
            GroupProducts gp;
            group_load_probes(probes + (size_t)exp_no * n_probes, n_probes, threadIdx.x, pr, l2_policy_evict_last());
            group_score<false>(v, pr, n_probes, threadIdx.x, gp, nullptr, 1.f);
            group_cutoff<1>(gp, n_probes, q, sm, threadIdx.x);
        }
        __syncthreads();
        if (threadIdx.x == 0) { *cutoff_out = sm.result; if (loops_out) *loops_out = sm.loops; }
        return;
    }
    float vals[kCutoffMaxPerThread];
    score_probes(v, probes, exp_no, n_probes, vals, l2_policy_evict_last());
    float c = block_bisect_cutoff<kCutoffThreads / 32>(vals, n_probes, q, sm, loops_out);
    if (threadIdx.x == 0) *cutoff_out = c;
}

}  // namespace effort
