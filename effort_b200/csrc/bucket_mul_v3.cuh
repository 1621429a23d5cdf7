// bucket_mul_v3.cuh -- the TMA pipeline variant of the fused bucketMul (FP16 buckets, slice-major device layout).
//
// Same operator and same prologue stages as bucket_mul_v2_kernel (cutoff -> selection -> gather-MAC -> reductions into
// `out`; reference: BucketMul.fullMul, bucketMul.swift:34-88 and bucketMul.metal:11-247), different machinery:
//
//  * one PRODUCER warp (warp 16) owns all HBM traffic.  In the slice-major layout the rank rows an input selects inside
//    this CTA's column slice are one contiguous byte range (n rows x 256 B), so a streaming unit = one input's run of
//    selected ranks = ONE cp.async.bulk (TMA unit) copy into a byte ring in shared memory, completion on an mbarrier
//    (expect_tx / complete_tx).  The producer walks the per-input selection masks 32 inputs at a time (the lanes issue
//    their copies in parallel), allocates ring space first-in-first-out and reclaims it as consumers release units
//    (one "empty" mbarrier per descriptor slot).  No selection list is built and the 16 consumer warps spend no
//    instruction on addresses or copies.
//  * 16 CONSUMER warps take units from a shared ticket counter (units differ in size: 1..16 rows), wait on the unit's
//    "full" barrier, run the read-modify-write accumulate over its rows four at a time (the rows of a unit belong to
//    one input: their updates never alias, see accumulate_unit_fp16) and release the unit.
//  * prologue: the exact-select cutoff exchanges its per-warp counts through double-buffered shared slots (no
//    shared-memory atomics on the critical path) and the 128 KB of accumulator tiles are zeroed inside its rounds (the
//    rounds are latency bound, the stores are free there); the overwrite protocol (zero + fence + arrive) is run by
//    the producer warp, which is idle until the masks exist.
#pragma once
#include "bucket_mul_v2.cuh"

namespace effort {

constexpr int kV3Threads = kV2Threads + 32;  // 16 consumer warps + 1 producer warp
constexpr int kV3Desc = 64;                  // descriptor slots (units in flight), a power of two
constexpr int kV3BatchBytes = 20 * 1024;     // the producer allocates at most this much ring space at a time

struct __align__(16) V3Desc {
    uint32_t off;    // byte offset of the unit's first row in the ring
    uint32_t n;      // rows (0 = poison: no more units)
    float val;       // the input's multiplier
    uint32_t rstride;  // bytes between consecutive rows of the unit in the ring (= bytes of a row slice)
};

struct V3Header {
    CutoffSmem cut;                      // bisect mode scratch
    uint2 sel_slot[2][kV2Warps + 1];     // select mode: per-warp packed counts, double buffered by round parity
    float red[kV2Warps + 1];
    int sel_rows;                        // rows selected by this CTA (statistics)
    unsigned ticket;                     // next unit sequence number a consumer may take
    unsigned long long full_bar[kV3Desc];
    unsigned long long empty_bar[kV3Desc];
    uint32_t usize[kV3Desc];             // ring bytes a unit holds (incl. a wrap skip charged to it)
    V3Desc desc[kV3Desc];
};

struct V3Smem {
    static constexpr int kTileFloats = 16 * 32 * 4;
    static constexpr int kTileBytes = kTileFloats * 4;
    static constexpr size_t kHdrBytes = (sizeof(V3Header) + 127) & ~size_t(127);
    static constexpr size_t kFixed = (size_t)kTileBytes /*alignment slack*/ + (size_t)kV2Warps * kTileBytes + kHdrBytes +
                                     (size_t)kV2MaxInputs * (4 + 4 + 4) + 128;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// exact select (see select_cutoff in bucket_mul_v2.cuh) with slot exchange; `zero_fn(round)` is called once per round
// between the counting and the barrier: independent work that hides in the round's latency.
template <typename ZeroFn>
__device__ __forceinline__ float select_cutoff_slots(const uint32_t (&keys)[4], int k, V3Header& hdr, int tid, int n_warps,
                                                     ZeroFn zero_fn) {
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t u = 0;
    bool any = true;
    const unsigned need = (unsigned)(k + 1);
#pragma unroll 1
    for (int round = 0; round < 8; round++) {
        const int b = 14 - 2 * round;
        uint32_t th1, th2, th3;
        if (round == 0) { th1 = 0u; th2 = 1u << 14; th3 = 0x7FFFu; }
        else { th1 = u | (1u << b); th2 = u | (2u << b); th3 = u | (3u << b); }
        const uint32_t c1 = count_gt2(keys, th1), c2 = count_gt2(keys, th2), c3 = count_gt2(keys, th3);
        const uint32_t a = __reduce_add_sync(0xffffffffu, c1 | (c2 << 16));
        const uint32_t bsum = __reduce_add_sync(0xffffffffu, c3);
        if (lane == 0) hdr.sel_slot[round & 1][warp] = make_uint2(a, bsum);
        zero_fn(round);
        __syncthreads();
        const uint2 s = (lane < n_warps) ? hdr.sel_slot[round & 1][lane] : make_uint2(0u, 0u);
        const uint32_t A = __reduce_add_sync(0xffffffffu, s.x), B = __reduce_add_sync(0xffffffffu, s.y);
        const unsigned g1 = A & 0xFFFFu, g2 = A >> 16, g3 = B;
        if (round == 0) {
            any = g1 >= need;
            if (g2 >= need) u = 1u << 14;
        } else {
            const unsigned j = (g1 >= need ? 1u : 0u) + (g2 >= need ? 1u : 0u) + (g3 >= need ? 1u : 0u);
            u |= j << b;
        }
    }
    const uint32_t t = any ? (u + 1u) : 0u;
    return __uint_as_float(t << 16);
}

template <int CUT>
__global__ void __launch_bounds__(kV3Threads, 1)
bucket_mul_v3_kernel(const __grid_constant__ V2Batch batch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int SLOTS = 16, VEC = 4;
    constexpr int NT = kV2Threads, NW = kV2Warps;
    constexpr int TF = V3Smem::kTileFloats, TW = 32 * VEC, LB = VEC * 2;
    constexpr int kRow = 32 * LB;  // 256: bytes of a full-width row slice

    int pi = 0;
#pragma unroll
    for (int k = 1; k < kMulBatchMax; k++) pi += (k < batch.n && (int)blockIdx.x >= batch.cta_begin[k]) ? 1 : 0;
    const V2Problem& pb = batch.p[pi];
    const int lb = (int)blockIdx.x - batch.cta_begin[pi];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == NW;
    const int slice = lb % pb.CS, rsp = lb / pb.CS;
    const int RS = pb.RS, P = pb.P, C = pb.C;

    // ---- carve shared memory ----
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t s1 = (s0 + (uint32_t)V3Smem::kTileBytes - 1u) & ~((uint32_t)V3Smem::kTileBytes - 1u);
    unsigned char* p = smem_raw + (s1 - s0);
    float* tiles = reinterpret_cast<float*>(p);
    const uint32_t tiles_saddr = s1;
    p += (size_t)NW * V3Smem::kTileBytes;
    V3Header& hdr = *reinterpret_cast<V3Header*>(p);
    p += V3Smem::kHdrBytes;
    uint32_t* sbase = reinterpret_cast<uint32_t*>(p);  // per local input: its rank-0 row slice, 16-byte units from bk16
    p += (size_t)kV2MaxInputs * 4;
    float* sval = reinterpret_cast<float*>(p);         // per local input: the multiplier v[i]
    p += (size_t)kV2MaxInputs * 4;
    uint32_t* smask = reinterpret_cast<uint32_t*>(p);  // per local input: selected ranks (bit rho)
    p += (size_t)kV2MaxInputs * 4;
    p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 127) & ~uintptr_t(127));
    const uint32_t ring_saddr = (uint32_t)__cvta_generic_to_shared(p);
    const uint32_t ring_bytes = (uint32_t)batch.ring_bytes;

    pdl_trigger();
    if (pb.exp_no) pdl_wait();
    const uint32_t e_no = pb.exp_no ? *pb.exp_no : 0u;
    V2_TRACE(0);

    // ---- 0. constant metadata before the dependency wait ----
    const uint64_t keep = l2_policy_evict_last();
    const int n_in = (pb.in > rsp) ? (pb.in - 1 - rsp) / RS + 1 : 0;
    float sel_stat[16];
#pragma unroll
    for (int rho = 0; rho < 16; rho++) sel_stat[rho] = 0.f;
    auto load_stats = [&](int j, float (&st)[16]) {
        const int i = rsp + j * RS;
        if (P == 16) {
            const uint4* sp = reinterpret_cast<const uint4*>(pb.st16 + ((size_t)e_no * pb.in + i) * 16);
            const uint4 a = ldg_keep_u4(sp, keep), b = ldg_keep_u4(sp + 1, keep);
            const uint32_t ws[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int q2 = 0; q2 < 8; q2++) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&ws[q2]));
                st[2 * q2] = f.x;
                st[2 * q2 + 1] = f.y;
            }
        } else {
#pragma unroll
            for (int rho = 0; rho < 16; rho++)
                if (rho < P) st[rho] = __half2float(pb.st16[((size_t)e_no * pb.in + i) * P + rho]);
        }
    };
    if (!producer && tid < n_in) load_stats(tid, sel_stat);
    uint4 prb = make_uint4(0u, 0u, 0u, 0u);
    if (CUT == kCutSelect && !producer)
        prb = ldg_keep_u4(reinterpret_cast<const uint4*>(pb.probes + (size_t)e_no * EFFORT_PROBES_MAX) + tid, keep);
    uint4 nwv = make_uint4(0u, 0u, 0u, 0u);
    if (pb.norm_w && !producer) nwv = *reinterpret_cast<const uint4*>(pb.norm_w + 8 * tid);
    float4* my_tile4 = reinterpret_cast<float4*>(tiles + (size_t)(producer ? 0 : warp) * TF);
    if (producer) {
        for (int s = lane; s < kV3Desc; s += 32) {
            mbar_init((uint32_t)__cvta_generic_to_shared(&hdr.full_bar[s]), 1);
            mbar_init((uint32_t)__cvta_generic_to_shared(&hdr.empty_bar[s]), 1);
        }
        if (lane == 0) { hdr.ticket = 0u; hdr.sel_rows = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    } else if (CUT != kCutSelect) {  // the bisection has no rounds to hide the tile zeroing in
        for (int i = lane; i < TF / 4; i += 32) my_tile4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    V2_TRACE(1);
    pdl_wait();

    // ---- 1. the input vector ----
    const int vmode = pb.norm_w ? kVNorm : (pb.v2 ? kVSilu : kVPlain);
    const int slice_cols = min(pb.W, C - slice * pb.W);
    float vv[8];
#pragma unroll
    for (int m = 0; m < 8; m++) vv[m] = 0.f;
    float my_v = 0.f, my_x3 = 0.f, my_nw = 1.f;
    if (!producer) {
        const float* src = (vmode == kVPlain) ? pb.v_cut : pb.v;
        const float4 a = *reinterpret_cast<const float4*>(src + 8 * tid), b = *reinterpret_cast<const float4*>(src + 8 * tid + 4);
        vv[0] = a.x; vv[1] = a.y; vv[2] = a.z; vv[3] = a.w; vv[4] = b.x; vv[5] = b.y; vv[6] = b.z; vv[7] = b.w;
        if (vmode == kVSilu) {
            const float4 c = *reinterpret_cast<const float4*>(pb.v2 + 8 * tid), d = *reinterpret_cast<const float4*>(pb.v2 + 8 * tid + 4);
            const float x3[8] = {c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
            for (int m = 0; m < 8; m++) vv[m] = silu_mul(vv[m], x3[m]);
        }
        if (tid < n_in) {
            const int i = rsp + tid * RS;
            my_v = pb.v[i];
            if (vmode == kVSilu) my_x3 = pb.v2[i];
            if (vmode == kVNorm) my_nw = __half2float(pb.norm_w[i]);
        }
    } else if (pb.out_mode == kOutOverwrite) {
        // overwrite semantics, run by the (still idle) producer warp: zero this CTA's share of the slice's outputs, make
        // the zeros visible, arrive on the slice counter -- the counter is inspected right before the reductions
        const int n4 = slice_cols * SLOTS / 4, per = (n4 + RS - 1) / RS;
        float4* o4 = reinterpret_cast<float4*>(pb.out + (size_t)slice * pb.W * SLOTS);
        for (int x = rsp * per + lane; x < min(n4, (rsp + 1) * per); x += 32) o4[x] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            atomicAdd(pb.sync + 2 * slice, 1u);
        }
    }
    float denom = 1.f;
    if (vmode == kVNorm) {  // rmsNorm32fast (aux.metal:113-152); in == 8 * NT (the producer's vv are zeros)
        float ss = 0.f;
#pragma unroll
        for (int m = 0; m < 8; m++) ss += vv[m] * vv[m];
        ss = warp_sum_f(ss);
        if (lane == 0) hdr.red[warp] = ss;
        __syncthreads();
        float t = (lane < NW) ? hdr.red[lane] : 0.f;
        t = warp_sum_f(t);
        denom = sqrtf(t / (float)pb.in + pb.norm_eps);
        const uint32_t nw[4] = {nwv.x, nwv.y, nwv.z, nwv.w};
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const float2 wf = __half22float2(*reinterpret_cast<const __half2*>(&nw[m]));
            vv[2 * m] = (vv[2 * m] / denom) * wf.x;
            vv[2 * m + 1] = (vv[2 * m + 1] / denom) * wf.y;
        }
        my_v = (my_v / denom) * my_nw;
    } else if (vmode == kVSilu) {
        my_v = silu_mul(my_v, my_x3);
    }
    V2_TRACE(2);

    // ---- 2. cutoff ----
    float cutoff;
    if constexpr (CUT == kCutSelect) {
        uint32_t keys[4];
        score8(vv, prb, keys);
        if (producer) { keys[0] = keys[1] = keys[2] = keys[3] = 0u; }  // zero products are above no threshold
        V2_TRACE(3);
        cutoff = select_cutoff_slots(keys, EFFORT_PROBES_MAX - pb.q, hdr, tid, NW + 1, [&](int round) {
            if (!producer) {  // 1/8 of this warp's accumulator tile per round
                my_tile4[round * 64 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
                my_tile4[round * 64 + 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        });
    } else {
        float* vtmp = reinterpret_cast<float*>(p);  // the ring is idle until the rows stream
        if (vmode == kVSilu) {
            if (!producer) {
#pragma unroll
                for (int m = 0; m < 8; m++) vtmp[8 * tid + m] = vv[m];
            }
            __syncthreads();
        }
        if (tid < kCutThreads) {
            GroupProbes gpr;
            group_load_probes(pb.probes + (size_t)e_no * EFFORT_PROBES_MAX, EFFORT_PROBES_MAX, tid, gpr, keep);
            GroupProducts gp;
            if (vmode == kVNorm) group_score<true>(pb.v, gpr, EFFORT_PROBES_MAX, tid, gp, pb.norm_w, denom);
            else if (vmode == kVPlain) group_score<false>(pb.v_cut, gpr, EFFORT_PROBES_MAX, tid, gp, nullptr, 1.f);
            else group_score<false>(vtmp, gpr, EFFORT_PROBES_MAX, tid, gp, nullptr, 1.f);
            group_cutoff<1>(gp, EFFORT_PROBES_MAX, pb.q, hdr.cut, tid, nullptr);
        }
        __syncthreads();
        cutoff = hdr.cut.result;
        if (vmode == kVSilu) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // ring bytes written above, then by TMA
    }
    if (pb.cutoff_out && lb == 0 && tid == 0) *pb.cutoff_out = cutoff;
    V2_TRACE(6);

    const int seg_bytes = slice_cols * 2;
    const bool full_width = slice_cols == TW;  // 128 columns: one 8-byte piece per lane and row, rows 256 B apart
    const int lpr = pb.lpr, R = pb.R;
    const int rowslot = lane / lpr, lcol = lane % lpr;
    const bool col_ok = lcol * VEC < slice_cols;
    const uint32_t base_lane = (tiles_saddr + (uint32_t)(producer ? 0 : warp) * V3Smem::kTileBytes) | (uint32_t)(lane * 4);
    const uint64_t pol = l2_policy_evict_first();
    const uint4* bk16 = reinterpret_cast<const uint4*>(pb.bk + (size_t)pb.in * P * ((size_t)slice * pb.W));  // slice-major
    const uint32_t rs16 = (uint32_t)(seg_bytes >> 4);  // 16-byte units per row slice

    uint32_t seq = 0, tail_seq = 0, head = 0, free_b = ring_bytes;  // producer state (units issued / reclaimed, ring)
    // ---- passes over the inputs of this row split (one pass for every Mistral shape) ----
    for (int j0 = 0; j0 < n_in; j0 += NT) {
        const int j = j0 + tid;
        if (j0 > 0) {
            __syncthreads();
#pragma unroll
            for (int rho = 0; rho < 16; rho++) sel_stat[rho] = 0.f;
            my_v = 0.f;
            if (!producer && j < n_in) {
                load_stats(j, sel_stat);
                const int i = rsp + j * RS;
                my_v = pb.v[i];
                if (vmode == kVNorm) my_v = (my_v / denom) * __half2float(pb.norm_w[i]);
                else if (vmode == kVSilu) my_v = silu_mul(my_v, pb.v2[i]);
            }
        }
        // 3. selection mask of this thread's input (prepareDispatch, bucketMul.metal:66)
        if (!producer) {
            unsigned mask = 0u;
#pragma unroll
            for (int rho = 0; rho < 16; rho++)
                if (rho < P && j < n_in && row_selected(cutoff, sel_stat[rho], my_v)) mask |= 1u << rho;
            smask[tid] = mask;
            if (j < n_in) {
                const int i = rsp + j * RS;
                sbase[tid] = (uint32_t)(((size_t)e_no * pb.in * P * C + (size_t)i * P * slice_cols) >> 3);
                sval[tid] = pb.out_scale ? my_v * *pb.out_scale : my_v;  // the selection above used the unscaled input
            }
            const int wrows = __reduce_add_sync(0xffffffffu, __popc(mask));
            if (lane == 0 && wrows) atomicAdd(&hdr.sel_rows, wrows);
        }
        __syncthreads();  // masks / bases / multipliers (and the zeroed tiles) visible
        V2_TRACE(8);

        if (producer) {
            // ---- 4a. producer: one bulk copy per run of selected ranks, 32 inputs per step ----
            const int n_pass = min(NT, n_in - j0);
            for (int jb = 0; jb < n_pass; jb += 32) {
                const int jj = jb + lane;
                unsigned m = (jj < n_pass) ? smask[jj] : 0u;
                const uint32_t sb = (jj < n_pass) ? sbase[jj] : 0u;
                const float sv = (jj < n_pass) ? sval[jj] : 0.f;
                while (__any_sync(0xffffffffu, m != 0u)) {
                    // this lane's next run of consecutive selected ranks
                    int st = 0, len = 0;
                    if (m) { st = __ffs((int)m) - 1; len = __ffs((int)~(m >> st)) - 1; }
                    const uint32_t bytes = (uint32_t)(len * seg_bytes);
                    // lanes go in lane order; a step takes the longest prefix of pending lanes within kV3BatchBytes
                    uint32_t incl = bytes;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    const bool take = (m != 0u) && (incl <= (uint32_t)kV3BatchBytes || incl == bytes);  // always >= 1 unit
                    const unsigned takers = __ballot_sync(0xffffffffu, take);
                    // (a pending lane behind a non-taking pending lane must wait: keep the lane order)
                    const unsigned pending = __ballot_sync(0xffffffffu, m != 0u);
                    const unsigned first_skip = pending & ~takers;
                    const unsigned allowed = first_skip ? ((1u << (__ffs((int)first_skip) - 1)) - 1u) : 0xffffffffu;
                    const unsigned go = takers & allowed;
                    const int n_go = __popc(go);
                    const bool mine = (go >> lane) & 1u;
                    const uint32_t T = __shfl_sync(0xffffffffu, incl, 31 - __clz((int)go));  // bytes of the step
                    // ring space: contiguous T bytes (skip the tail of the ring if they do not fit before its end)
                    uint32_t skip = (head + T > ring_bytes) ? (ring_bytes - head) : 0u;
                    while (free_b < T + skip || seq + (uint32_t)n_go - tail_seq > (uint32_t)kV3Desc) {
                        const uint32_t ts = tail_seq & (kV3Desc - 1);
                        if (!mbar_wait((uint32_t)__cvta_generic_to_shared(&hdr.empty_bar[ts]), (tail_seq / kV3Desc) & 1u)) {
                            if (pb.err_flag && lane == 0) atomicExch(pb.err_flag, 3u);
                            break;
                        }
                        free_b += hdr.usize[ts];
                        tail_seq++;
                    }
                    const uint32_t start = skip ? 0u : head;
                    if (mine) {
                        const uint32_t my_seq = seq + (uint32_t)__popc(go & ((1u << lane) - 1u));
                        const uint32_t slot = my_seq & (kV3Desc - 1);
                        const uint32_t off = start + (incl - bytes);
                        const bool first = (go & ((1u << lane) - 1u)) == 0u;
                        hdr.usize[slot] = bytes + (first ? skip : 0u);
                        V3Desc d;
                        d.off = off; d.n = (uint32_t)len; d.val = sv; d.rstride = (uint32_t)seg_bytes;
                        *reinterpret_cast<uint4*>(&hdr.desc[slot]) = *reinterpret_cast<const uint4*>(&d);
                        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&hdr.full_bar[slot]);
                        mbar_expect_tx(bar, (int)bytes);
                        bulk_g2s(ring_saddr + off, bk16 + (size_t)(sb + (uint32_t)st * rs16), (int)bytes, bar, pol);
                        m &= ~(((1u << len) - 1u) << st);
                    }
                    seq += (uint32_t)n_go;
                    head = start + T;
                    free_b -= T + skip;
                    if (head >= ring_bytes) head = 0u;
                    __syncwarp();
                }
            }
            // poison units: one per consumer warp (n = 0, nothing to copy)
            while (seq + (uint32_t)NW - tail_seq > (uint32_t)kV3Desc) {
                const uint32_t ts = tail_seq & (kV3Desc - 1);
                if (!mbar_wait((uint32_t)__cvta_generic_to_shared(&hdr.empty_bar[ts]), (tail_seq / kV3Desc) & 1u)) break;
                free_b += hdr.usize[ts];
                tail_seq++;
            }
            if (lane < NW) {
                const uint32_t slot = (seq + (uint32_t)lane) & (kV3Desc - 1);
                hdr.usize[slot] = 0u;
                V3Desc d;
                d.off = 0u; d.n = 0u; d.val = 0.f; d.rstride = 0u;
                *reinterpret_cast<uint4*>(&hdr.desc[slot]) = *reinterpret_cast<const uint4*>(&d);
                mbar_arrive((uint32_t)__cvta_generic_to_shared(&hdr.full_bar[slot]));
            }
            seq += (uint32_t)NW;
            __syncwarp();
        } else {
            // ---- 4b. consumers ----
            const uint32_t ticket_saddr = (uint32_t)__cvta_generic_to_shared(&hdr.ticket);
#pragma unroll 1
            for (;;) {
                uint32_t t = 0;
                if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(t) : "r"(ticket_saddr) : "memory");
                t = __shfl_sync(0xffffffffu, t, 0);
                const uint32_t slot = t & (kV3Desc - 1);
                if (!mbar_wait((uint32_t)__cvta_generic_to_shared(&hdr.full_bar[slot]), (t / kV3Desc) & 1u)) {
                    if (pb.err_flag && lane == 0) atomicExch(pb.err_flag, 2u);
                    break;
                }
                const uint4 dq = *reinterpret_cast<const uint4*>(&hdr.desc[slot]);
                const int n = (int)dq.y;
                if (n == 0) {  // poison: release the slot and stop
                    if (lane == 0) mbar_arrive((uint32_t)__cvta_generic_to_shared(&hdr.empty_bar[slot]));
                    break;
                }
                const float val = __uint_as_float(dq.z);
                const uint32_t u0 = ring_saddr + dq.x;
                if (full_width) {
                    uint32_t a0 = u0 + (uint32_t)(lane * LB);
                    int r = 0;
                    for (; r + 4 <= n; r += 4, a0 += 4 * kRow) accumulate_unit_fp16<VEC, 4, kRow>(base_lane, val, a0);
                    switch (n - r) {
                        case 1: accumulate_unit_fp16<VEC, 1, kRow>(base_lane, val, a0); break;
                        case 2: accumulate_unit_fp16<VEC, 2, kRow>(base_lane, val, a0); break;
                        case 3: accumulate_unit_fp16<VEC, 3, kRow>(base_lane, val, a0); break;
                        default: break;
                    }
                } else {  // narrow slice: R rows per step, lanes past the slice idle
                    for (int st = 0; st * R < n; st++) {
                        const int r = st * R + rowslot;
                        const bool ok = (rowslot < R) && (r < n) && col_ok;
                        uint32_t ww[2] = {0u, 0u};
                        if (ok) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(ww[0]), "=r"(ww[1]) : "r"(u0 + (uint32_t)(r * seg_bytes + lcol * LB)));
                        accumulate_words<SLOTS, VEC>(base_lane, ok ? val : 0.f, ww);
                    }
                }
                __syncwarp();  // every lane has read the unit's bytes
                if (lane == 0) mbar_arrive((uint32_t)__cvta_generic_to_shared(&hdr.empty_bar[slot]));
            }
        }
    }
    __syncthreads();
    if (pb.sel_counts && slice == 0 && tid == 0) pb.sel_counts[rsp] = (uint32_t)hdr.sel_rows;
    V2_TRACE(9);

    // ---- 5. CTA epilogue: sum the 16 warp tiles and add into out (as bucket_mul_v2_kernel) ----
    if (!producer) {
        constexpr int NG = NT / TW, SPT = SLOTS / NG;
        static_assert(SPT == 4, "one 16-byte reduction per thread");
        const int cl = tid % TW, sg = tid / TW;
        const int k = cl >> 5, ln = cl & 31;
        float acc[SPT] = {0.f, 0.f, 0.f, 0.f};
        const bool col_on = (ln < lpr) && (ln * VEC + k < slice_cols);
        if (col_on) {
            for (int rs2 = 0; rs2 < R; rs2++) {
                const int word0 = (sg * SPT) * TW + k * 32 + ln + rs2 * lpr;
#pragma unroll
                for (int w = 0; w < NW; w++)
#pragma unroll
                    for (int s = 0; s < SPT; s++) acc[s] += tiles[(size_t)w * TF + word0 + s * TW];
            }
        }
        if (pb.out_mode == kOutOverwrite) {
            if (tid == 0) {
                const unsigned* cnt = pb.sync + 2 * slice;
                const unsigned long long t0 = gtime_ns();
                while (ld_acquire_u32(cnt) < (unsigned)RS) {
                    if (gtime_ns() - t0 > 2000000000ull) {
                        if (pb.err_flag) atomicExch(pb.err_flag, 1u);
                        break;
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");  // the 16 consumer warps
        }
        if (col_on) {
            const int col = slice * pb.W + ln * VEC + k;
            red_add_v4(pb.out + (size_t)col * SLOTS + sg * SPT, acc[0], acc[1], acc[2], acc[3]);
        }
        if (pb.out_mode == kOutOverwrite) {
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
            if (tid == 0) {
                unsigned* sy = pb.sync + 2 * slice;
                const unsigned old = atomicAdd(sy + 1, 1u);
                if (old == (unsigned)RS - 1u) { sy[0] = 0u; sy[1] = 0u; }
            }
        }
    }
    V2_TRACE(10);
}

}  // namespace effort
