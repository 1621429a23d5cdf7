// decode.cuh -- the non-GEMV steps either side of the hot path in the single-stream decode loop
// (reference: runNetwork.swift:113-209 with kernels from aux.metal / matrix.metal).  These are the "next"
// rows of the scope table: kept simple and HBM/latency-lean; every
// approximate GEMV in the loop goes through the bucketMul path.
#pragma once
#include "common.cuh"

namespace effort {

// fetchRow16to32 (aux.metal:355-363): x = float(tok_embeddings[token]); also the per-block partial sums of x^2 that
// the first layer's fused rmsNorm-on-load consumes (sumsq[gridDim.x], may be null).
// buffers a kernel of the decode chain clears for a LATER producer that accumulates into them (bucket_mul_v2 adds its
// result into `out` with reductions; the consumer of the previous use has finished by stream order)
struct ZeroList {
    float* p[6];
    int n[6];
};
__device__ __forceinline__ void zero_lists(const ZeroList& z, int part, int parts, int tid, int nt) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
        if (!z.p[k]) continue;
        const int n4 = z.n[k] >> 2;  // lengths are multiples of 4
        float4* q = reinterpret_cast<float4*>(z.p[k]);
        for (int i = part * nt + tid; i < n4; i += parts * nt) q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__global__ void __launch_bounds__(1024)
embed_kernel(const int* __restrict__ token, const __half* __restrict__ emb, int dim, int vocab, float* __restrict__ x,
             float* __restrict__ sumsq, const ZeroList zl) {
    __shared__ float red[32];
    pdl_trigger();
    pdl_wait();
    zero_lists(zl, blockIdx.x, gridDim.x, threadIdx.x, blockDim.x);
    int t = *token;
    t = (t < 0 || t >= vocab) ? 0 : t;  // never index the table with a garbage token (e.g. argmax over NaN logits)
    float ss = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += gridDim.x * blockDim.x) {
        const float xv = __half2float(emb[(size_t)t * dim + i]);
        x[i] = xv;
        ss += xv * xv;
    }
    ss = warp_sum_f(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
        v = warp_sum_f(v);
        if (threadIdx.x == 0 && sumsq) sumsq[blockIdx.x] = v;
    }
}

// h (+= add) ; out = rmsNormFast(h) * w     (rmsNorm32fast aux.metal:113-152: x / sqrt(mean(x^2) + 1e-5),
// then mulVec32by16 aux.metal:269).  One CTA of 1024 threads; `add` may be null.  h is updated in place
// when add != null (h.add(by:), runNetwork.swift:172,183).
__global__ void __launch_bounds__(1024)
add_rmsnorm_kernel(float* __restrict__ h, const float* __restrict__ add, const __half* __restrict__ w,
                   int dim, float eps, float* __restrict__ out) {
    __shared__ float red[32];
    __shared__ float total;
    pdl_trigger();
    pdl_wait();
    float ss = 0.f;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        float x = h[i];
        if (add) { x += add[i]; h[i] = x; }
        ss += x * x;
    }
    ss = warp_sum_f(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
        t = warp_sum_f(t);
        if (threadIdx.x == 0) total = t;
    }
    __syncthreads();
    const float denom = sqrtf(total / (float)dim + eps);
    for (int i = threadIdx.x; i < dim; i += blockDim.x) out[i] = (h[i] / denom) * __half2float(w[i]);
}

// h += add (last residual of the token loop, runNetwork.swift:183)
__global__ void add_kernel(float* __restrict__ h, const float* __restrict__ add, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) h[i] += add[i];
}

// silu32b (matrix.metal:25-34): out = x3 * x1 / (1 + exp(-x1))
__global__ void silu_mul_kernel(const float* __restrict__ x1, const float* __restrict__ x3, int n,
                                float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = x3[i] * x1[i] / (1.f + expf(-x1[i]));
}

// Attention for one new token (runNetwork.swift:136-164): rope_mx on q and k (aux.metal:218-231, HF
// rotate-half, theta = 1e6: freqs = 1e-6^(j/64), model.swift:693-717), append k/v to the cache, scores =
// q.k/sqrt(headDim) (dotSetScore2 aux.metal:397-447), softmax = exp(s)/sum(exp(s)) WITHOUT max subtraction
// (aux.metal:185-199), out = sum_t p_t * v_t (sumScores32 aux.metal:379-393).
// The reference materialises K/V repeated x4 for GQA (runNetwork.swift:136-137); here the 8 KV heads are
// stored once and head h reads KV head h / (n_heads/n_kv_heads) -- same arithmetic, a quarter of the bytes.
// One CTA per query head, 8 warps striding over the tokens, lane <-> 4 of the 128 head dims.
// pos is read from device memory so that the whole token can live in one replayable CUDA graph.
__global__ void __launch_bounds__(256)
attention_kernel(const float* __restrict__ xq, const float* __restrict__ xk, const float* __restrict__ xv,
                 float* __restrict__ kcache, float* __restrict__ vcache,  // [max_seq][n_kv][128]
                 const int* __restrict__ pos_dev, int n_heads, int n_kv, float theta, int max_seq,
                 float* __restrict__ attn_out, const ZeroList zl) {
    constexpr int HD = 128;
    __shared__ __align__(16) float q[HD];
    __shared__ __align__(16) float kcur[HD];  // this token's roped key (every head keeps its own copy: no cross-CTA wait)
    __shared__ float acc_s[8][HD];
    __shared__ float sum_s[8];
    const int h = blockIdx.x, kvh = h / (n_heads / n_kv);
    pdl_trigger();
    pdl_wait();
    int pos = *pos_dev;
    pos = pos < 0 ? 0 : (pos >= max_seq ? max_seq - 1 : pos);  // backstop: the host refuses steps past max_seq
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    zero_lists(zl, blockIdx.x, gridDim.x, tid, blockDim.x);
    // rope(q) and rope(k) into smem; the first head of each KV group also appends k/v to the cache
    if (tid < HD) {
        const int j = tid & 63;
        const float freq = powf(1e-6f * (1e6f / theta), (float)j / 64.f);  // theta=1e6 -> 1e-6^(j/64)
        const float ang = (float)pos * freq;
        const float c = cosf(ang), s = sinf(ang);
        const float* qh = xq + (size_t)h * HD;
        const float* kh = xk + (size_t)kvh * HD;
        const float qa = qh[tid], qb = (tid < 64) ? qh[tid + 64] : qh[tid - 64];
        const float ka = kh[tid], kb = (tid < 64) ? kh[tid + 64] : kh[tid - 64];
        q[tid] = (tid < 64) ? qa * c - qb * s : qa * c + qb * s;
        const float kr = (tid < 64) ? ka * c - kb * s : ka * c + kb * s;
        kcur[tid] = kr;
        if (h % (n_heads / n_kv) == 0) {
            kcache[((size_t)pos * n_kv + kvh) * HD + tid] = kr;
            vcache[((size_t)pos * n_kv + kvh) * HD + tid] = xv[(size_t)kvh * HD + tid];
        }
    }
    __syncthreads();
    const float4 q4 = *reinterpret_cast<const float4*>(&q[lane * 4]);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float sum = 0.f;
    const float scale = rsqrtf((float)HD);
    // warp w takes tokens w, w+8, ...; four of them per step so that eight 16-byte loads are in flight (the loop is a chain
    // of L2 round trips otherwise).  The entry this token appends is read from its source, not from the cache another CTA
    // of the group is writing.  Same summation order as one token per step.
    for (int tb = warp; tb <= pos; tb += 32) {
        float4 k4[4], v4[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int t = tb + 8 * i;
            k4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            v4[i] = k4[i];
            if (t < pos) {
                k4[i] = *reinterpret_cast<const float4*>(kcache + ((size_t)t * n_kv + kvh) * HD + lane * 4);
                v4[i] = *reinterpret_cast<const float4*>(vcache + ((size_t)t * n_kv + kvh) * HD + lane * 4);
            } else if (t == pos) {
                k4[i] = *reinterpret_cast<const float4*>(&kcur[lane * 4]);
                v4[i] = *reinterpret_cast<const float4*>(xv + (size_t)kvh * HD + lane * 4);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (tb + 8 * i <= pos) {
                float d = q4.x * k4[i].x + q4.y * k4[i].y + q4.z * k4[i].z + q4.w * k4[i].w;
                d = warp_sum_f(d);
                const float p = expf(d * scale);
                sum += p;
                acc.x += p * v4[i].x; acc.y += p * v4[i].y; acc.z += p * v4[i].z; acc.w += p * v4[i].w;
            }
        }
    }
    *reinterpret_cast<float4*>(&acc_s[warp][lane * 4]) = acc;
    if (lane == 0) sum_s[warp] = sum;
    __syncthreads();
    if (tid < HD) {
        float a = 0.f, s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) { a += acc_s[w][tid]; s += sum_s[w]; }
        attn_out[(size_t)h * HD + tid] = a / s;
    }
}

// greedy next token: argmax over the logits (mpsTopK index 0, runNetwork.swift:235-257); also bumps pos.
__global__ void __launch_bounds__(1024)
argmax_advance_kernel(const float* __restrict__ logits, int n, int* __restrict__ next_token,
                      int* __restrict__ pos_dev) {
    __shared__ float bv[32];
    __shared__ int bi[32];
    pdl_trigger();
    pdl_wait();
    float best = -INFINITY;
    int idx = 0x7fffffff;  // all-NaN logits leave it there: clamped to 0 below
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float x = logits[i];
        if (x > best || (x == best && i < idx)) { best = x; idx = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = bv[threadIdx.x]; idx = bi[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
        }
        if (threadIdx.x == 0) { *next_token = (idx < 0 || idx >= n) ? 0 : idx; *pos_dev = *pos_dev + 1; }
    }
}

// ---- MoE routing (runNetwork.swift:185-189): gateOut = basicMul(fxn, ffnGate) with fxn = rmsNormFast(h) * ffn_norm (v cast
// to fp16 first, helpers/mps.swift:19), the two largest gate logits (mpsTopK, helpers/mps.swift:49-80; ties: lower expert
// index first -- MPS is closed, unpinned) and gateVals.softmax() = exp(x) / sum(exp(x)) (aux.metal:185-199).  One CTA, one
// warp per expert (n_experts <= 8 per pass).  The indices stay on the device: the expert GEMVs read them as `expNo`.
__global__ void __launch_bounds__(256)
moe_gate_kernel(const float* __restrict__ h, const __half* __restrict__ norm_w, float eps, const __half* __restrict__ gate,
                int n_experts, int dim, uint32_t* __restrict__ gate_idx, float* __restrict__ gate_val, float* __restrict__ h_keep) {
    __shared__ float red[8];
    __shared__ float logit[64];
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float ss = 0.f;
    // both routed experts normalise the SAME hidden state (fxn is computed once, runNetwork.swift:176), but the first
    // expert's w2 already adds into h: keep a copy for the expert GEMVs to read
    for (int i = threadIdx.x; i < dim; i += blockDim.x) { const float x = h[i]; ss += x * x; h_keep[i] = x; }
    ss = warp_sum_f(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float t = (lane < 8) ? red[lane] : 0.f;
    t = warp_sum_f(t);
    const float denom = sqrtf(t / (float)dim + eps);
    for (int e = warp; e < n_experts && e < 64; e += 8) {
        const __half* row = gate + (size_t)e * dim;
        float acc = 0.f;
        for (int i = lane; i < dim; i += 32) {
            const float x = __half2float(__float2half_rn((h[i] / denom) * __half2float(norm_w[i])));
            acc = fmaf(x, __half2float(row[i]), acc);
        }
        acc = warp_sum_f(acc);
        if (lane == 0) logit[e] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int i0 = 0, i1 = -1;
        for (int e = 1; e < n_experts && e < 64; e++)
            if (logit[e] > logit[i0]) i0 = e;
        for (int e = 0; e < n_experts && e < 64; e++)
            if (e != i0 && (i1 < 0 || logit[e] > logit[i1])) i1 = e;
        if (i1 < 0) i1 = i0;
        const float e0 = expf(logit[i0]), e1 = expf(logit[i1]);
        gate_idx[0] = (uint32_t)i0;
        gate_idx[1] = (uint32_t)i1;
        gate_val[0] = e0 / (e0 + e1);
        gate_val[1] = e1 / (e0 + e1);
    }
}

// ---- fused head: final rmsNorm * w on load, dense lm_head GEMV (basicMul, helpers/mps.swift:14-47), greedy argmax ----
// (runNetwork.swift:206-209 + mpsTopK index 0, :235-257).  One warp per vocabulary row; every CTA keeps its best
// (logit, index) and the last CTA to finish (atomic ticket) reduces the per-CTA candidates -- lowest index wins ties, the
// same rule as argmax_advance_kernel -- writes the next token and advances the position.  v_norm_w == null: v is used as is.
__global__ void __launch_bounds__(256)
head_kernel(const float* __restrict__ h, const __half* __restrict__ norm_w, float eps, const __half* __restrict__ W,
            int out, int in, int row0, float* __restrict__ logits, float2* __restrict__ cand, unsigned* __restrict__ ticket,
            int* __restrict__ next_token, int* __restrict__ pos_dev, int do_argmax) {
    extern __shared__ float vs[];
    __shared__ float red[8];
    __shared__ float bestv[8];
    __shared__ int besti[8];
    __shared__ int is_last;
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float denom = 1.f;
    if (norm_w) {
        float ss = 0.f;
        for (int i = threadIdx.x; i < in; i += blockDim.x) { const float x = h[i]; ss += x * x; }
        ss = warp_sum_f(ss);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float t = (lane < 8) ? red[lane] : 0.f;
        t = warp_sum_f(t);
        denom = sqrtf(t / (float)in + eps);
    }
    for (int i = threadIdx.x; i < in; i += blockDim.x) {
        float x = h[i];
        if (norm_w) x = (x / denom) * __half2float(norm_w[i]);
        vs[i] = __half2float(__float2half_rn(x));  // v is cast to fp16 first (mps.swift:19)
    }
    __syncthreads();
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    for (int o = blockIdx.x * 8 + warp; o < out; o += gridDim.x * 8) {
        const __half* row = W + (size_t)o * in;
        float acc = 0.f;
        auto mac8 = [&](const uint4 d, int c) {
            const uint32_t ws[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&ws[j]));
                acc = fmaf(vs[c + 2 * j], f.x, acc);
                acc = fmaf(vs[c + 2 * j + 1], f.y, acc);
            }
        };
        if (in == 4096) {  // the whole 8 KB row in flight before the first multiply (same summation order)
            uint4 d[16];
#pragma unroll
            for (int k = 0; k < 16; k++) d[k] = ldg_stream_u4(row + lane * 8 + 256 * k);
#pragma unroll
            for (int k = 0; k < 16; k++) mac8(d[k], lane * 8 + 256 * k);
        } else {
            for (int c = lane * 8; c < in; c += 256) mac8(ldg_stream_u4(row + c), c);
        }
        acc = warp_sum_f(acc);
        if (lane == 0) logits[o] = acc;
        const int gi = row0 + o;
        if (acc > best || (acc == best && gi < bidx)) { best = acc; bidx = gi; }
    }
    if (!do_argmax) return;
    if (lane == 0) { bestv[warp] = best; besti[warp] = bidx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++)
            if (bestv[w] > best || (bestv[w] == best && besti[w] < bidx)) { best = bestv[w]; bidx = besti[w]; }
        cand[blockIdx.x] = make_float2(best, __int_as_float(bidx));
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    best = -INFINITY;
    bidx = 0x7fffffff;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
        float2 c;
        asm volatile("ld.volatile.global.v2.f32 {%0,%1}, [%2];" : "=f"(c.x), "=f"(c.y) : "l"(cand + b));
        const int ci = __float_as_int(c.y);
        if (c.x > best || (c.x == best && ci < bidx)) { best = c.x; bidx = ci; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane == 0) { bestv[warp] = best; besti[warp] = bidx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < 8; w++)
            if (bestv[w] > best || (bestv[w] == best && besti[w] < bidx)) { best = bestv[w]; bidx = besti[w]; }
        *next_token = (bidx == 0x7fffffff) ? 0 : bidx;  // all-NaN logits: token 0, never an out-of-range index
        *pos_dev = *pos_dev + 1;
        *ticket = 0u;  // re-armed for the next token (graph replay)
    }
}

}  // namespace effort
