// q4.cuh -- GPU Q4 converter (reference: q4_draft.py:70-322, function convert(core2), after the global
// top-2 % outlier extraction which the host does once, see effort_b200/convert.py).
//   per input row i, groups of 8 consecutive outputs sorted by |w| descending (np.argsort(-|w|) on 8 elements =
//   insertion sort = stable: ties keep the lower index first, q4_draft.py:119-138); row order i*8 + rank
//   (:150-168); nibble = (8 if w < 0) + pos, 4 nibbles per 16-bit word, first group in the TOP nibble (:264-303);
//   bucket.stats = (avg, avg) with avg = float16(np.mean(|values|)) computed with float32 pairwise summation
//   (:180-195, :244-245); probes = diag(core) (:240).
#pragma once
#include "common.cuh"

namespace effort {

// thread <-> (input row i, word column c): 32 consecutive outputs = 4 groups of 8.
__global__ void __launch_bounds__(256)
q4_bucketize_kernel(const uint16_t* __restrict__ wT, int in, int out, uint16_t* __restrict__ buckets,
                    uint16_t* __restrict__ absvals /* [in*8][out/8] fp16 |w| per rank row */) {
    const int C4 = out / 32, G8 = out / 8;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)in * C4) return;
    const int i = (int)(t / C4), c = (int)(t % C4);
    const uint4* src = reinterpret_cast<const uint4*>(wT + (size_t)i * out + (size_t)c * 32);
    uint32_t word[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const uint4 d = src[m];
        const uint32_t raw[4] = {d.x, d.y, d.z, d.w};
        uint16_t b[8];
#pragma unroll
        for (int k = 0; k < 4; k++) { b[2 * k] = (uint16_t)(raw[k] & 0xFFFFu); b[2 * k + 1] = (uint16_t)(raw[k] >> 16); }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint16_t aj = b[j] & 0x7FFFu;
            int rank = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint16_t ak = b[k] & 0x7FFFu;
                rank += (ak > aj || (ak == aj && k < j)) ? 1 : 0;
            }
            // "8 if value < 0" (q4_draft.py:265): -0.0 and NaN are not < 0
            const bool neg = (b[j] & 0x8000u) && aj != 0 && aj <= 0x7C00u;
            const uint32_t nib = (neg ? 8u : 0u) | (uint32_t)j;
#pragma unroll
            for (int r = 0; r < 8; r++)
                if (r == rank) {
                    word[r] |= nib << (12 - 4 * m);
                    absvals[((size_t)i * 8 + r) * G8 + (size_t)c * 4 + m] = aj;
                }
        }
    }
#pragma unroll
    for (int r = 0; r < 8; r++) buckets[((size_t)i * 8 + r) * C4 + c] = (uint16_t)word[r];
}

// numpy's float32 pairwise summation (numpy/core/src/umath/loops_utils.h.src, PW_BLOCKSIZE = 128) over a
// contiguous fp16 vector, so that the mean matches np.mean(float16 array) bit for bit.
__device__ float np_pairwise_sum_f16(const __half* a, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; i++) res = __fadd_rn(res, __half2float(a[i]));
        return res;
    } else if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = __half2float(a[j]);
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] = __fadd_rn(r[j], __half2float(a[i + j]));
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; i++) res = __fadd_rn(res, __half2float(a[i]));
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return __fadd_rn(np_pairwise_sum_f16(a, n2), np_pairwise_sum_f16(a + n2, n - n2));
    }
}

__global__ void q4_stats_kernel(const __half* __restrict__ absvals, size_t rows, int n, float* __restrict__ stats2) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float s = np_pairwise_sum_f16(absvals + r * n, n);
    const float mean = __half2float(__float2half_rn(__fdiv_rn(s, (float)n)));  // float16(ret / rcount)
    stats2[r * 2 + 0] = mean;
    stats2[r * 2 + 1] = mean;
}

__global__ void q4_probes_kernel(const uint16_t* __restrict__ wT, int out, int n, uint16_t* __restrict__ probes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) probes[i] = wT[(size_t)i * out + i];
}

}  // namespace effort
