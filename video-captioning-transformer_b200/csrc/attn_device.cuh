// Warp-level building blocks of the attention core, shared by the stand-alone kernels (attn_core.cu) and the
// fused in-projection + attention kernel (attn_fused.cu).  All operate on fp32 tiles in shared memory.
#pragma once
#include "common.cuh"

namespace vct {

constexpr int RB = 4;                      // query rows (or keys) register-blocked per pass

// out[i][j] = sum_c rows[i][c] * keys[j][c] for this warp's rows [r0, r1) and every key j (lane = key).
// The lane's key row lives in registers; rows are read as broadcast float4.
template <int DHP>
__device__ __forceinline__ void rows_dot_keys(const float* __restrict__ rows, int QS, const float* __restrict__ keys, int KS,
                                              int Lk, int LkP, int r0, int r1, int lane, float* __restrict__ out) {
    for (int slot = 0; slot * 32 < Lk; ++slot) {
        const int j = lane + 32 * slot;
        float kreg[DHP];
#pragma unroll
        for (int c = 0; c < DHP; ++c) kreg[c] = j < Lk ? keys[j * KS + c] : 0.f;
#pragma unroll 1
        for (int i = r0; i < r1; ++i) {
            const float4* qrow = reinterpret_cast<const float4*>(rows + i * QS);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < DHP / 4; ++c4) {
                const float4 qv = qrow[c4];
                a0 = fmaf(qv.x, kreg[4 * c4 + 0], a0);
                a1 = fmaf(qv.y, kreg[4 * c4 + 1], a1);
                a2 = fmaf(qv.z, kreg[4 * c4 + 2], a2);
                a3 = fmaf(qv.w, kreg[4 * c4 + 3], a3);
            }
            if (j < Lk) out[i * LkP + j] = (a0 + a1) + (a2 + a3);
        }
    }
}

// masked softmax of one score row held as (lane, lane + 32); returns probabilities (0 where masked)
__device__ __forceinline__ void softmax_row(const float* __restrict__ srow, int Lk, int i, int lane, bool causal,
                                            const unsigned char* __restrict__ pad_row, float scale, float& p0, float& p1) {
    float s[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int j = lane + 32 * t;
        const bool ok = j < Lk && !(causal && j > i) && !(pad_row != nullptr && pad_row[j]);
        s[t] = ok ? srow[j] * scale : -INFINITY;
    }
    const float m = warp_max(fmaxf(s[0], s[1]));
    if (m == -INFINITY) { p0 = p1 = 0.f; return; }    // fully masked row (cannot happen on this path, Q8)
    const float e0 = s[0] == -INFINITY ? 0.f : expf(s[0] - m);
    const float e1 = s[1] == -INFINITY ? 0.f : expf(s[1] - m);
    const float inv = 1.f / warp_sum(e0 + e1);
    p0 = e0 * inv;
    p1 = e1 * inv;
}

// acc[k][cc] = sum_j wt[j][ib + k] * mat[j][lane + 32 cc]  for k < RB   (wt transposed: [Lk][LqP])
template <int DHP>
__device__ __forceinline__ void weighted_rows(const float* __restrict__ wt, int LqP, const float* __restrict__ mat,
                                              int MS, int Lk, int ib, int nrows, int lane, float (&acc)[RB][DHP / 32]) {
#pragma unroll
    for (int k = 0; k < RB; ++k)
#pragma unroll
        for (int cc = 0; cc < DHP / 32; ++cc) acc[k][cc] = 0.f;
    for (int j = 0; j < Lk; ++j) {
        float w[RB];
#pragma unroll
        for (int k = 0; k < RB; ++k) w[k] = k < nrows ? wt[j * LqP + ib + k] : 0.f;
#pragma unroll
        for (int cc = 0; cc < DHP / 32; ++cc) {
            const float v = mat[j * MS + lane + 32 * cc];
#pragma unroll
            for (int k = 0; k < RB; ++k) acc[k][cc] = fmaf(w[k], v, acc[k][cc]);
        }
    }
}

// dropout keep-multipliers of one probability row for the key slots (lane, lane + 32): lane t draws the 8
// decisions of keys 8t..8t+7 (one Philox call per row and warp), the bits are exchanged with shuffles.
// Index space: row r = (b*H + h)*Lq + i owns groups r*8 .. r*8+7 (Lk <= 64).
__device__ __forceinline__ void row_dropout(const Rng& rng, unsigned int site, long long row, int lane, float& sc0, float& sc1) {
    const uint32_t bits = dropout_bits8(rng, site, (unsigned long long)row * 8ull + (unsigned long long)(lane & 7));
    const uint32_t m0 = __shfl_sync(0xffffffffu, bits, lane >> 3);
    const uint32_t m1 = __shfl_sync(0xffffffffu, bits, 4 + (lane >> 3));
    sc0 = ((m0 >> (lane & 7)) & 1u) ? rng.inv_keep : 0.f;
    sc1 = ((m1 >> (lane & 7)) & 1u) ? rng.inv_keep : 0.f;
}


}  // namespace vct
