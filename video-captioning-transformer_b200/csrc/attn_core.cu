// Attention core: softmax(scale * Q K^T + mask) V, forward and backward.
//
// One CTA (4 warps) per (batch, head).  Sequence lengths of the shipped configs are 13..33, so the
// whole head -- Q, K, V (and dO) -- is staged once in shared memory as fp32; each warp owns a contiguous block
// of query rows.  (Sequences beyond 64 rows: the tiled kernels under "Long sequences" below.)
//   scores   : one key per lane; the lane keeps its K row in REGISTERS and the query row is read from shared
//              memory as broadcast float4 -> FMA-bound, not shared-memory-bound
//   softmax  : warp-shuffle max / sum over the key lanes; key-padding and causal masks are implicit
//   P V      : one head-dim column per lane, 4 query rows register-blocked per pass over the keys
//   backward : recomputes P (same Philox dropout mask), dP with V rows in registers, dS; dQ like P V;
//              then the keys are split over the warps for dK = dS^T Q and dV = P^T dO (no atomics)
#include <cstdlib>

#include "attn_device.cuh"

using namespace vct;

namespace {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr size_t kSmemBudget = 200 * 1024;

struct Dims {
    int B, H, Lq, Lk, dh;
    long long q_ld, k_ld, v_ld, o_ld, do_ld, dq_ld, dk_ld, dv_ld;
    long long q_bs, k_bs, v_bs, o_bs, do_bs, dq_bs, dk_bs, dv_bs;   // batch strides (elements)
    int causal;
    float scale;
};

// rows x dh tile (global, dtype T) -> shared fp32 [rows][stride], zero-padded to DHP columns
template <typename T, int DHP>
__device__ __forceinline__ void stage_tile(const T* __restrict__ g, long long ld, int rows, int dh, float* s, int stride) {
    constexpr int NV = DHP / 4;
    for (int idx = threadIdx.x; idx < rows * NV; idx += kThreads) {
        const int r = idx / NV, c = (idx % NV) * 4;
        float4 v = c < dh ? ld4(g + (long long)r * ld + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        float* d = s + r * stride + c;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
}

template <typename T, int DHP>
__global__ void __launch_bounds__(kThreads)
attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ o,
                const unsigned char* __restrict__ key_pad, float* __restrict__ probs, Dims D, float drop_p,
                const unsigned long long* __restrict__ rng_state, unsigned int site) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x, b = bh / D.H, h = bh % D.H;
    const int dh = D.dh, Lk = D.Lk, Lq = D.Lq;
    constexpr int KS = DHP + 1;
    const int LkP = (Lk + 3) & ~3, LqP = (Lq + 3) & ~3;
    float* Qs = sm;                       // [Lq][DHP]
    float* Ks = Qs + Lq * DHP;            // [Lk][KS]
    float* Vs = Ks + Lk * KS;             // [Lk][KS]
    float* Ss = Vs + Lk * KS;             // [Lq][LkP]   raw scores
    float* Pt = Ss + Lq * LkP;            // [Lk][LqP]   (dropped) probabilities, transposed
    const Rng rng = make_rng(rng_state, drop_p);

    stage_tile<T, DHP>(q + (long long)b * D.q_bs + h * dh, D.q_ld, Lq, dh, Qs, DHP);
    stage_tile<T, DHP>(k + (long long)b * D.k_bs + h * dh, D.k_ld, Lk, dh, Ks, KS);
    stage_tile<T, DHP>(v + (long long)b * D.v_bs + h * dh, D.v_ld, Lk, dh, Vs, KS);
    __syncthreads();

    const int R = (Lq + kWarps - 1) / kWarps;
    const int r0 = warp * R, r1 = min(Lq, r0 + R);
    if (r0 >= r1) return;
    rows_dot_keys<DHP>(Qs, DHP, Ks, KS, Lk, LkP, r0, r1, lane, Ss);
    __syncwarp();
    const unsigned char* pad_row = key_pad ? key_pad + (long long)b * Lk : nullptr;
    for (int i = r0; i < r1; ++i) {
        float p0, p1;
        softmax_row(Ss + i * LkP, Lk, i, lane, D.causal != 0, pad_row, D.scale, p0, p1);
        const long long pbase = ((long long)bh * Lq + i) * Lk;
        if (probs) {
            if (lane < Lk) probs[pbase + lane] = p0;
            if (lane + 32 < Lk) probs[pbase + lane + 32] = p1;
        }
        if (rng.p > 0.f) {
            float sc0, sc1;
            row_dropout(rng, site, (long long)bh * Lq + i, lane, sc0, sc1);
            p0 *= sc0;
            p1 *= sc1;
        }
        if (lane < Lk) Pt[lane * LqP + i] = p0;
        if (lane + 32 < Lk) Pt[(lane + 32) * LqP + i] = p1;
    }
    __syncwarp();
    for (int ib = r0; ib < r1; ib += RB) {
        const int nrows = min(RB, r1 - ib);
        float acc[RB][DHP / 32];
        weighted_rows<DHP>(Pt, LqP, Vs, KS, Lk, ib, nrows, lane, acc);
#pragma unroll
        for (int kk = 0; kk < RB; ++kk) {
            if (kk < nrows) {
                T* orow = o + (long long)b * D.o_bs + (long long)(ib + kk) * D.o_ld + h * dh;
#pragma unroll
                for (int cc = 0; cc < DHP / 32; ++cc) {
                    const int c = lane + 32 * cc;
                    if (c < dh) orow[c] = from_f32<T>(acc[kk][cc]);
                }
            }
        }
    }
}

template <typename T, int DHP>
__global__ void __launch_bounds__(kThreads, 4)      // <= 128 registers: four CTAs per SM, so that B * H = 512 CTAs are one wave
attn_bwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ d_o,
                T* __restrict__ dq, T* __restrict__ dk, T* __restrict__ dv, const unsigned char* __restrict__ key_pad,
                Dims D, float drop_p, const unsigned long long* __restrict__ rng_state, unsigned int site,
                float* __restrict__ dbias, float* __restrict__ dbias_part, unsigned int* dbias_cnt) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    __shared__ float colpart[kWarps][3][DHP];      // per-warp column sums of this head's dq, dk, dv rows (in-proj bias gradient)
    __shared__ bool is_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x, b = bh / D.H, h = bh % D.H;
    const int dh = D.dh, Lk = D.Lk, Lq = D.Lq;
    // (accumulated in shared memory, not registers: the kernel sits exactly at the 128-register limit that lets four
    // CTAs share an SM, and 512 CTAs need all 4 x 148 slots to run as one wave)
#pragma unroll
    for (int cc = 0; cc < DHP / 32; ++cc) colpart[warp][0][lane + 32 * cc] = colpart[warp][1][lane + 32 * cc] = colpart[warp][2][lane + 32 * cc] = 0.f;
    constexpr int KS = DHP + 1;
    const int LkP = (Lk + 3) & ~3, LqP = (Lq + 3) & ~3;
    float* Qs = sm;                        // [Lq][DHP]
    float* dOs = Qs + Lq * DHP;            // [Lq][DHP]
    float* Ks = dOs + Lq * DHP;            // [Lk][KS]
    float* Vs = Ks + Lk * KS;              // [Lk][KS]
    float* Ss = Vs + Lk * KS;              // [Lq][LkP]   raw scores
    float* dPs = Ss + Lq * LkP;            // [Lq][LkP]   dO . V
    float* dSt = dPs + Lq * LkP;           // [Lk][LqP]   dS transposed
    float* Pdt = dSt + Lk * LqP;           // [Lk][LqP]   dropped probabilities transposed
    const Rng rng = make_rng(rng_state, drop_p);

    stage_tile<T, DHP>(q + (long long)b * D.q_bs + h * dh, D.q_ld, Lq, dh, Qs, DHP);
    stage_tile<T, DHP>(d_o + (long long)b * D.do_bs + h * dh, D.do_ld, Lq, dh, dOs, DHP);
    stage_tile<T, DHP>(k + (long long)b * D.k_bs + h * dh, D.k_ld, Lk, dh, Ks, KS);
    stage_tile<T, DHP>(v + (long long)b * D.v_bs + h * dh, D.v_ld, Lk, dh, Vs, KS);
    __syncthreads();

    const int R = (Lq + kWarps - 1) / kWarps;
    const int r0 = warp * R, r1 = min(Lq, r0 + R);
    if (r0 < r1) {
        rows_dot_keys<DHP>(Qs, DHP, Ks, KS, Lk, LkP, r0, r1, lane, Ss);
        rows_dot_keys<DHP>(dOs, DHP, Vs, KS, Lk, LkP, r0, r1, lane, dPs);
        __syncwarp();
        const unsigned char* pad_row = key_pad ? key_pad + (long long)b * Lk : nullptr;
        for (int i = r0; i < r1; ++i) {
            float p[2];
            softmax_row(Ss + i * LkP, Lk, i, lane, D.causal != 0, pad_row, D.scale, p[0], p[1]);
            float sc[2] = {1.f, 1.f}, dP[2] = {0.f, 0.f};
            float dsum = 0.f;
            if (rng.p > 0.f) row_dropout(rng, site, (long long)bh * Lq + i, lane, sc[0], sc[1]);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int j = lane + 32 * t;
                if (j < Lk) {
                    dP[t] = dPs[i * LkP + j] * sc[t];
                    dsum += p[t] * dP[t];
                }
            }
            dsum = warp_sum(dsum);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int j = lane + 32 * t;
                if (j < Lk) {
                    dSt[j * LqP + i] = p[t] * (dP[t] - dsum) * D.scale;
                    Pdt[j * LqP + i] = p[t] * sc[t];
                }
            }
        }
        __syncwarp();
        // dQ[i][c] = sum_j dS[i][j] K[j][c]
        for (int ib = r0; ib < r1; ib += RB) {
            const int nrows = min(RB, r1 - ib);
            float acc[RB][DHP / 32];
            weighted_rows<DHP>(dSt, LqP, Ks, KS, Lk, ib, nrows, lane, acc);
#pragma unroll
            for (int kk = 0; kk < RB; ++kk) {
                if (kk < nrows) {
                    T* row = dq + (long long)b * D.dq_bs + (long long)(ib + kk) * D.dq_ld + h * dh;
#pragma unroll
                    for (int cc = 0; cc < DHP / 32; ++cc) {
                        const int c = lane + 32 * cc;
                        if (c < dh) row[c] = from_f32<T>(acc[kk][cc]);
                        if (dbias != nullptr) colpart[warp][0][c] += acc[kk][cc];
                    }
                }
            }
        }
    }
    __syncthreads();
    // keys split over the warps: dK[j][c] = sum_i dS[i][j] Q[i][c],  dV[j][c] = sum_i Pd[i][j] dO[i][c]
    const int KR = (Lk + kWarps - 1) / kWarps;
    const int j0 = warp * KR, j1 = min(Lk, j0 + KR);
    for (int jb = j0; jb < j1; jb += RB) {
        const int nk = min(RB, j1 - jb);
        float ak[RB][DHP / 32], av[RB][DHP / 32];
#pragma unroll
        for (int kk = 0; kk < RB; ++kk)
#pragma unroll
            for (int cc = 0; cc < DHP / 32; ++cc) ak[kk][cc] = av[kk][cc] = 0.f;
        for (int i = 0; i < Lq; ++i) {
            float ws[RB], wp[RB];
#pragma unroll
            for (int kk = 0; kk < RB; ++kk) {
                ws[kk] = kk < nk ? dSt[(jb + kk) * LqP + i] : 0.f;
                wp[kk] = kk < nk ? Pdt[(jb + kk) * LqP + i] : 0.f;
            }
#pragma unroll
            for (int cc = 0; cc < DHP / 32; ++cc) {
                const float qv = Qs[i * DHP + lane + 32 * cc], gv = dOs[i * DHP + lane + 32 * cc];
#pragma unroll
                for (int kk = 0; kk < RB; ++kk) {
                    ak[kk][cc] = fmaf(ws[kk], qv, ak[kk][cc]);
                    av[kk][cc] = fmaf(wp[kk], gv, av[kk][cc]);
                }
            }
        }
#pragma unroll
        for (int kk = 0; kk < RB; ++kk) {
            if (kk < nk) {
                T* krow = dk + (long long)b * D.dk_bs + (long long)(jb + kk) * D.dk_ld + h * dh;
                T* vrow = dv + (long long)b * D.dv_bs + (long long)(jb + kk) * D.dv_ld + h * dh;
#pragma unroll
                for (int cc = 0; cc < DHP / 32; ++cc) {
                    const int c = lane + 32 * cc;
                    if (c < dh) { krow[c] = from_f32<T>(ak[kk][cc]); vrow[c] = from_f32<T>(av[kk][cc]); }
                    if (dbias != nullptr) {
                        colpart[warp][1][c] += ak[kk][cc];
                        colpart[warp][2][c] += av[kk][cc];
                    }
                }
            }
        }
    }
    // ---- in-projection bias gradient: column sums of dq | dk | dv over all (batch, position) rows, without a separate
    //      pass over the gradient matrix.  CTA (b, h) writes its [3][dh] sums to dbias_part[b]; the last CTA of head h to
    //      finish adds the B partials in batch order (deterministic) into dbias (layout q | k | v, each H * dh wide). ----
    if (dbias == nullptr) return;
    __syncthreads();
    const int Hd = D.H * dh;
    for (int t = threadIdx.x; t < 3 * dh; t += kThreads) {
        const int sec = t / dh, c = t % dh;
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) a += colpart[w][sec][c];
        dbias_part[(long long)b * 3 * Hd + sec * Hd + h * dh + c] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(dbias_cnt + h, 1u);
        is_last = (prev == (unsigned int)D.B - 1u);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int t = threadIdx.x; t < 3 * dh; t += kThreads) {
            const int sec = t / dh, c = t % dh;
            const long long col = sec * Hd + h * dh + c;
            float a = 0.f;
            int bb = 0;
            for (; bb + 16 <= D.B; bb += 16) {                  // 16 independent loads in flight, summed in batch order
                float v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = __ldcg(dbias_part + (long long)(bb + u) * 3 * Hd + col);
#pragma unroll
                for (int u = 0; u < 16; ++u) a += v[u];
            }
            for (; bb < D.B; ++bb) a += __ldcg(dbias_part + (long long)bb * 3 * Hd + col);
            dbias[col] = a;
        }
        if (threadIdx.x == 0) dbias_cnt[h] = 0u;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Long sequences (Lq or Lk > 64: more frames than CLIP4Clip's 12, captions beyond 64 word pieces -- the reference has no
// limit: train.py does not truncate captions and T is whatever the feature file holds).  Same arithmetic as the
// kernels above, tiled over the sequence:
//   forward     CTA = (batch * head, tile of 16 query rows, 4 per warp).  K, then V, stream through shared memory in
//               blocks of 32 rows (lane = key for the scores, lane = head-dim column for P V); the tile's complete score
//               rows [16][Lk] stay in shared memory, so the softmax is the plain two-pass one
//   backward Q  same tiling: recomputes P and dP = dO V^T, dS = P (dP - sum_j P dP) scale, dQ = dS K, and leaves
//               (row max, 1 / row sum, sum_j P dP) of every query row in `row_stats`
//   backward KV CTA = (batch * head, tile of 16 keys): S^T and dP^T of its keys against every query block, P rebuilt from
//               the row statistics, dK = dS^T Q, dV = Pd^T dO.  No atomics: bit-reproducible
// Dropout index space of this path: probability row r = (b*H + h)*Lq + i owns ceil(Lk / 8) groups of 8 keys, group g of
// row r is Philox index r * ceil(Lk / 8) + g (the short kernels give every row exactly 8 groups).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kLongRows = 16;       // tile rows per CTA
constexpr int kLongBlk = 32;        // rows of the streamed operand per block (one per lane)
constexpr int kLongMax = 1024;      // longest sequence: 2 x [16][Lk] fp32 score tiles + operands fit 200 KB

__device__ __forceinline__ uint32_t long_group_bits(const Rng& rng, unsigned int site, long long row, int nG, int g) {
    return g < nG ? dropout_bits8(rng, site, (unsigned long long)row * (unsigned long long)nG + (unsigned long long)g) : 0u;
}

// masked, scaled scores of one row -> exponentials (in place); m = row maximum (-inf: every key masked), inv = 1 / row sum
__device__ __forceinline__ void long_softmax_row(float* __restrict__ srow, int Lk, int gi, int lane, bool causal,
                                                 const unsigned char* __restrict__ pad_row, float scale, float& m, float& inv) {
    float mx = -INFINITY;
    for (int j = lane; j < Lk; j += 32) {
        const bool ok = !(causal && j > gi) && !(pad_row != nullptr && pad_row[j]);
        float s = -INFINITY;
        if (ok) s = srow[j] * scale;             // (columns of causally skipped blocks were never written: not read either)
        srow[j] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Lk; j += 32) {
        const float s = srow[j];
        const float e = s == -INFINITY ? 0.f : expf(s - mx);
        srow[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    m = mx;
    inv = sum > 0.f ? 1.f / sum : 0.f;
}

// acc[k][cc] += sum_{j < n} w[(r0 + k) * ldw + j] * mat[j][lane + 32 cc]  for k < nrows <= 4  (w broadcast, mat coalesced)
template <int DHP>
__device__ __forceinline__ void long_accumulate(const float* __restrict__ w, int ldw, int r0, int nrows,
                                                const float* __restrict__ mat, int MS, int n, int lane,
                                                float (&acc)[4][DHP / 32]) {
    for (int j = 0; j < n; ++j) {
        float wk[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) wk[kk] = kk < nrows ? w[(r0 + kk) * ldw + j] : 0.f;
#pragma unroll
        for (int cc = 0; cc < DHP / 32; ++cc) {
            const float x = mat[j * MS + lane + 32 * cc];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) acc[kk][cc] = fmaf(wk[kk], x, acc[kk][cc]);
        }
    }
}

template <typename T, int DHP>
__device__ __forceinline__ void long_store_rows(T* __restrict__ base, long long ld, int nrows, int dh, int lane,
                                                const float (&acc)[4][DHP / 32]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        if (kk < nrows) {
#pragma unroll
            for (int cc = 0; cc < DHP / 32; ++cc) {
                const int c = lane + 32 * cc;
                if (c < dh) base[(long long)kk * ld + c] = from_f32<T>(acc[kk][cc]);
            }
        }
    }
}

template <int DHP>
__device__ __forceinline__ void long_zero(float (&acc)[4][DHP / 32]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int cc = 0; cc < DHP / 32; ++cc) acc[kk][cc] = 0.f;
}

template <typename T, int DHP>
__global__ void __launch_bounds__(kThreads)
attn_long_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ o,
                     const unsigned char* __restrict__ key_pad, float* __restrict__ probs, Dims D, float drop_p,
                     const unsigned long long* __restrict__ rng_state, unsigned int site) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x, b = bh / D.H, h = bh % D.H;
    const int dh = D.dh, Lk = D.Lk, Lq = D.Lq;
    const int q0 = blockIdx.y * kLongRows, nq = min(kLongRows, Lq - q0);
    constexpr int KS = DHP + 1;
    const int LkP = (Lk + 3) & ~3;
    float* Qs = sm;                               // [16][DHP]   the tile's query rows
    float* Bs = Qs + kLongRows * DHP;             // [32][KS]    streamed block of K, later V
    float* Ss = Bs + kLongBlk * KS;               // [16][LkP]   scores -> (dropped) probabilities
    const Rng rng = make_rng(rng_state, drop_p);
    const int r0 = warp * 4, r1 = min(nq, r0 + 4);                  // this warp's rows of the tile
    const int kend = D.causal ? min(Lk, q0 + nq) : Lk;              // keys >= kend are masked for every row of the tile
    const T* kb = k + (long long)b * D.k_bs + h * dh;
    const T* vb = v + (long long)b * D.v_bs + h * dh;
    stage_tile<T, DHP>(q + (long long)b * D.q_bs + (long long)q0 * D.q_ld + h * dh, D.q_ld, nq, dh, Qs, DHP);
    for (int j0 = 0; j0 < kend; j0 += kLongBlk) {
        const int nk = min(kLongBlk, Lk - j0);
        __syncthreads();
        stage_tile<T, DHP>(kb + (long long)j0 * D.k_ld, D.k_ld, nk, dh, Bs, KS);
        __syncthreads();
        if (r0 < r1) rows_dot_keys<DHP>(Qs, DHP, Bs, KS, nk, LkP, r0, r1, lane, Ss + j0);
    }
    __syncwarp();
    const unsigned char* pad_row = key_pad ? key_pad + (long long)b * Lk : nullptr;
    const int nG = (Lk + 7) >> 3;
    for (int il = r0; il < r1; ++il) {
        const int gi = q0 + il;
        float* srow = Ss + il * LkP;
        float m, inv;
        long_softmax_row(srow, Lk, gi, lane, D.causal != 0, pad_row, D.scale, m, inv);
        const long long row = (long long)bh * Lq + gi;
        for (int c0 = 0; c0 < Lk; c0 += 256) {                      // 256 keys = 32 groups: one Philox call per lane
            const uint32_t bits = rng.p > 0.f ? long_group_bits(rng, site, row, nG, (c0 >> 3) + lane) : 0xFFu;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (c0 + 32 * t < Lk) {                             // warp-uniform
                    const uint32_t gb = __shfl_sync(0xffffffffu, bits, 4 * t + (lane >> 3));
                    const int j = c0 + 32 * t + lane;
                    if (j < Lk) {
                        float p = srow[j] * inv;
                        if (probs) probs[row * Lk + j] = p;
                        if (rng.p > 0.f) p = ((gb >> (lane & 7)) & 1u) ? p * rng.inv_keep : 0.f;
                        srow[j] = p;
                    }
                }
            }
        }
    }
    __syncwarp();
    float acc[4][DHP / 32];
    long_zero<DHP>(acc);
    for (int j0 = 0; j0 < kend; j0 += kLongBlk) {
        const int nk = min(kLongBlk, Lk - j0);
        __syncthreads();
        stage_tile<T, DHP>(vb + (long long)j0 * D.v_ld, D.v_ld, nk, dh, Bs, KS);
        __syncthreads();
        if (r0 < r1) long_accumulate<DHP>(Ss + j0, LkP, r0, r1 - r0, Bs, KS, nk, lane, acc);
    }
    if (r0 < r1)
        long_store_rows<T, DHP>(o + (long long)b * D.o_bs + (long long)(q0 + r0) * D.o_ld + h * dh, D.o_ld, r1 - r0, dh, lane, acc);
}

template <typename T, int DHP>
__global__ void __launch_bounds__(kThreads)
attn_long_bwd_q_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ d_o,
                       T* __restrict__ dq, const unsigned char* __restrict__ key_pad, Dims D, float drop_p,
                       const unsigned long long* __restrict__ rng_state, unsigned int site, float* __restrict__ row_stats) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x, b = bh / D.H, h = bh % D.H;
    const int dh = D.dh, Lk = D.Lk, Lq = D.Lq;
    const int q0 = blockIdx.y * kLongRows, nq = min(kLongRows, Lq - q0);
    constexpr int KS = DHP + 1;
    const int LkP = (Lk + 3) & ~3;
    float* Qs = sm;                               // [16][DHP]
    float* dOs = Qs + kLongRows * DHP;            // [16][DHP]
    float* Bs = dOs + kLongRows * DHP;            // [32][KS]    streamed block of K / V
    float* Ss = Bs + kLongBlk * KS;               // [16][LkP]   scores -> P -> dS
    float* dPs = Ss + kLongRows * LkP;            // [16][LkP]   dO . V -> dropped dP
    const Rng rng = make_rng(rng_state, drop_p);
    const int r0 = warp * 4, r1 = min(nq, r0 + 4);
    const int kend = D.causal ? min(Lk, q0 + nq) : Lk;
    const T* kb = k + (long long)b * D.k_bs + h * dh;
    const T* vb = v + (long long)b * D.v_bs + h * dh;
    stage_tile<T, DHP>(q + (long long)b * D.q_bs + (long long)q0 * D.q_ld + h * dh, D.q_ld, nq, dh, Qs, DHP);
    stage_tile<T, DHP>(d_o + (long long)b * D.do_bs + (long long)q0 * D.do_ld + h * dh, D.do_ld, nq, dh, dOs, DHP);
    for (int j0 = 0; j0 < kend; j0 += kLongBlk) {
        const int nk = min(kLongBlk, Lk - j0);
        __syncthreads();
        stage_tile<T, DHP>(kb + (long long)j0 * D.k_ld, D.k_ld, nk, dh, Bs, KS);
        __syncthreads();
        if (r0 < r1) rows_dot_keys<DHP>(Qs, DHP, Bs, KS, nk, LkP, r0, r1, lane, Ss + j0);
        __syncthreads();
        stage_tile<T, DHP>(vb + (long long)j0 * D.v_ld, D.v_ld, nk, dh, Bs, KS);
        __syncthreads();
        if (r0 < r1) rows_dot_keys<DHP>(dOs, DHP, Bs, KS, nk, LkP, r0, r1, lane, dPs + j0);
    }
    __syncwarp();
    const unsigned char* pad_row = key_pad ? key_pad + (long long)b * Lk : nullptr;
    const int nG = (Lk + 7) >> 3;
    for (int il = r0; il < r1; ++il) {
        const int gi = q0 + il;
        float* srow = Ss + il * LkP;
        float* drow = dPs + il * LkP;
        float m, inv;
        long_softmax_row(srow, Lk, gi, lane, D.causal != 0, pad_row, D.scale, m, inv);
        const long long row = (long long)bh * Lq + gi;
        float dsum = 0.f;
        for (int c0 = 0; c0 < Lk; c0 += 256) {
            const uint32_t bits = rng.p > 0.f ? long_group_bits(rng, site, row, nG, (c0 >> 3) + lane) : 0xFFu;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (c0 + 32 * t < Lk) {
                    const uint32_t gb = __shfl_sync(0xffffffffu, bits, 4 * t + (lane >> 3));
                    const int j = c0 + 32 * t + lane;
                    if (j < Lk) {
                        const float p = srow[j] * inv;
                        const float sc = rng.p > 0.f ? (((gb >> (lane & 7)) & 1u) ? rng.inv_keep : 0.f) : 1.f;
                        const float dP = p != 0.f ? drow[j] * sc : 0.f;     // (masked columns may hold unwritten memory)
                        dsum += p * dP;
                        srow[j] = p;
                        drow[j] = dP;
                    }
                }
            }
        }
        dsum = warp_sum(dsum);
        for (int j = lane; j < Lk; j += 32) srow[j] = srow[j] * (drow[j] - dsum) * D.scale;
        if (lane == 0) *reinterpret_cast<float4*>(row_stats + row * 4) = make_float4(m, inv, dsum, 0.f);
    }
    __syncwarp();
    float acc[4][DHP / 32];
    long_zero<DHP>(acc);
    for (int j0 = 0; j0 < kend; j0 += kLongBlk) {
        const int nk = min(kLongBlk, Lk - j0);
        __syncthreads();
        stage_tile<T, DHP>(kb + (long long)j0 * D.k_ld, D.k_ld, nk, dh, Bs, KS);
        __syncthreads();
        if (r0 < r1) long_accumulate<DHP>(Ss + j0, LkP, r0, r1 - r0, Bs, KS, nk, lane, acc);
    }
    if (r0 < r1)
        long_store_rows<T, DHP>(dq + (long long)b * D.dq_bs + (long long)(q0 + r0) * D.dq_ld + h * dh, D.dq_ld, r1 - r0, dh, lane, acc);
}

template <typename T, int DHP>
__global__ void __launch_bounds__(kThreads)
attn_long_bwd_kv_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ d_o,
                        T* __restrict__ dk, T* __restrict__ dv, const unsigned char* __restrict__ key_pad, Dims D,
                        float drop_p, const unsigned long long* __restrict__ rng_state, unsigned int site,
                        const float* __restrict__ row_stats) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x, b = bh / D.H, h = bh % D.H;
    const int dh = D.dh, Lk = D.Lk, Lq = D.Lq;
    const int k0 = blockIdx.y * kLongRows, nkk = min(kLongRows, Lk - k0);
    constexpr int KS = DHP + 1;
    const int LqP = (Lq + 3) & ~3;
    float* Kt = sm;                               // [16][DHP]   the tile's keys
    float* Vt = Kt + kLongRows * DHP;             // [16][DHP]   ... and values
    float* Bs = Vt + kLongRows * DHP;             // [32][KS]    streamed block of Q / dO
    float* St = Bs + kLongBlk * KS;               // [16][LqP]   S^T -> dS^T
    float* Pt = St + kLongRows * LqP;             // [16][LqP]   dP^T -> dropped P^T
    const Rng rng = make_rng(rng_state, drop_p);
    const int r0 = warp * 4, r1 = min(nkk, r0 + 4);                 // this warp's keys of the tile
    const int i_begin = D.causal ? (k0 & ~(kLongBlk - 1)) : 0;      // causal: queries i < k0 see none of these keys
    const T* qb = q + (long long)b * D.q_bs + h * dh;
    const T* gb_ = d_o + (long long)b * D.do_bs + h * dh;
    stage_tile<T, DHP>(k + (long long)b * D.k_bs + (long long)k0 * D.k_ld + h * dh, D.k_ld, nkk, dh, Kt, DHP);
    stage_tile<T, DHP>(v + (long long)b * D.v_bs + (long long)k0 * D.v_ld + h * dh, D.v_ld, nkk, dh, Vt, DHP);
    for (int i0 = i_begin; i0 < Lq; i0 += kLongBlk) {
        const int ni = min(kLongBlk, Lq - i0);
        __syncthreads();
        stage_tile<T, DHP>(qb + (long long)i0 * D.q_ld, D.q_ld, ni, dh, Bs, KS);
        __syncthreads();
        if (r0 < r1) rows_dot_keys<DHP>(Kt, DHP, Bs, KS, ni, LqP, r0, r1, lane, St + i0);
        __syncthreads();
        stage_tile<T, DHP>(gb_ + (long long)i0 * D.do_ld, D.do_ld, ni, dh, Bs, KS);
        __syncthreads();
        if (r0 < r1) rows_dot_keys<DHP>(Vt, DHP, Bs, KS, ni, LqP, r0, r1, lane, Pt + i0);
    }
    __syncwarp();
    if (r0 < r1) {
        const unsigned char* pad_row = key_pad ? key_pad + (long long)b * Lk : nullptr;
        const int nG = (Lk + 7) >> 3;
        const int g = (k0 + r0) >> 3, bit0 = (k0 + r0) & 7;        // the warp's 4 keys sit in ONE group of 8 (bit0 = 0 or 4)
        for (int i = i_begin + lane; i < Lq; i += 32) {
            const long long row = (long long)bh * Lq + i;
            const float4 st = *reinterpret_cast<const float4*>(row_stats + row * 4);      // row max, 1 / row sum, sum_j P dP
            const uint32_t bits = rng.p > 0.f ? long_group_bits(rng, site, row, nG, g) : 0xFFu;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int jl = r0 + kk;
                if (jl < r1) {
                    const int j = k0 + jl;
                    const bool ok = !(D.causal && j > i) && !(pad_row != nullptr && pad_row[j]);
                    float p = 0.f;
                    if (ok && st.x != -INFINITY) p = expf(St[jl * LqP + i] * D.scale - st.x) * st.y;
                    const float sc = rng.p > 0.f ? (((bits >> (bit0 + kk)) & 1u) ? rng.inv_keep : 0.f) : 1.f;
                    const float dP = p != 0.f ? Pt[jl * LqP + i] * sc : 0.f;
                    St[jl * LqP + i] = p * (dP - st.z) * D.scale;
                    Pt[jl * LqP + i] = p * sc;
                }
            }
        }
    }
    __syncwarp();
    float ak[4][DHP / 32], av[4][DHP / 32];
    long_zero<DHP>(ak);
    long_zero<DHP>(av);
    for (int i0 = i_begin; i0 < Lq; i0 += kLongBlk) {
        const int ni = min(kLongBlk, Lq - i0);
        __syncthreads();
        stage_tile<T, DHP>(qb + (long long)i0 * D.q_ld, D.q_ld, ni, dh, Bs, KS);
        __syncthreads();
        if (r0 < r1) long_accumulate<DHP>(St + i0, LqP, r0, r1 - r0, Bs, KS, ni, lane, ak);
        __syncthreads();
        stage_tile<T, DHP>(gb_ + (long long)i0 * D.do_ld, D.do_ld, ni, dh, Bs, KS);
        __syncthreads();
        if (r0 < r1) long_accumulate<DHP>(Pt + i0, LqP, r0, r1 - r0, Bs, KS, ni, lane, av);
    }
    if (r0 < r1) {
        long_store_rows<T, DHP>(dk + (long long)b * D.dk_bs + (long long)(k0 + r0) * D.dk_ld + h * dh, D.dk_ld, r1 - r0, dh, lane, ak);
        long_store_rows<T, DHP>(dv + (long long)b * D.dv_bs + (long long)(k0 + r0) * D.dv_ld + h * dh, D.dv_ld, r1 - r0, dh, lane, av);
    }
}

int validate(const vct_attn_args* a, const char* who) {
    VCT_REQUIRE(a != nullptr, "%s: null args", who);
    VCT_REQUIRE(a->B > 0 && a->H > 0 && a->Lq > 0 && a->Lk > 0, "%s: empty problem", who);
    VCT_REQUIRE(a->Lq <= kLongMax && a->Lk <= kLongMax, "%s: Lq, Lk must be <= %d (got %d, %d)", who, kLongMax, a->Lq, a->Lk);
    VCT_REQUIRE(a->dh % 4 == 0 && a->dh <= 128, "%s: dh must be a multiple of 4 and <= 128 (got %d)", who, a->dh);
    VCT_REQUIRE(!a->causal || a->Lq == a->Lk, "%s: causal needs Lq == Lk", who);
    VCT_REQUIRE(a->q_ld % 4 == 0 && a->k_ld % 4 == 0 && a->v_ld % 4 == 0 && a->o_ld % 4 == 0,
                "%s: row strides must be multiples of 4 elements", who);
    VCT_REQUIRE(a->dtype == VCT_F32 || a->dtype == VCT_BF16, "%s: bad dtype", who);
    return 0;
}

Dims make_dims(const vct_attn_args* a) {
    Dims D;
    D.B = a->B; D.H = a->H; D.Lq = a->Lq; D.Lk = a->Lk; D.dh = a->dh;
    D.q_ld = a->q_ld; D.k_ld = a->k_ld; D.v_ld = a->v_ld; D.o_ld = a->o_ld;
    D.do_ld = a->do_ld; D.dq_ld = a->dq_ld; D.dk_ld = a->dk_ld; D.dv_ld = a->dv_ld;
    D.q_bs = a->q_bs ? a->q_bs : (long long)a->Lq * a->q_ld;
    D.k_bs = a->k_bs ? a->k_bs : (long long)a->Lk * a->k_ld;
    D.v_bs = a->v_bs ? a->v_bs : (long long)a->Lk * a->v_ld;
    D.o_bs = a->o_bs ? a->o_bs : (long long)a->Lq * a->o_ld;
    D.do_bs = a->do_bs ? a->do_bs : (long long)a->Lq * a->do_ld;
    D.dq_bs = a->dq_bs ? a->dq_bs : (long long)a->Lq * a->dq_ld;
    D.dk_bs = a->dk_bs ? a->dk_bs : (long long)a->Lk * a->dk_ld;
    D.dv_bs = a->dv_bs ? a->dv_bs : (long long)a->Lk * a->dv_ld;
    D.causal = a->causal; D.scale = a->scale;
    return D;
}

template <typename T, int DHP>
int launch_fwd(const vct_attn_args* a, cudaStream_t st) {
    const int LkP = (a->Lk + 3) & ~3, LqP = (a->Lq + 3) & ~3;
    const size_t smem = sizeof(float) * ((size_t)a->Lq * DHP + 2 * (size_t)a->Lk * (DHP + 1) + (size_t)a->Lq * LkP + (size_t)a->Lk * LqP);
    VCT_REQUIRE(smem <= kSmemBudget, "vct_attn_fwd: head does not fit shared memory");
    auto kern = attn_fwd_kernel<T, DHP>;
    static bool once = false;
    if (!once) { VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)); once = true; }
    vct::launch(kern, dim3(a->B * a->H), dim3(kThreads), smem, st, (const T*)a->q, (const T*)a->k, (const T*)a->v, (T*)a->o, a->key_pad, a->probs,
                                              make_dims(a), a->drop_p, a->rng_state, a->site);
    return check_launch("vct_attn_fwd");
}

template <typename T, int DHP>
int launch_bwd(const vct_attn_args* a, cudaStream_t st) {
    const int LkP = (a->Lk + 3) & ~3, LqP = (a->Lq + 3) & ~3;
    const size_t smem = sizeof(float) * (2 * (size_t)a->Lq * DHP + 2 * (size_t)a->Lk * (DHP + 1) + 2 * (size_t)a->Lq * LkP + 2 * (size_t)a->Lk * LqP);
    VCT_REQUIRE(smem <= kSmemBudget, "vct_attn_bwd: head does not fit shared memory");
    auto kern = attn_bwd_kernel<T, DHP>;
    static bool once = false;
    if (!once) { VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)); once = true; }
    vct::launch(kern, dim3(a->B * a->H), dim3(kThreads), smem, st, (const T*)a->q, (const T*)a->k, (const T*)a->v, (const T*)a->d_o, (T*)a->dq,
                                              (T*)a->dk, (T*)a->dv, a->key_pad, make_dims(a), a->drop_p, a->rng_state, a->site,
                                              a->dbias, a->dbias_partials, a->dbias_counters);
    return check_launch("vct_attn_bwd");
}


// ---------------------------------------------------------------------------------------------------------------------
// Decode-step attention (Lq = 1: the new token of greedy decoding attends over the K/V cache or the projected memory,
// model/CapDecoder.py:62-79 with a K/V cache).  One WARP per (batch, head), nothing staged in shared memory:
//   scores : lane = key; the lane streams its K row (dh contiguous elements, 16-byte loads) against the query row, which
//            every lane holds in registers
//   softmax: warp shuffles over the key lanes (keys lane, lane + 32)
//   values : lane = 4 consecutive head-dim columns (dh / 4 <= 32 lanes); probabilities broadcast by shuffle, V rows read
//            coalesced
// The general kernel above spends a whole 4-warp CTA (and a shared-memory round trip) on the one query row: 18 us for
// B = 256, H = 8 against ~3 us here.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDecWarps = 8;

template <typename T, int DHP>
__global__ void __launch_bounds__(kDecWarps * 32)
attn_decode_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ o,
                   const unsigned char* __restrict__ key_pad, float* __restrict__ probs, Dims D) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(16) float qs[kDecWarps][DHP];            // the warp's query row (read as broadcast float4)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x * kDecWarps + warp;
    if (bh >= D.B * D.H) return;
    const int b = bh / D.H, h = bh % D.H;
    const int dh = D.dh, Lk = D.Lk;
    {
        const T* qrow = q + (long long)b * D.q_bs + h * dh;
        for (int c = 4 * lane; c < DHP; c += 128) {
            const float4 t = c < dh ? ld4(qrow + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(&qs[warp][c]) = t;
        }
    }
    __syncwarp();
    const unsigned char* pad_row = key_pad ? key_pad + (long long)b * Lk : nullptr;
    float s[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int j = lane + 32 * t;
        s[t] = -INFINITY;
        if (j < Lk && !(pad_row != nullptr && pad_row[j])) {
            const T* krow = k + (long long)b * D.k_bs + (long long)j * D.k_ld + h * dh;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int c8 = 0; c8 < DHP / 8; ++c8) {
                if (8 * c8 < dh) {                               // dh % 8 == 0 on this path (checked by the launcher)
                    float kv[8];
                    ld8(krow + 8 * c8, kv);
                    const float4 q0 = *reinterpret_cast<const float4*>(&qs[warp][8 * c8]);
                    const float4 q1 = *reinterpret_cast<const float4*>(&qs[warp][8 * c8 + 4]);
                    a0 = fmaf(q0.x, kv[0], a0); a1 = fmaf(q0.y, kv[1], a1); a2 = fmaf(q0.z, kv[2], a2); a3 = fmaf(q0.w, kv[3], a3);
                    a0 = fmaf(q1.x, kv[4], a0); a1 = fmaf(q1.y, kv[5], a1); a2 = fmaf(q1.z, kv[6], a2); a3 = fmaf(q1.w, kv[7], a3);
                }
            }
            s[t] = ((a0 + a1) + (a2 + a3)) * D.scale;
        }
    }
    const float m = warp_max(fmaxf(s[0], s[1]));
    const float e0 = s[0] == -INFINITY ? 0.f : expf(s[0] - m);
    const float e1 = s[1] == -INFINITY ? 0.f : expf(s[1] - m);
    const float den = warp_sum(e0 + e1);
    const float inv = den > 0.f ? 1.f / den : 0.f;
    const float p0 = e0 * inv, p1 = e1 * inv;
    if (probs) {
        float* pr = probs + (long long)bh * Lk;          // [B, H, 1, Lk]
        if (lane < Lk) pr[lane] = p0;
        if (lane + 32 < Lk) pr[lane + 32] = p1;
    }
    // o[c] = sum_j p_j V[j][c]; lane owns columns 4 lane .. 4 lane + 3; four V rows in flight per pass
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool colok = 4 * lane < dh;
    const T* vbase = v + (long long)b * D.v_bs + h * dh + 4 * lane;
    for (int j0 = 0; j0 < Lk; j0 += 4) {
        float4 vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            vv[u] = (colok && j0 + u < Lk) ? ld4(vbase + (long long)(j0 + u) * D.v_ld) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            const float pj = __shfl_sync(0xffffffffu, j < 32 ? p0 : p1, j & 31);
            acc.x = fmaf(pj, vv[u].x, acc.x); acc.y = fmaf(pj, vv[u].y, acc.y);
            acc.z = fmaf(pj, vv[u].z, acc.z); acc.w = fmaf(pj, vv[u].w, acc.w);
        }
    }
    if (colok) st4(o + (long long)b * D.o_bs + h * dh + 4 * lane, acc);
}

template <typename T, int DHP>
int launch_decode(const vct_attn_args* a, cudaStream_t st) {
    const int warps = a->B * a->H;
    vct::launch(attn_decode_kernel<T, DHP>, dim3((warps + kDecWarps - 1) / kDecWarps), dim3(kDecWarps * 32), 0, st, (const T*)a->q,
                (const T*)a->k, (const T*)a->v, (T*)a->o, a->key_pad, a->probs, make_dims(a));
    return check_launch("vct_attn_fwd(decode)");
}

// ---- long sequences: forward = one launch, backward = the query-tile kernel, then the key-tile kernel ----
template <typename T, int DHP>
int launch_long_fwd(const vct_attn_args* a, cudaStream_t st) {
    const int LkP = (a->Lk + 3) & ~3;
    const size_t smem = sizeof(float) * ((size_t)kLongRows * DHP + (size_t)kLongBlk * (DHP + 1) + (size_t)kLongRows * LkP);
    VCT_REQUIRE(smem <= kSmemBudget, "vct_attn_fwd: key rows do not fit shared memory");
    auto kern = attn_long_fwd_kernel<T, DHP>;
    static bool once = false;
    if (!once) { VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)); once = true; }
    vct::launch(kern, dim3(a->B * a->H, (a->Lq + kLongRows - 1) / kLongRows), dim3(kThreads), smem, st, (const T*)a->q, (const T*)a->k,
                (const T*)a->v, (T*)a->o, a->key_pad, a->probs, make_dims(a), a->drop_p, a->rng_state, a->site);
    return check_launch("vct_attn_fwd(long)");
}

template <typename T, int DHP>
int launch_long_bwd(const vct_attn_args* a, cudaStream_t st) {
    VCT_REQUIRE(a->row_stats != nullptr && (reinterpret_cast<uintptr_t>(a->row_stats) & 15) == 0,
                "vct_attn_bwd: sequences longer than 64 need the 16-byte aligned row_stats workspace (fp32 [B*H*Lq, 4])");
    VCT_REQUIRE(a->dbias == nullptr, "vct_attn_bwd: dbias is produced in-kernel only for Lq, Lk <= 64 (use vct_colsum on dq | dk | dv)");
    const int LP = ((a->Lq > a->Lk ? a->Lq : a->Lk) + 3) & ~3;
    const size_t smem = sizeof(float) * (2 * (size_t)kLongRows * DHP + (size_t)kLongBlk * (DHP + 1) + 2 * (size_t)kLongRows * LP);
    VCT_REQUIRE(smem <= kSmemBudget, "vct_attn_bwd: score rows do not fit shared memory");
    auto kq = attn_long_bwd_q_kernel<T, DHP>;
    auto kkv = attn_long_bwd_kv_kernel<T, DHP>;
    static bool once = false;
    if (!once) {
        VCT_CUDA(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        VCT_CUDA(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        once = true;
    }
    const Dims D = make_dims(a);
    vct::launch(kq, dim3(a->B * a->H, (a->Lq + kLongRows - 1) / kLongRows), dim3(kThreads), smem, st, (const T*)a->q, (const T*)a->k,
                (const T*)a->v, (const T*)a->d_o, (T*)a->dq, a->key_pad, D, a->drop_p, a->rng_state, a->site, a->row_stats);
    if (int e = check_launch("vct_attn_bwd(long, q)")) return e;
    vct::launch(kkv, dim3(a->B * a->H, (a->Lk + kLongRows - 1) / kLongRows), dim3(kThreads), smem, st, (const T*)a->q, (const T*)a->k,
                (const T*)a->v, (const T*)a->d_o, (T*)a->dk, (T*)a->dv, a->key_pad, D, a->drop_p, a->rng_state, a->site,
                (const float*)a->row_stats);
    return check_launch("vct_attn_bwd(long, kv)");
}

template <typename T>
int dispatch(const vct_attn_args* a, cudaStream_t st, bool bwd) {
    const int dh = a->dh;
    if (a->Lq > 64 || a->Lk > 64) {
        if (dh <= 32) return bwd ? launch_long_bwd<T, 32>(a, st) : launch_long_fwd<T, 32>(a, st);
        if (dh <= 64) return bwd ? launch_long_bwd<T, 64>(a, st) : launch_long_fwd<T, 64>(a, st);
        if (dh <= 96) return bwd ? launch_long_bwd<T, 96>(a, st) : launch_long_fwd<T, 96>(a, st);
        return bwd ? launch_long_bwd<T, 128>(a, st) : launch_long_fwd<T, 128>(a, st);
    }
    // one query row, no dropout (greedy decoding): the warp-per-head kernel
    static const bool dec_on = [] { const char* e = getenv("VCT_ATTN_DECODE"); return e == nullptr || e[0] != '0'; }();
    if (!bwd && dec_on && a->Lq == 1 && !(a->drop_p > 0.f)) {
        const bool al = ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
                          reinterpret_cast<uintptr_t>(a->o)) & 15) == 0 && (a->q_bs % 4 == 0) && (a->k_bs % 4 == 0) && (a->v_bs % 4 == 0) &&
                        (a->o_bs % 8 == 0) && a->dh % 8 == 0 && a->q_bs % 8 == 0 && a->k_bs % 8 == 0 && a->k_ld % 8 == 0;
        if (al) {
            if (dh <= 32) return launch_decode<T, 32>(a, st);
            if (dh <= 64) return launch_decode<T, 64>(a, st);
            if (dh <= 96) return launch_decode<T, 96>(a, st);
            return launch_decode<T, 128>(a, st);
        }
    }
    if (dh <= 32) return bwd ? launch_bwd<T, 32>(a, st) : launch_fwd<T, 32>(a, st);
    if (dh <= 64) return bwd ? launch_bwd<T, 64>(a, st) : launch_fwd<T, 64>(a, st);
    if (dh <= 96) return bwd ? launch_bwd<T, 96>(a, st) : launch_fwd<T, 96>(a, st);
    return bwd ? launch_bwd<T, 128>(a, st) : launch_fwd<T, 128>(a, st);
}

}  // namespace

extern "C" int vct_attn_fwd(const vct_attn_args* a, vct_stream_t stream) {
    if (int e = validate(a, "vct_attn_fwd")) return e;
    VCT_REQUIRE(a->q && a->k && a->v && a->o, "vct_attn_fwd: null tensor");
    if (a->dtype == VCT_BF16) return dispatch<__nv_bfloat16>(a, (cudaStream_t)stream, false);
    return dispatch<float>(a, (cudaStream_t)stream, false);
}

namespace vct {
int attn_bwd_tc(const vct_attn_args* m, cudaStream_t st);   // attn_fused.cu
}

extern "C" int vct_attn_bwd(const vct_attn_args* a, vct_stream_t stream) {
    if (int e = validate(a, "vct_attn_bwd")) return e;
    VCT_REQUIRE(a->q && a->k && a->v && a->d_o && a->dq && a->dk && a->dv, "vct_attn_bwd: null tensor");
    VCT_REQUIRE(a->do_ld % 4 == 0 && a->dq_ld % 4 == 0 && a->dk_ld % 4 == 0 && a->dv_ld % 4 == 0,
                "vct_attn_bwd: gradient row strides must be multiples of 4 elements");
    VCT_REQUIRE(a->dbias == nullptr || (a->dbias_partials != nullptr && a->dbias_counters != nullptr),
                "vct_attn_bwd: dbias needs dbias_partials and dbias_counters");
    // bf16, sequences up to 32 queries / 64 keys: every contraction on tcgen05 (attn_fused.cu); otherwise the SIMT kernel
    {
        const int r = vct::attn_bwd_tc(a, (cudaStream_t)stream);
        if (r <= 0) return r;
    }
    if (a->dtype == VCT_BF16) return dispatch<__nv_bfloat16>(a, (cudaStream_t)stream, true);
    return dispatch<float>(a, (cudaStream_t)stream, true);
}
