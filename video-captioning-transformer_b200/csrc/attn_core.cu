// Attention core: softmax(scale * Q K^T + mask) V, forward and backward.
//
// One CTA (4 warps) per (batch, head).  Sequence lengths on this path are 13..33, so the
// whole head -- Q, K, V (and dO) -- is staged once in shared memory as fp32; each warp owns a contiguous block
// of query rows.
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

int validate(const vct_attn_args* a, const char* who) {
    VCT_REQUIRE(a != nullptr, "%s: null args", who);
    VCT_REQUIRE(a->B > 0 && a->H > 0 && a->Lq > 0 && a->Lk > 0, "%s: empty problem", who);
    VCT_REQUIRE(a->Lq <= 64 && a->Lk <= 64, "%s: Lq, Lk must be <= 64 (got %d, %d)", who, a->Lq, a->Lk);
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

template <typename T>
int dispatch(const vct_attn_args* a, cudaStream_t st, bool bwd) {
    const int dh = a->dh;
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
