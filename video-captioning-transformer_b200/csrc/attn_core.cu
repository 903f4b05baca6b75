// Attention core: softmax(scale * Q K^T + mask) V, forward and backward.
// One warp per (batch, head); K/V of the head live in shared memory (fp32), scores and the
// softmax are spread one key per lane (two when Lk > 32) and reduced with warp shuffles.
// Sequence lengths on this path are 13..33, so a whole head fits a warp.
#include "common.cuh"

using namespace vct;

namespace {

constexpr int kMaxWarps = 4;
constexpr size_t kSmemBudget = 200 * 1024;

struct Dims {
    int B, H, Lq, Lk, dh;
    long long q_ld, k_ld, v_ld, o_ld, do_ld, dq_ld, dk_ld, dv_ld;
    long long q_bs, k_bs, v_bs, o_bs, do_bs, dq_bs, dk_bs, dv_bs;   // batch strides (elements)
    int causal;
    float scale;
};

// scores of query row i against key slots (lane, lane+32); returns probabilities p0,p1 (0 where masked)
__device__ __forceinline__ void row_softmax(const float* __restrict__ qs, const float* __restrict__ Ks, int dh, int Lk,
                                            int i, int lane, bool causal, const unsigned char* __restrict__ pad_row,
                                            float scale, float& p0, float& p1) {
    float s[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int j = lane + 32 * t;
        float a = -INFINITY;
        if (j < Lk && !(causal && j > i) && !(pad_row != nullptr && pad_row[j])) {
            const float* kr = Ks + j * (dh + 1);
            a = 0.f;
            for (int c = 0; c < dh; ++c) a += qs[c] * kr[c];
            a *= scale;
        }
        s[t] = a;
    }
    const float m = warp_max(fmaxf(s[0], s[1]));
    if (m == -INFINITY) { p0 = p1 = 0.f; return; }   // fully masked row (cannot happen on this path, Q8)
    const float e0 = s[0] == -INFINITY ? 0.f : expf(s[0] - m);
    const float e1 = s[1] == -INFINITY ? 0.f : expf(s[1] - m);
    const float inv = 1.f / warp_sum(e0 + e1);
    p0 = e0 * inv;
    p1 = e1 * inv;
}

template <typename T>
__global__ void __launch_bounds__(kMaxWarps * 32)
attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ o,
                const unsigned char* __restrict__ key_pad, float* __restrict__ probs, Dims D, float drop_p,
                const unsigned long long* __restrict__ rng_state, unsigned int site, int per_warp_floats) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x * (blockDim.x >> 5) + warp;
    if (bh >= D.B * D.H) return;
    const int b = bh / D.H, h = bh % D.H;
    const int dh = D.dh, Lk = D.Lk, Lq = D.Lq, nv = dh >> 2;
    float* Ks = smem + (size_t)warp * per_warp_floats;   // [Lk][dh+1]
    float* Vs = Ks + Lk * (dh + 1);                      // [Lk][dh]
    float* qs = Vs + Lk * dh;                            // [dh]
    const Rng rng = make_rng(rng_state, drop_p);

    for (int idx = lane; idx < Lk * nv; idx += 32) {
        const int j = idx / nv, c = (idx % nv) * 4;
        float4 kv = ld4(k + (long long)b * D.k_bs + (long long)j * D.k_ld + h * dh + c);
        float4 vv = ld4(v + (long long)b * D.v_bs + (long long)j * D.v_ld + h * dh + c);
        float* kd = Ks + j * (dh + 1) + c;
        kd[0] = kv.x; kd[1] = kv.y; kd[2] = kv.z; kd[3] = kv.w;
        float* vd = Vs + j * dh + c;
        vd[0] = vv.x; vd[1] = vv.y; vd[2] = vv.z; vd[3] = vv.w;
    }
    const unsigned char* pad_row = key_pad ? key_pad + (long long)b * Lk : nullptr;
    for (int i = 0; i < Lq; ++i) {
        __syncwarp();
        if (lane < nv) {
            float4 qv = ld4(q + (long long)b * D.q_bs + (long long)i * D.q_ld + h * dh + lane * 4);
            qs[lane * 4 + 0] = qv.x; qs[lane * 4 + 1] = qv.y; qs[lane * 4 + 2] = qv.z; qs[lane * 4 + 3] = qv.w;
        }
        __syncwarp();
        float p0, p1;
        row_softmax(qs, Ks, dh, Lk, i, lane, D.causal != 0, pad_row, D.scale, p0, p1);
        const long long pbase = ((long long)bh * Lq + i) * Lk;
        if (probs) {
            if (lane < Lk) probs[pbase + lane] = p0;
            if (lane + 32 < Lk) probs[pbase + lane + 32] = p1;
        }
        if (rng.p > 0.f) {
            if (lane < Lk) p0 *= dropout_scale1(rng, site, (unsigned long long)(pbase + lane));
            if (lane + 32 < Lk) p1 *= dropout_scale1(rng, site, (unsigned long long)(pbase + lane + 32));
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < Lk; ++j) {
            const float pj = __shfl_sync(0xffffffffu, j < 32 ? p0 : p1, j & 31);
            const float* vr = Vs + j * dh;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = lane + 32 * cc;
                if (c < dh) acc[cc] += pj * vr[c];
            }
        }
        T* orow = o + (long long)b * D.o_bs + (long long)i * D.o_ld + h * dh;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = lane + 32 * cc;
            if (c < dh) orow[c] = from_f32<T>(acc[cc]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kMaxWarps * 32)
attn_bwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ d_o,
                T* __restrict__ dq, T* __restrict__ dk, T* __restrict__ dv, const unsigned char* __restrict__ key_pad,
                Dims D, float drop_p, const unsigned long long* __restrict__ rng_state, unsigned int site,
                int per_warp_floats) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x * (blockDim.x >> 5) + warp;
    if (bh >= D.B * D.H) return;
    const int b = bh / D.H, h = bh % D.H;
    const int dh = D.dh, Lk = D.Lk, Lq = D.Lq, nv = dh >> 2;
    float* Ks = smem + (size_t)warp * per_warp_floats;   // [Lk][dh+1]
    float* Vs = Ks + Lk * (dh + 1);                      // [Lk][dh+1]
    float* dKs = Vs + Lk * (dh + 1);                     // [Lk][dh]
    float* dVs = dKs + Lk * dh;                          // [Lk][dh]
    float* qs = dVs + Lk * dh;                           // [dh]
    float* dos = qs + dh;                                // [dh]
    const Rng rng = make_rng(rng_state, drop_p);

    for (int idx = lane; idx < Lk * nv; idx += 32) {
        const int j = idx / nv, c = (idx % nv) * 4;
        float4 kv = ld4(k + (long long)b * D.k_bs + (long long)j * D.k_ld + h * dh + c);
        float4 vv = ld4(v + (long long)b * D.v_bs + (long long)j * D.v_ld + h * dh + c);
        float* kd = Ks + j * (dh + 1) + c;
        kd[0] = kv.x; kd[1] = kv.y; kd[2] = kv.z; kd[3] = kv.w;
        float* vd = Vs + j * (dh + 1) + c;
        vd[0] = vv.x; vd[1] = vv.y; vd[2] = vv.z; vd[3] = vv.w;
    }
    for (int idx = lane; idx < Lk * dh; idx += 32) { dKs[idx] = 0.f; dVs[idx] = 0.f; }
    const unsigned char* pad_row = key_pad ? key_pad + (long long)b * Lk : nullptr;

    for (int i = 0; i < Lq; ++i) {
        __syncwarp();
        if (lane < nv) {
            float4 qv = ld4(q + (long long)b * D.q_bs + (long long)i * D.q_ld + h * dh + lane * 4);
            qs[lane * 4 + 0] = qv.x; qs[lane * 4 + 1] = qv.y; qs[lane * 4 + 2] = qv.z; qs[lane * 4 + 3] = qv.w;
            float4 gv = ld4(d_o + (long long)b * D.do_bs + (long long)i * D.do_ld + h * dh + lane * 4);
            dos[lane * 4 + 0] = gv.x; dos[lane * 4 + 1] = gv.y; dos[lane * 4 + 2] = gv.z; dos[lane * 4 + 3] = gv.w;
        }
        __syncwarp();
        float p[2];
        row_softmax(qs, Ks, dh, Lk, i, lane, D.causal != 0, pad_row, D.scale, p[0], p[1]);
        const long long pbase = ((long long)bh * Lq + i) * Lk;
        float sc[2] = {1.f, 1.f}, dP[2] = {0.f, 0.f};
        float dsum = 0.f;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int j = lane + 32 * t;
            if (j < Lk) {
                if (rng.p > 0.f) sc[t] = dropout_scale1(rng, site, (unsigned long long)(pbase + j));
                const float* vr = Vs + j * (dh + 1);
                float a = 0.f;
                for (int c = 0; c < dh; ++c) a += dos[c] * vr[c];
                dP[t] = a * sc[t];
                dsum += p[t] * dP[t];
            }
        }
        dsum = warp_sum(dsum);
        const float dS0 = p[0] * (dP[0] - dsum) * D.scale, dS1 = p[1] * (dP[1] - dsum) * D.scale;
        const float pd0 = p[0] * sc[0], pd1 = p[1] * sc[1];
        float qreg[4], doreg[4], dqacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = lane + 32 * cc;
            qreg[cc] = c < dh ? qs[c] : 0.f;
            doreg[cc] = c < dh ? dos[c] : 0.f;
        }
        for (int j = 0; j < Lk; ++j) {
            const float dsj = __shfl_sync(0xffffffffu, j < 32 ? dS0 : dS1, j & 31);
            const float pdj = __shfl_sync(0xffffffffu, j < 32 ? pd0 : pd1, j & 31);
            const float* kr = Ks + j * (dh + 1);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = lane + 32 * cc;
                if (c < dh) {
                    dqacc[cc] += dsj * kr[c];
                    dKs[j * dh + c] += dsj * qreg[cc];
                    dVs[j * dh + c] += pdj * doreg[cc];
                }
            }
        }
        T* dqrow = dq + (long long)b * D.dq_bs + (long long)i * D.dq_ld + h * dh;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = lane + 32 * cc;
            if (c < dh) dqrow[c] = from_f32<T>(dqacc[cc]);
        }
    }
    __syncwarp();
    for (int idx = lane; idx < Lk * nv; idx += 32) {
        const int j = idx / nv, c = (idx % nv) * 4;
        const float* a = dKs + j * dh + c;
        const float* g = dVs + j * dh + c;
        st4(dk + (long long)b * D.dk_bs + (long long)j * D.dk_ld + h * dh + c, make_float4(a[0], a[1], a[2], a[3]));
        st4(dv + (long long)b * D.dv_bs + (long long)j * D.dv_ld + h * dh + c, make_float4(g[0], g[1], g[2], g[3]));
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

}  // namespace

extern "C" int vct_attn_fwd(const vct_attn_args* a, vct_stream_t stream) {
    if (int e = validate(a, "vct_attn_fwd")) return e;
    VCT_REQUIRE(a->q && a->k && a->v && a->o, "vct_attn_fwd: null tensor");
    const int per_warp = a->Lk * (a->dh + 1) + a->Lk * a->dh + a->dh;
    int warps = (int)(kSmemBudget / ((size_t)per_warp * sizeof(float)));
    warps = warps > kMaxWarps ? kMaxWarps : warps;
    VCT_REQUIRE(warps >= 1, "vct_attn_fwd: head does not fit shared memory");
    const size_t smem = (size_t)warps * per_warp * sizeof(float);
    const int blocks = (a->B * a->H + warps - 1) / warps;
    const Dims D = make_dims(a);
    cudaStream_t st = (cudaStream_t)stream;
    if (a->dtype == VCT_BF16) {
        auto kern = attn_fwd_kernel<__nv_bfloat16>;
        static bool once = false;
        if (!once) { VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)); once = true; }
        kern<<<blocks, warps * 32, smem, st>>>((const __nv_bfloat16*)a->q, (const __nv_bfloat16*)a->k,
                                               (const __nv_bfloat16*)a->v, (__nv_bfloat16*)a->o, a->key_pad, a->probs, D,
                                               a->drop_p, a->rng_state, a->site, per_warp);
    } else {
        auto kern = attn_fwd_kernel<float>;
        static bool once = false;
        if (!once) { VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)); once = true; }
        kern<<<blocks, warps * 32, smem, st>>>((const float*)a->q, (const float*)a->k, (const float*)a->v, (float*)a->o,
                                               a->key_pad, a->probs, D, a->drop_p, a->rng_state, a->site, per_warp);
    }
    return check_launch("vct_attn_fwd");
}

extern "C" int vct_attn_bwd(const vct_attn_args* a, vct_stream_t stream) {
    if (int e = validate(a, "vct_attn_bwd")) return e;
    VCT_REQUIRE(a->q && a->k && a->v && a->d_o && a->dq && a->dk && a->dv, "vct_attn_bwd: null tensor");
    VCT_REQUIRE(a->do_ld % 4 == 0 && a->dq_ld % 4 == 0 && a->dk_ld % 4 == 0 && a->dv_ld % 4 == 0,
                "vct_attn_bwd: gradient row strides must be multiples of 4 elements");
    const int per_warp = 2 * a->Lk * (a->dh + 1) + 2 * a->Lk * a->dh + 2 * a->dh;
    int warps = (int)(kSmemBudget / ((size_t)per_warp * sizeof(float)));
    warps = warps > kMaxWarps ? kMaxWarps : warps;
    VCT_REQUIRE(warps >= 1, "vct_attn_bwd: head does not fit shared memory");
    const size_t smem = (size_t)warps * per_warp * sizeof(float);
    const int blocks = (a->B * a->H + warps - 1) / warps;
    const Dims D = make_dims(a);
    cudaStream_t st = (cudaStream_t)stream;
    if (a->dtype == VCT_BF16) {
        auto kern = attn_bwd_kernel<__nv_bfloat16>;
        static bool once = false;
        if (!once) { VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)); once = true; }
        kern<<<blocks, warps * 32, smem, st>>>((const __nv_bfloat16*)a->q, (const __nv_bfloat16*)a->k,
                                               (const __nv_bfloat16*)a->v, (const __nv_bfloat16*)a->d_o,
                                               (__nv_bfloat16*)a->dq, (__nv_bfloat16*)a->dk, (__nv_bfloat16*)a->dv,
                                               a->key_pad, D, a->drop_p, a->rng_state, a->site, per_warp);
    } else {
        auto kern = attn_bwd_kernel<float>;
        static bool once = false;
        if (!once) { VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)); once = true; }
        kern<<<blocks, warps * 32, smem, st>>>((const float*)a->q, (const float*)a->k, (const float*)a->v,
                                               (const float*)a->d_o, (float*)a->dq, (float*)a->dk, (float*)a->dv,
                                               a->key_pad, D, a->drop_p, a->rng_state, a->site, per_warp);
    }
    return check_launch("vct_attn_bwd");
}
