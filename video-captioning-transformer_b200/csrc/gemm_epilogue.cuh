// Fused GEMM epilogue shared by the SIMT and the tcgen05 kernels.
// Order: + bias[n]; + row_table[m % period, n]; activation; + addend[m, n]; store C (and C2).
// The activation kind is a template parameter so that the plain epilogue (most launches) carries neither the
// Philox nor the erf code: instruction-cache footprint matters for the many ~5 us GEMMs of a step.
#pragma once
#include "common.cuh"

namespace vct {

struct Epilogue {
    int M, N;
    void* C; int c_dtype; long long ldc;
    void* C2; int c2_dtype; long long ldc2;
    const float* bias;
    const float* row_table; int row_period;
    const float* addend; long long ld_addend;
    int act;
    const void* aux; int aux_dtype; long long ld_aux;
    float drop_p; const unsigned long long* rng_state; unsigned int site;
    int vec_ok;   // leading dimensions / pointers allow 16-byte loads and stores for full groups of 4 columns
    long long* trace;   // debug: CTA (0,0,0) writes clock64 timestamps of its pipeline phases here (NULL = off)
};

inline Epilogue make_epilogue(const vct_gemm_args* a) {
    Epilogue e;
    e.M = a->M; e.N = a->N;
    e.C = a->C; e.c_dtype = a->c_dtype; e.ldc = a->ldc;
    e.C2 = a->C2; e.c2_dtype = a->c2_dtype; e.ldc2 = a->ldc2;
    e.bias = a->bias;
    e.row_table = a->row_table; e.row_period = a->row_period;
    e.addend = a->addend; e.ld_addend = a->ld_addend;
    e.act = a->act;
    e.aux = a->aux; e.aux_dtype = a->aux_dtype; e.ld_aux = a->ld_aux;
    e.drop_p = a->drop_p; e.rng_state = a->rng_state; e.site = a->site;
    e.trace = nullptr;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    e.vec_ok = (a->ldc % 4 == 0) && (a->C2 == nullptr || a->ldc2 % 4 == 0) &&
               (a->addend == nullptr || (a->ld_addend % 4 == 0 && al16(a->addend))) &&
               (a->aux == nullptr || (a->ld_aux % 4 == 0 && al16(a->aux))) && (a->bias == nullptr || al16(a->bias)) &&
               (a->row_table == nullptr || (al16(a->row_table) && a->N % 4 == 0)) &&
               (a->act == VCT_ACT_NONE || (a->N % 8 == 0 && a->ldc % 8 == 0 && (a->C2 == nullptr || a->ldc2 % 8 == 0) &&
                                           (a->aux == nullptr || a->ld_aux % 8 == 0) &&
                                           (a->addend == nullptr || a->ld_addend % 8 == 0)));
    return e;
}

__device__ __forceinline__ void store4(void* base, int dtype, long long off, float4 v) {
    if (dtype == VCT_BF16) st4(reinterpret_cast<__nv_bfloat16*>(base) + off, v);
    else st4(reinterpret_cast<float*>(base) + off, v);
}
__device__ __forceinline__ void store1(void* base, int dtype, long long off, float v) {
    if (dtype == VCT_BF16) reinterpret_cast<__nv_bfloat16*>(base)[off] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(base)[off] = v;
}
__device__ __forceinline__ float load1(const void* base, int dtype, long long off) {
    return dtype == VCT_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[off])
                             : reinterpret_cast<const float*>(base)[off];
}

// slow path: ragged N or unaligned strides; one element at a time
template <int ACT>
__device__ __noinline__ void epilogue_scalar(const Epilogue& e, const Rng& rng, int m, int n, float4 acc) {
    const float v4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
        const int nn = n + q;
        if (nn >= e.N) break;
        float v = v4[q];
        if (e.bias) v += e.bias[nn];
        if (e.row_table) v += e.row_table[(long long)(m % e.row_period) * e.N + nn];
        float sc = 1.f;
        if (ACT != VCT_ACT_NONE && ACT != VCT_ACT_MUL_AUX) sc = dropout_scale1(rng, e.site, (unsigned long long)m * (unsigned long long)e.N + nn);
        if (ACT == VCT_ACT_GELU_FWD || ACT == VCT_ACT_GELU_FWD_F) {
            store1(e.C, e.c_dtype, (long long)m * e.ldc + nn, ACT == VCT_ACT_GELU_FWD ? v : dgelu_f(v) * sc);
            if (e.C2) store1(e.C2, e.c2_dtype, (long long)m * e.ldc2 + nn, gelu_f(v) * sc);
            continue;
        }
        if (ACT == VCT_ACT_GELU_BWD) v *= dgelu_f(load1(e.aux, e.aux_dtype, (long long)m * e.ld_aux + nn)) * sc;
        if (ACT == VCT_ACT_MUL_AUX) v *= load1(e.aux, e.aux_dtype, (long long)m * e.ld_aux + nn);
        if (e.addend) v += e.addend[(long long)m * e.ld_addend + nn];
        store1(e.C, e.c_dtype, (long long)m * e.ldc + nn, v);
        if (e.C2) store1(e.C2, e.c2_dtype, (long long)m * e.ldc2 + nn, v);
    }
}

// plain epilogue: acc = accumulators of row m, columns n..n+3 (n % 4 == 0).  Rows >= M / columns >= N are dropped.
__device__ __forceinline__ void epilogue_plain4(const Epilogue& e, const Rng& rng, int m, int n, float4 acc) {
    if (m >= e.M || n >= e.N) return;
    if (!e.vec_ok || n + 4 > e.N) { epilogue_scalar<VCT_ACT_NONE>(e, rng, m, n, acc); return; }
    if (e.bias) {
        const float4 b = ld4(e.bias + n);
        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    if (e.row_table) {
        const float4 t = ld4(e.row_table + (long long)(m % e.row_period) * e.N + n);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    if (e.addend) {
        const float4 ad = ld4(e.addend + (long long)m * e.ld_addend + n);
        acc.x += ad.x; acc.y += ad.y; acc.z += ad.z; acc.w += ad.w;
    }
    store4(e.C, e.c_dtype, (long long)m * e.ldc + n, acc);
    if (e.C2) store4(e.C2, e.c2_dtype, (long long)m * e.ldc2 + n, acc);
}

// a0, a1 = accumulators of row m, columns n..n+7 (n % 8 == 0): one Philox draw covers the 8 dropout decisions
template <int ACT>
__device__ __forceinline__ void epilogue_store8(const Epilogue& e, const Rng& rng, int m, int n, float4 a0, float4 a1) {
    if (ACT == VCT_ACT_NONE) {
        epilogue_plain4(e, rng, m, n, a0);
        epilogue_plain4(e, rng, m, n + 4, a1);
        return;
    }
    if (m >= e.M || n >= e.N) return;
    if (!e.vec_ok || n + 8 > e.N) {
        epilogue_scalar<ACT>(e, rng, m, n, a0);
        if (n + 4 < e.N) epilogue_scalar<ACT>(e, rng, m, n + 4, a1);
        return;
    }
    float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    if (e.bias) {
        float b[8];
        ld8(e.bias + n, b);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] += b[q];
    }
    if (e.row_table) {
        float t[8];
        ld8(e.row_table + (long long)(m % e.row_period) * e.N + n, t);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] += t[q];
    }
    float sc[8];
    if (ACT != VCT_ACT_MUL_AUX) dropout_scale8(rng, e.site, ((unsigned long long)m * (unsigned long long)e.N + (unsigned long long)n) >> 3, sc);
    if (ACT == VCT_ACT_GELU_FWD || ACT == VCT_ACT_GELU_FWD_F) {
        const long long o = (long long)m * e.ldc + n;
        float f[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] = ACT == VCT_ACT_GELU_FWD ? v[q] : dgelu_f(v[q]) * sc[q];
        if (e.c_dtype == VCT_BF16) st8(reinterpret_cast<__nv_bfloat16*>(e.C) + o, f);
        else st8(reinterpret_cast<float*>(e.C) + o, f);
        if (e.C2) {
            float h[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) h[q] = gelu_f(v[q]) * sc[q];
            const long long o2 = (long long)m * e.ldc2 + n;
            if (e.c2_dtype == VCT_BF16) st8(reinterpret_cast<__nv_bfloat16*>(e.C2) + o2, h);
            else st8(reinterpret_cast<float*>(e.C2) + o2, h);
        }
        return;
    }
    {   // GELU backward
        float z[8];
        const long long ao = (long long)m * e.ld_aux + n;
        if (e.aux_dtype == VCT_BF16) ld8(reinterpret_cast<const __nv_bfloat16*>(e.aux) + ao, z);
        else ld8(reinterpret_cast<const float*>(e.aux) + ao, z);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] *= ACT == VCT_ACT_MUL_AUX ? z[q] : dgelu_f(z[q]) * sc[q];
    }
    if (e.addend) {
        float ad[8];
        ld8(e.addend + (long long)m * e.ld_addend + n, ad);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] += ad[q];
    }
    const long long o = (long long)m * e.ldc + n;
    if (e.c_dtype == VCT_BF16) st8(reinterpret_cast<__nv_bfloat16*>(e.C) + o, v);
    else st8(reinterpret_cast<float*>(e.C) + o, v);
    if (e.C2) {
        const long long o2 = (long long)m * e.ldc2 + n;
        if (e.c2_dtype == VCT_BF16) st8(reinterpret_cast<__nv_bfloat16*>(e.C2) + o2, v);
        else st8(reinterpret_cast<float*>(e.C2) + o2, v);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Tile epilogues of the tcgen05 kernels.  128 threads (t = 0..127) store a [rows x BN] fp32 accumulator tile
// that has been staged in shared memory (row stride RS floats, `stage` = shared-space byte address).  Everything
// that does not depend on the row (column offset, bias, output pointers, feature flags) is hoisted out of the row
// loop; rows are processed four at a time -- four explicit ld.shared first, then the adds and the stores -- because
// the compiler may not move a load above a store that might alias it, which would serialise every iteration on the
// shared-memory latency.  A warp instruction covers 512 contiguous bytes of an output row.  The generic per-element
// functions above remain the path for ragged / unaligned tiles.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

struct PlainCtx {           // row-invariant state of epilogue_tile_plain
    float4 b;
    const float* table; int period, N, n;
    bool c_bf16, c2_bf16;
    char* c; char* c2; const float* ad;
    long long c_step, c2_step, ad_step;
};

__device__ __forceinline__ void plain_row(PlainCtx& x, float4 v, float4 a4, int m) {
    v.x += x.b.x; v.y += x.b.y; v.z += x.b.z; v.w += x.b.w;
    if (x.table) {
        const float4 tb = ld4(x.table + (long long)(m % x.period) * x.N + x.n);
        v.x += tb.x; v.y += tb.y; v.z += tb.z; v.w += tb.w;
    }
    v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w;
    if (x.c_bf16) st4(reinterpret_cast<__nv_bfloat16*>(x.c), v);
    else st4(reinterpret_cast<float*>(x.c), v);
    x.c += x.c_step;
    if (x.c2) {
        if (x.c2_bf16) st4(reinterpret_cast<__nv_bfloat16*>(x.c2), v);
        else st4(reinterpret_cast<float*>(x.c2), v);
        x.c2 += x.c2_step;
    }
}

template <int BN, int NT>
__device__ __forceinline__ void epilogue_tile_plain(const Epilogue& e, const Rng& rng, uint32_t stage, int RS, int m0, int n0, int t) {
    constexpr int CG4 = BN / 4, RSTEP = NT / CG4;           // 4-column groups per row; rows covered per pass of NT threads
    if (t >= RSTEP * CG4) return;
    const int c4 = t % CG4, n = n0 + c4 * 4;
    if (n >= e.N) return;
    const int rows = min(128, e.M - m0);
    int rr = t / CG4;
    uint32_t src = stage + (uint32_t)(rr * RS + c4 * 4) * 4u;
    const uint32_t src_step = (uint32_t)(RSTEP * RS) * 4u;
    if (!e.vec_ok || n + 4 > e.N) {
        for (; rr < rows; rr += RSTEP, src += src_step) epilogue_scalar<VCT_ACT_NONE>(e, rng, m0 + rr, n, lds128(src));
        return;
    }
    PlainCtx x;
    x.b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.bias) x.b = ld4(e.bias + n);
    x.table = e.row_table; x.period = e.row_period; x.N = e.N; x.n = n;
    x.c_bf16 = e.c_dtype == VCT_BF16; x.c2_bf16 = e.c2_dtype == VCT_BF16;
    const long long m = m0 + rr;
    x.c = reinterpret_cast<char*>(e.C) + (m * e.ldc + n) * (x.c_bf16 ? 2 : 4);
    x.c2 = e.C2 ? reinterpret_cast<char*>(e.C2) + (m * e.ldc2 + n) * (x.c2_bf16 ? 2 : 4) : nullptr;
    x.ad = e.addend ? e.addend + m * e.ld_addend + n : nullptr;
    x.c_step = (long long)RSTEP * e.ldc * (x.c_bf16 ? 2 : 4);
    x.c2_step = (long long)RSTEP * e.ldc2 * (x.c2_bf16 ? 2 : 4);
    x.ad_step = (long long)RSTEP * e.ld_addend;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (; rr + 3 * RSTEP < rows; rr += 4 * RSTEP, src += 4 * src_step) {
        float4 v[4], a4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = lds128(src + u * src_step);
#pragma unroll
        for (int u = 0; u < 4; ++u) a4[u] = x.ad ? ld4(x.ad + u * x.ad_step) : zero;
        if (x.ad) x.ad += 4 * x.ad_step;
#pragma unroll
        for (int u = 0; u < 4; ++u) plain_row(x, v[u], a4[u], m0 + rr + u * RSTEP);
    }
    for (; rr < rows; rr += RSTEP, src += src_step) {
        const float4 v = lds128(src);
        const float4 a4 = x.ad ? ld4(x.ad) : zero;
        if (x.ad) x.ad += x.ad_step;
        plain_row(x, v, a4, m0 + rr);
    }
}

// raw fp32 partial sums of a split-K slice: ws[m * ldw + n] (ws already offset to the slice)
template <int BN, int NT>
__device__ __forceinline__ void epilogue_tile_partial(float* __restrict__ ws, long long ldw, int M, int N, uint32_t stage, int RS,
                                                      int m0, int n0, int t) {
    constexpr int CG4 = BN / 4, RSTEP = NT / CG4;
    if (t >= RSTEP * CG4) return;
    const int c4 = t % CG4, n = n0 + c4 * 4;
    if (n >= N) return;                                     // ldw is a multiple of 8 >= N: whole float4 groups are in bounds
    const int rows = min(128, M - m0);
    int rr = t / CG4;
    uint32_t src = stage + (uint32_t)(rr * RS + c4 * 4) * 4u;
    const uint32_t src_step = (uint32_t)(RSTEP * RS) * 4u;
    float* dst = ws + (long long)(m0 + rr) * ldw + n;
    const long long dst_step = (long long)RSTEP * ldw;
    for (; rr + 3 * RSTEP < rows; rr += 4 * RSTEP, src += 4 * src_step, dst += 4 * dst_step) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = lds128(src + u * src_step);
#pragma unroll
        for (int u = 0; u < 4; ++u) *reinterpret_cast<float4*>(dst + u * dst_step) = v[u];
    }
    for (; rr < rows; rr += RSTEP, src += src_step, dst += dst_step) *reinterpret_cast<float4*>(dst) = lds128(src);
}

// activation epilogues (GELU forward / backward): 8-column groups per thread so that one Philox draw covers the
// group's 8 dropout decisions and bf16 outputs are stored 16 bytes at a time
template <int ACT, int BN, int NT>
__device__ __forceinline__ void epilogue_tile_act(const Epilogue& e, const Rng& rng, uint32_t stage, int RS, int m0, int n0, int t) {
    constexpr int CG = BN / 8, RSTEP = NT / CG;
    if (t >= RSTEP * CG) return;
    const int cg = t % CG, n = n0 + cg * 8;
    if (n >= e.N) return;
    const int rows = min(128, e.M - m0);
    int rr = t / CG;
    uint32_t src = stage + (uint32_t)(rr * RS + cg * 8) * 4u;
    const uint32_t src_step = (uint32_t)(RSTEP * RS) * 4u;
    if (!e.vec_ok || n + 8 > e.N) {
        for (; rr < rows; rr += RSTEP, src += src_step) epilogue_store8<ACT>(e, rng, m0 + rr, n, lds128(src), lds128(src + 16));
        return;
    }
    float b[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) b[q] = 0.f;
    if (e.bias) ld8(e.bias + n, b);
    const bool c_bf16 = e.c_dtype == VCT_BF16, c2_bf16 = e.c2_dtype == VCT_BF16, aux_bf16 = e.aux_dtype == VCT_BF16;
    const long long ldc = e.ldc, ldc2 = e.ldc2, ldx = e.ld_aux, lda = e.ld_addend;
    const unsigned long long N = (unsigned long long)e.N;
    const unsigned int site = e.site;
    // two rows per pass; the global operands of the NEXT pass (z and the addend of the GELU backward) are requested before
    // the math of the current one, so that their L2 latency is hidden behind ~500 instructions of work
    float z[2][8], ad[2][8], zn[2][8], adn[2][8];
    constexpr bool kBwd = ACT == VCT_ACT_GELU_BWD || ACT == VCT_ACT_MUL_AUX;
    constexpr bool kFwd = ACT == VCT_ACT_GELU_FWD || ACT == VCT_ACT_GELU_FWD_F;
    auto fetch = [&](int r0, float (&zz)[2][8], float (&aa)[2][8]) {
        if (!kBwd) return;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int rq = r0 + u * RSTEP;
            if (rq < rows) {
                const long long mm = m0 + rq;
                if (aux_bf16) ld8(reinterpret_cast<const __nv_bfloat16*>(e.aux) + mm * ldx + n, zz[u]);
                else ld8(reinterpret_cast<const float*>(e.aux) + mm * ldx + n, zz[u]);
                if (e.addend) ld8(e.addend + mm * lda + n, aa[u]);
            }
        }
    };
    fetch(rr, z, ad);
    for (; rr < rows; rr += 2 * RSTEP, src += 2 * src_step) {
        const bool two = rr + RSTEP < rows;
        float4 a[2][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            a[u][0] = lds128(src + u * src_step);
            a[u][1] = lds128(src + u * src_step + 16);
        }
        fetch(rr + 2 * RSTEP, zn, adn);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            const long long m = m0 + rr + u * RSTEP;
            float v[8] = {a[u][0].x + b[0], a[u][0].y + b[1], a[u][0].z + b[2], a[u][0].w + b[3],
                          a[u][1].x + b[4], a[u][1].y + b[5], a[u][1].z + b[6], a[u][1].w + b[7]};
            float sc[8];
            if (ACT != VCT_ACT_MUL_AUX) dropout_scale8(rng, site, ((unsigned long long)m * N + (unsigned long long)n) >> 3, sc);
            if (kFwd) {
                float hh[8], ff[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float cdf, pdf;
                    normal_cdf_pdf_fast(v[q], cdf, pdf);
                    hh[q] = v[q] * cdf * sc[q];
                    ff[q] = ACT == VCT_ACT_GELU_FWD ? v[q] : fmaf(v[q], pdf, cdf) * sc[q];
                }
                const long long o = m * ldc + n;
                if (c_bf16) st8(reinterpret_cast<__nv_bfloat16*>(e.C) + o, ff);
                else st8(reinterpret_cast<float*>(e.C) + o, ff);
                if (e.C2) {
                    const long long o2 = m * ldc2 + n;
                    if (c2_bf16) st8(reinterpret_cast<__nv_bfloat16*>(e.C2) + o2, hh);
                    else st8(reinterpret_cast<float*>(e.C2) + o2, hh);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] *= ACT == VCT_ACT_MUL_AUX ? z[u][q] : dgelu_fast(z[u][q]) * sc[q];
                if (e.addend) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] += ad[u][q];
                }
                const long long o = m * ldc + n;
                if (c_bf16) st8(reinterpret_cast<__nv_bfloat16*>(e.C) + o, v);
                else st8(reinterpret_cast<float*>(e.C) + o, v);
                if (e.C2) {
                    const long long o2 = m * ldc2 + n;
                    if (c2_bf16) st8(reinterpret_cast<__nv_bfloat16*>(e.C2) + o2, v);
                    else st8(reinterpret_cast<float*>(e.C2) + o2, v);
                }
            }
        }
        if (kBwd) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int q = 0; q < 8; ++q) { z[u][q] = zn[u][q]; ad[u][q] = adn[u][q]; }
        }
    }
}

}  // namespace vct
