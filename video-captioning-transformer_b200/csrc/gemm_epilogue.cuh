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
        if (ACT != VCT_ACT_NONE) sc = dropout_scale1(rng, e.site, (unsigned long long)m * (unsigned long long)e.N + nn);
        if (ACT == VCT_ACT_GELU_FWD) {
            store1(e.C, e.c_dtype, (long long)m * e.ldc + nn, v);
            if (e.C2) store1(e.C2, e.c2_dtype, (long long)m * e.ldc2 + nn, gelu_f(v) * sc);
            continue;
        }
        if (ACT == VCT_ACT_GELU_BWD) v *= dgelu_f(load1(e.aux, e.aux_dtype, (long long)m * e.ld_aux + nn)) * sc;
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
    dropout_scale8(rng, e.site, ((unsigned long long)m * (unsigned long long)e.N + (unsigned long long)n) >> 3, sc);
    if (ACT == VCT_ACT_GELU_FWD) {
        const long long o = (long long)m * e.ldc + n;
        if (e.c_dtype == VCT_BF16) st8(reinterpret_cast<__nv_bfloat16*>(e.C) + o, v);
        else st8(reinterpret_cast<float*>(e.C) + o, v);
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
        for (int q = 0; q < 8; ++q) v[q] *= dgelu_f(z[q]) * sc[q];
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

}  // namespace vct
