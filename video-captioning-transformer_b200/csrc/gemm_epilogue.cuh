// Fused GEMM epilogue shared by the SIMT and the tcgen05 kernels.
// Order: + bias[n]; + row_table[m % period, n]; activation; + addend[m, n]; store C (and C2).
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
    return e;
}

__device__ __forceinline__ void store_vals(void* base, int dtype, long long off, const float* v, int cnt, bool vec) {
    if (dtype == VCT_BF16) {
        __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base) + off;
        if (vec) st4(p, make_float4(v[0], v[1], v[2], v[3]));
        else for (int q = 0; q < cnt; ++q) p[q] = __float2bfloat16_rn(v[q]);
    } else {
        float* p = reinterpret_cast<float*>(base) + off;
        if (vec) st4(p, make_float4(v[0], v[1], v[2], v[3]));
        else for (int q = 0; q < cnt; ++q) p[q] = v[q];
    }
}

// v[0..3] = accumulators of row m, columns n..n+3 (n % 4 == 0).  Columns >= N are dropped.
__device__ __forceinline__ void epilogue_store4(const Epilogue& e, const Rng& rng, int m, int n, float* v) {
    if (m >= e.M || n >= e.N) return;
    const int cnt = e.N - n < 4 ? e.N - n : 4;
    if (e.bias) {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (q < cnt) v[q] += e.bias[n + q];
    }
    if (e.row_table) {
        const float* t = e.row_table + (long long)(m % e.row_period) * e.N + n;
#pragma unroll
        for (int q = 0; q < 4; ++q) if (q < cnt) v[q] += t[q];
    }
    float sc[4] = {1.f, 1.f, 1.f, 1.f};
    if (e.act != VCT_ACT_NONE && rng.p > 0.f) {
        const unsigned long long idx = (unsigned long long)m * (unsigned long long)e.N + (unsigned long long)n;
        if ((idx & 3ull) == 0ull) {
            float4 s4 = dropout_scale4(rng, e.site, idx >> 2);
            sc[0] = s4.x; sc[1] = s4.y; sc[2] = s4.z; sc[3] = s4.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) sc[q] = dropout_scale1(rng, e.site, idx + q);
        }
    }
    if (e.act == VCT_ACT_GELU_FWD) {
        const bool vec1 = cnt == 4 && (e.ldc & 3) == 0;
        store_vals(e.C, e.c_dtype, (long long)m * e.ldc + n, v, cnt, vec1);
        float h[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) h[q] = gelu_f(v[q]) * sc[q];
        if (e.C2) store_vals(e.C2, e.c2_dtype, (long long)m * e.ldc2 + n, h, cnt, cnt == 4 && (e.ldc2 & 3) == 0);
        return;
    }
    if (e.act == VCT_ACT_GELU_BWD) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q < cnt) {
                const long long o = (long long)m * e.ld_aux + n + q;
                const float z = e.aux_dtype == VCT_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(e.aux)[o])
                                                        : reinterpret_cast<const float*>(e.aux)[o];
                v[q] *= dgelu_f(z) * sc[q];
            }
        }
    }
    if (e.addend) {
        const float* ad = e.addend + (long long)m * e.ld_addend + n;
#pragma unroll
        for (int q = 0; q < 4; ++q) if (q < cnt) v[q] += ad[q];
    }
    store_vals(e.C, e.c_dtype, (long long)m * e.ldc + n, v, cnt, cnt == 4 && (e.ldc & 3) == 0);
    if (e.C2) store_vals(e.C2, e.c2_dtype, (long long)m * e.ldc2 + n, v, cnt, cnt == 4 && (e.ldc2 & 3) == 0);
}

}  // namespace vct
