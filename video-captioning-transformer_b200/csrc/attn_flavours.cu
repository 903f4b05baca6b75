// The three attention flavours of the reference as single C-ABI calls.  In bf16 with the tcgen05 GEMM each flavour runs
// as ONE fused kernel (attn_fused.cu: projection, scores, softmax, dropout and value gather, all contractions on the
// tensor cores); the remaining cases (fp32, sequences longer than 32, cross-attention whose K/V are not projected
// yet) compose the packed in-projection GEMM with the stand-alone attention core.
#include "common.cuh"

using namespace vct;

namespace vct {
int attn_fused_self(const vct_mha_args* m, int causal, cudaStream_t st);   // attn_fused.cu
int attn_fused_cross(const vct_mha_args* m, cudaStream_t st);              // attn_fused.cu
}

namespace {

size_t esize(int dtype) { return dtype == VCT_BF16 ? 2 : 4; }

int project(const void* x, int rows, int d, int n_out, const void* w, const float* b, void* out, int dtype, int impl,
            vct_stream_t stream, void* split_ws = nullptr, long long split_ws_bytes = 0) {
    vct_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.M = rows; g.N = n_out; g.K = d;
    g.A = x; g.a_dtype = dtype; g.lda = d; g.a_trans = 0;
    g.B = w; g.b_dtype = dtype; g.ldb = d; g.b_trans = 0;
    g.C = out; g.c_dtype = dtype; g.ldc = n_out;
    g.bias = b;
    g.act = VCT_ACT_NONE;
    g.impl = impl;
    g.split_ws = split_ws; g.split_ws_bytes = split_ws_bytes;
    return vct_gemm(&g, stream);
}

int self_attention(const vct_mha_args* a, int causal, vct_stream_t stream, const char* who) {
    VCT_REQUIRE(a && a->x && a->w_in && a->qkv && a->o, "%s: null argument", who);
    VCT_REQUIRE(a->d % a->H == 0, "%s: d %% H != 0", who);
    const int d = a->d, rows = a->B * a->L;
    // bf16 + tcgen05: ONE kernel does the in-projection and the attention (attn_fused.cu); other dtypes / head sizes
    // compose the projection GEMM with the stand-alone attention core
    {
        const int r = vct::attn_fused_self(a, causal, (cudaStream_t)stream);
        if (r <= 0) return r;
    }
    if (int e = project(a->x, rows, d, 3 * d, a->w_in, a->b_in, a->qkv, a->dtype, a->gemm_impl, stream, a->split_ws, a->split_ws_bytes)) return e;
    vct_attn_args t;
    memset(&t, 0, sizeof(t));
    const char* base = (const char*)a->qkv;
    t.B = a->B; t.H = a->H; t.Lq = a->L; t.Lk = a->L; t.dh = d / a->H; t.dtype = a->dtype;
    t.q = base; t.q_ld = 3 * d;
    t.k = base + (size_t)d * esize(a->dtype); t.k_ld = 3 * d;
    t.v = base + (size_t)2 * d * esize(a->dtype); t.v_ld = 3 * d;
    t.o = a->o; t.o_ld = d;
    t.key_pad = a->key_pad; t.causal = causal;
    t.scale = 1.0f / sqrtf((float)t.dh);
    t.drop_p = a->drop_p; t.rng_state = a->rng_state; t.site = a->site;
    t.probs = a->probs;
    return vct_attn_fwd(&t, stream);
}

}  // namespace

extern "C" int vct_attn_enc_self_fwd(const vct_mha_args* a, vct_stream_t stream) {
    return self_attention(a, 0, stream, "vct_attn_enc_self_fwd");
}

extern "C" int vct_attn_dec_self_fwd(const vct_mha_args* a, vct_stream_t stream) {
    return self_attention(a, 1, stream, "vct_attn_dec_self_fwd");
}

extern "C" int vct_attn_dec_cross_fwd(const vct_mha_args* a, vct_stream_t stream) {
    VCT_REQUIRE(a && a->x && a->w_in && a->qkv && a->kv && a->o, "vct_attn_dec_cross_fwd: null argument");
    VCT_REQUIRE(a->kv_ready || a->mem, "vct_attn_dec_cross_fwd: mem is required unless kv_ready");
    VCT_REQUIRE(a->d % a->H == 0, "vct_attn_dec_cross_fwd: d %% H != 0");
    const int d = a->d;
    const size_t es = esize(a->dtype);
    // bf16 + tcgen05 with the memory's K/V already projected: ONE kernel (q projection + attention, attn_fused.cu)
    {
        const int r = vct::attn_fused_cross(a, (cudaStream_t)stream);
        if (r <= 0) return r;
    }
    // q = x W_in[0:d]^T + b_in[0:d]
    if (int e = project(a->x, a->B * a->L, d, d, a->w_in, a->b_in, a->qkv, a->dtype, a->gemm_impl, stream, a->split_ws, a->split_ws_bytes)) return e;
    if (!a->kv_ready) {
        const char* w_kv = (const char*)a->w_in + (size_t)d * d * es;
        if (int e = project(a->mem, a->B * a->Lk, d, 2 * d, w_kv, a->b_in ? a->b_in + d : nullptr, a->kv, a->dtype,
                            a->gemm_impl, stream, a->split_ws, a->split_ws_bytes))
            return e;
    }
    vct_attn_args t;
    memset(&t, 0, sizeof(t));
    t.B = a->B; t.H = a->H; t.Lq = a->L; t.Lk = a->Lk; t.dh = d / a->H; t.dtype = a->dtype;
    t.q = a->qkv; t.q_ld = d;
    t.k = a->kv; t.k_ld = 2 * d;
    t.v = (const char*)a->kv + (size_t)d * es; t.v_ld = 2 * d;
    t.o = a->o; t.o_ld = d;
    t.key_pad = nullptr; t.causal = 0;    // cross-attention is never masked
    t.scale = 1.0f / sqrtf((float)t.dh);
    t.drop_p = a->drop_p; t.rng_state = a->rng_state; t.site = a->site;
    t.probs = a->probs;
    return vct_attn_fwd(&t, stream);
}
