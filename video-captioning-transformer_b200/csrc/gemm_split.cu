// Reference-precision GEMM on the tensor cores ("split-bf16"): fp32 operands are decomposed into bf16 pieces
//     x = h + m (+ l),   h = bf16(x), m = bf16(x - h), l = bf16(x - h - m)       (the subtractions are exact in fp32)
// and the product is rebuilt from the leading cross terms with fp32 accumulation in TMEM,
//     3 terms:  a_h b_h + a_h b_m + a_m b_h                                       (|err| ~ 2^-16 |a||b| per product)
//     6 terms:  ... + a_m b_m + a_h b_l + a_l b_h                                 (|err| ~ 2^-23 |a||b|: fp32 class)
// The terms are not separate GEMMs: the pieces of A and B are laid side by side along the CONTRACTION axis
//     A' = [a_h | a_h | a_m | ...],   B' = [b_h | b_m | b_h | ...]     =>     A' B'^T = sum of the terms
// so ONE launch of the ordinary tcgen05 kernel with K' = terms * K produces the result, epilogue included.  The reference
// computes in fp32 (TF32 off, no autocast); tcgen05 has no fp32 MMA kind, and `north_star` asks for token-id
// argmax exactness, which bf16 operands cannot guarantee (62/64 agreement).  This mode is what
// precision = "bf16x3" / "bf16x6" of the engine selects: fp32 storage everywhere, every dense contraction on tcgen05.
#include "common.cuh"

using namespace vct;

namespace vct {
int gemm_tcgen05(const vct_gemm_args* a, cudaStream_t st);   // gemm_tc.cu
}

namespace {

// which piece (0 = h, 1 = m, 2 = l) each term takes from the A side / the B side
__constant__ int kPieceA[2][6] = {{0, 0, 1, 0, 0, 0}, {0, 0, 1, 1, 0, 2}};
__constant__ int kPieceB[2][6] = {{0, 1, 0, 0, 0, 0}, {0, 1, 0, 1, 2, 0}};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// src fp32 [rows, cols] (row stride ld_src) -> dst bf16.
//   axis 1 (contraction along the columns, K = cols):  dst [rows, terms * Kp], term j at columns [j * Kp, j * Kp + Kp)
//   axis 0 (contraction along the rows,    K = rows):  dst [terms * Kp, ld_dst], term j at rows    [j * Kp, j * Kp + Kp)
// Kp = K rounded up to 8; the padding (columns / rows K .. Kp-1) is written as zeros.  One thread = 8 consecutive columns.
template <int TERMS>
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ src, long long ld_src, int rows, int cols, int axis, int side,
                  __nv_bfloat16* __restrict__ dst, long long ld_dst, int Kp) {
    pdl_launch_dependents();
    pdl_wait();
    const int chunks = axis == 1 ? Kp / 8 : (cols + 7) / 8;
    const int out_rows = axis == 1 ? rows : Kp;
    const long long total = (long long)out_rows * chunks;
    const bool vec = (ld_src % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / chunks), c = (int)(idx % chunks);
        float x[8];
        const int c0 = c * 8;
        if (r < rows && vec && c0 + 8 <= cols) {
            const float4 a = *reinterpret_cast<const float4*>(src + (long long)r * ld_src + c0);
            const float4 b = *reinterpret_cast<const float4*>(src + (long long)r * ld_src + c0 + 4);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = (r < rows && c0 + q < cols) ? src[(long long)r * ld_src + c0 + q] : 0.f;
        }
        uint4 piece[3];
        {
            float h[8], m[8], l[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                h[q] = __bfloat162float(__float2bfloat16_rn(x[q]));
                const float r1 = x[q] - h[q];
                m[q] = __bfloat162float(__float2bfloat16_rn(r1));
                l[q] = r1 - m[q];
            }
            piece[0] = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
            piece[1] = make_uint4(pack2(m[0], m[1]), pack2(m[2], m[3]), pack2(m[4], m[5]), pack2(m[6], m[7]));
            piece[2] = make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
        }
        const int* tab = side == 0 ? kPieceA[TERMS == 6] : kPieceB[TERMS == 6];
#pragma unroll
        for (int j = 0; j < TERMS; ++j) {
            const int p = tab[j];
            const uint4 v = p == 0 ? piece[0] : (p == 1 ? piece[1] : piece[2]);
            const long long off = axis == 1 ? (long long)r * ld_dst + (long long)j * Kp + c0
                                            : ((long long)j * Kp + r) * ld_dst + c0;
            *reinterpret_cast<uint4*>(dst + off) = v;
        }
    }
}

inline long long rup(long long v, long long m) { return (v + m - 1) / m * m; }

int launch_split(const float* src, long long ld_src, int rows, int cols, int axis, int side, int terms, __nv_bfloat16* dst,
                 long long ld_dst, int Kp, cudaStream_t st) {
    const long long total = (long long)(axis == 1 ? rows : Kp) * (axis == 1 ? Kp / 8 : (cols + 7) / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
    if (blocks < 1) blocks = 1;
    if (terms == 3)
        vct::launch(split_bf16_kernel<3>, dim3((unsigned)blocks), dim3(256), 0, st, src, ld_src, rows, cols, axis, side, dst, ld_dst, Kp);
    else
        vct::launch(split_bf16_kernel<6>, dim3((unsigned)blocks), dim3(256), 0, st, src, ld_src, rows, cols, axis, side, dst, ld_dst, Kp);
    return check_launch("vct_split_bf16");
}

}  // namespace

extern "C" long long vct_gemm_split_workspace_bytes(int M, int N, int K, int a_trans, int b_trans, int terms) {
    const long long Kp = rup(K, 8), t = terms;
    const long long a = a_trans ? t * Kp * rup(M, 8) : (long long)M * t * Kp;
    const long long b = b_trans ? t * Kp * rup(N, 8) : (long long)N * t * Kp;
    return rup(a * 2, 256) + rup(b * 2, 256);
}

extern "C" int vct_split_bf16(const float* src, long long ld_src, int rows, int cols, int axis, int side, int terms, void* dst,
                              long long ld_dst, vct_stream_t stream) {
    VCT_REQUIRE(src && dst && rows > 0 && cols > 0, "vct_split_bf16: null / empty argument");
    VCT_REQUIRE(terms == 3 || terms == 6, "vct_split_bf16: terms must be 3 or 6");
    VCT_REQUIRE((axis == 0 || axis == 1) && (side == 0 || side == 1), "vct_split_bf16: axis / side must be 0 or 1");
    VCT_REQUIRE(ld_dst % 8 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "vct_split_bf16: dst must be 16-byte aligned, ld_dst %% 8 == 0");
    const int Kp = (int)rup(axis == 1 ? cols : rows, 8);
    VCT_REQUIRE(ld_dst >= (axis == 1 ? (long long)terms * Kp : rup(cols, 8)), "vct_split_bf16: ld_dst too small");
    return launch_split(src, ld_src, rows, cols, axis, side, terms, (__nv_bfloat16*)dst, ld_dst, Kp, (cudaStream_t)stream);
}

namespace vct {

// vct_gemm with impl = VCT_GEMM_TCGEN05_X3 / _X6: fp32 A and B, split into the caller's workspace, one tcgen05 launch.
int gemm_split(const vct_gemm_args* a, int terms, cudaStream_t st) {
    VCT_REQUIRE(a->a_dtype == VCT_F32, "vct_gemm(split-bf16): operands must be fp32 (bf16 operands go to VCT_GEMM_TCGEN05)");
    VCT_REQUIRE(a->split_ws != nullptr && (reinterpret_cast<uintptr_t>(a->split_ws) & 255) == 0,
                "vct_gemm(split-bf16): split_ws (256-byte aligned) is required");
    const long long need = vct_gemm_split_workspace_bytes(a->M, a->N, a->K, a->a_trans, a->b_trans, terms);
    VCT_REQUIRE(need <= a->split_ws_bytes, "vct_gemm(split-bf16): split_ws too small (%lld > %lld bytes)", need, a->split_ws_bytes);
    const int Kp = (int)rup(a->K, 8);
    const long long t = terms;
    __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(a->split_ws);
    const long long a_elems = a->a_trans ? t * Kp * rup(a->M, 8) : (long long)a->M * t * Kp;
    __nv_bfloat16* Bs = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(a->split_ws) + rup(a_elems * 2, 256));
    const long long lda = a->a_trans ? rup(a->M, 8) : t * Kp, ldb = a->b_trans ? rup(a->N, 8) : t * Kp;
    // A stored [M, K] (contraction along columns) or [K, M] (along rows); same for B with N
    if (int e = launch_split((const float*)a->A, a->lda, a->a_trans ? a->K : a->M, a->a_trans ? a->M : a->K, a->a_trans ? 0 : 1, 0,
                             terms, As, lda, Kp, st)) return e;
    if (int e = launch_split((const float*)a->B, a->ldb, a->b_trans ? a->K : a->N, a->b_trans ? a->N : a->K, a->b_trans ? 0 : 1, 1,
                             terms, Bs, ldb, Kp, st)) return e;
    vct_gemm_args g = *a;
    g.A = As; g.a_dtype = VCT_BF16; g.lda = lda;
    g.B = Bs; g.b_dtype = VCT_BF16; g.ldb = ldb;
    g.K = (int)(t * Kp);
    g.impl = VCT_GEMM_TCGEN05;
    return gemm_tcgen05(&g, st);
}

}  // namespace vct
