// SIMT (fp32 FFMA) GEMM with the fused epilogue: the exact-arithmetic path (fp32 operands,
// used for bring-up parity and argmax-exact decoding) and the cross-check for the tcgen05 kernel
// (bf16 operands, identical rounding points).  128x128x16 tiles, 256 threads, 8x8 per thread.
#include "gemm_epilogue.cuh"

using namespace vct;

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, THREADS = 256;

// 8 consecutive elements starting at p (elements beyond `valid` read as 0).  `vec` = 16B-aligned fast path.
template <typename T>
__device__ __forceinline__ void load8(const T* p, int valid, bool vec, float* out) {
    if (vec && valid >= 8) {
        float4 a = ld4(p), b = ld4(p + 4);
        out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
        out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) out[e] = e < valid ? to_f32(p[e]) : 0.f;
    }
}

// Stage a [ROWS=128][BK] operand tile into smem as S[k][row].
//  TRANS = false: memory is [rows][K] (K contiguous);  TRANS = true: memory is [K][rows] (rows contiguous)
template <typename T, bool TRANS>
__device__ __forceinline__ void fetch_tile(const T* __restrict__ base, long long ld, int rows_total, int K, int row0,
                                           int k0, bool vec_ok, float* reg) {
    const int t = threadIdx.x;
    if (!TRANS) {
        const int r = t >> 1, kk = (t & 1) * 8;
        const int row = row0 + r, k = k0 + kk;
        int valid = row < rows_total ? K - k : 0;
        valid = valid < 0 ? 0 : valid;
        load8(base + (long long)row * ld + k, valid, vec_ok, reg);
    } else {
        const int kk = t >> 4, r0 = (t & 15) * 8;
        const int k = k0 + kk, row = row0 + r0;
        int valid = k < K ? rows_total - row : 0;
        valid = valid < 0 ? 0 : valid;
        load8(base + (long long)k * ld + row, valid, vec_ok, reg);
    }
}

template <bool TRANS>
__device__ __forceinline__ void stash_tile(float (*S)[BM + PAD], const float* reg) {
    const int t = threadIdx.x;
    if (!TRANS) {
        const int r = t >> 1, kk = (t & 1) * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) S[kk + e][r] = reg[e];
    } else {
        const int kk = t >> 4, r0 = (t & 15) * 8;
        *reinterpret_cast<float4*>(&S[kk][r0]) = make_float4(reg[0], reg[1], reg[2], reg[3]);
        *reinterpret_cast<float4*>(&S[kk][r0 + 4]) = make_float4(reg[4], reg[5], reg[6], reg[7]);
    }
}

template <typename T, bool A_TRANS, bool B_TRANS, int ACT>
__global__ void __launch_bounds__(THREADS)
gemm_simt_kernel(const T* __restrict__ A, long long lda, const T* __restrict__ B, long long ldb, int M, int N, int K,
                 bool a_vec, bool b_vec, Epilogue epi) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    fetch_tile<T, A_TRANS>(A, lda, M, K, m0, 0, a_vec, ra);
    fetch_tile<T, B_TRANS>(B, ldb, N, K, n0, 0, b_vec, rb);
    for (int k0 = 0; k0 < K; k0 += BK) {
        stash_tile<A_TRANS>(As, ra);
        stash_tile<B_TRANS>(Bs, rb);
        __syncthreads();
        if (k0 + BK < K) {
            fetch_tile<T, A_TRANS>(A, lda, M, K, m0, k0 + BK, a_vec, ra);
            fetch_tile<T, B_TRANS>(B, ldb, N, K, n0, k0 + BK, b_vec, rb);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    const Rng rng = make_rng(epi.rng_state, ACT != VCT_ACT_NONE ? epi.drop_p : 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        epilogue_store8<ACT>(epi, rng, m, n0 + tx * 8, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]),
                             make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
    }
}

template <typename T>
int launch_simt(const vct_gemm_args* a, cudaStream_t st) {
    const Epilogue epi = make_epilogue(a);
    dim3 grid((a->N + BN - 1) / BN, (a->M + BM - 1) / BM);
    const int esz = (int)sizeof(T);
    const long long vecel = 16 / esz;   // elements per 16 bytes
    const bool a_vec = (a->lda % vecel == 0) && ((reinterpret_cast<uintptr_t>(a->A) & 15) == 0);
    const bool b_vec = (a->ldb % vecel == 0) && ((reinterpret_cast<uintptr_t>(a->B) & 15) == 0);
    const T* A = (const T*)a->A;
    const T* B = (const T*)a->B;
#define GO(AT, BT, ACT) vct::launch(gemm_simt_kernel<T, AT, BT, ACT>, dim3(grid), dim3(THREADS), 0, st, A, a->lda, B, a->ldb, a->M, a->N, a->K, a_vec, b_vec, epi)
#define GO_ACT(AT, BT)                                      \
    if (a->act == VCT_ACT_GELU_FWD) GO(AT, BT, VCT_ACT_GELU_FWD);   \
    else if (a->act == VCT_ACT_GELU_BWD) GO(AT, BT, VCT_ACT_GELU_BWD); \
    else if (a->act == VCT_ACT_GELU_FWD_F) GO(AT, BT, VCT_ACT_GELU_FWD_F); \
    else if (a->act == VCT_ACT_MUL_AUX) GO(AT, BT, VCT_ACT_MUL_AUX); \
    else GO(AT, BT, VCT_ACT_NONE)
    if (!a->a_trans && !a->b_trans) { GO_ACT(false, false); }
    else if (!a->a_trans && a->b_trans) { GO_ACT(false, true); }
    else if (a->a_trans && !a->b_trans) { GO(true, false, VCT_ACT_NONE); }
    else { GO(true, true, VCT_ACT_NONE); }
#undef GO_ACT
#undef GO
    return check_launch("vct_gemm(simt)");
}

}  // namespace

namespace vct {
int gemm_simt(const vct_gemm_args* a, cudaStream_t st) {
    if (a->a_dtype == VCT_BF16) return launch_simt<__nv_bfloat16>(a, st);
    return launch_simt<float>(a, st);
}
int gemm_tcgen05(const vct_gemm_args* a, cudaStream_t st);   // gemm_tc.cu
int gemm_split(const vct_gemm_args* a, int terms, cudaStream_t st);   // gemm_split.cu
int gemm_grouped(const vct_gemm_args* args, int count, cudaStream_t st);   // gemm_tc.cu
void gemm_tune(int bn, int splits, int low);                  // gemm_tc.cu
void gemm_trace(long long* dev_buf);                          // gemm_tc.cu
}  // namespace vct

extern "C" int vct_gemm_tune(int block_n, int splits, int ring) {
    VCT_REQUIRE(block_n == 0 || block_n == 64 || block_n == 128 || block_n == 256, "vct_gemm_tune: block_n must be 0, 64, 128 or 256");
    VCT_REQUIRE(splits >= 0 && splits <= 8 && ring >= -1 && ring <= 1, "vct_gemm_tune: bad splits / ring");
    vct::gemm_tune(block_n, splits, ring);
    return 0;
}

extern "C" int vct_gemm_trace(void* dev_buf) {
    vct::gemm_trace(reinterpret_cast<long long*>(dev_buf));
    return 0;
}

extern "C" int vct_gemm(const vct_gemm_args* a, vct_stream_t stream) {
    VCT_REQUIRE(a != nullptr, "vct_gemm: null args");
    VCT_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "vct_gemm: empty problem (M=%d N=%d K=%d)", a->M, a->N, a->K);
    VCT_REQUIRE(a->A && a->B && a->C, "vct_gemm: null operand");
    VCT_REQUIRE(a->a_dtype == a->b_dtype, "vct_gemm: A and B must share a dtype");
    VCT_REQUIRE((a->a_dtype == VCT_F32 || a->a_dtype == VCT_BF16) && (a->c_dtype == VCT_F32 || a->c_dtype == VCT_BF16),
                "vct_gemm: bad dtype");
    VCT_REQUIRE(a->lda >= (a->a_trans ? a->M : a->K) && a->ldb >= (a->b_trans ? a->N : a->K) && a->ldc >= a->N,
                "vct_gemm: leading dimension too small");
    VCT_REQUIRE((reinterpret_cast<uintptr_t>(a->C) & 15) == 0 && (a->C2 == nullptr || (reinterpret_cast<uintptr_t>(a->C2) & 15) == 0),
                "vct_gemm: outputs must be 16-byte aligned");
    VCT_REQUIRE(a->act >= VCT_ACT_NONE && a->act <= VCT_ACT_MUL_AUX, "vct_gemm: unknown activation %d", a->act);
    VCT_REQUIRE((a->act != VCT_ACT_GELU_BWD && a->act != VCT_ACT_MUL_AUX) || a->aux != nullptr,
                "vct_gemm: GELU_BWD / MUL_AUX need aux (the pre-activation / the saved factor)");
    VCT_REQUIRE(a->act == VCT_ACT_NONE || !a->a_trans, "vct_gemm: activation epilogues are built for a_trans = 0 only");
    VCT_REQUIRE(a->row_table == nullptr || a->row_period > 0, "vct_gemm: row_table needs row_period > 0");
    VCT_REQUIRE(a->addend == nullptr || a->ld_addend >= a->N, "vct_gemm: ld_addend too small");
    if (a->impl == VCT_GEMM_TCGEN05) return vct::gemm_tcgen05(a, (cudaStream_t)stream);
    if (a->impl == VCT_GEMM_TCGEN05_X3) return vct::gemm_split(a, 3, (cudaStream_t)stream);
    if (a->impl == VCT_GEMM_TCGEN05_X6) return vct::gemm_split(a, 6, (cudaStream_t)stream);
    return vct::gemm_simt(a, (cudaStream_t)stream);
}

// Several independent GEMMs in one call.  Weight-gradient groups (bf16 operands stored [K, M] / [K, N], fp32 output, no
// epilogue extras, <= 8 problems, tcgen05) run as ONE persistent launch over the tiles of all problems; any other group is
// issued problem by problem.
extern "C" int vct_gemm_grouped(const vct_gemm_args* args, int count, vct_stream_t stream) {
    VCT_REQUIRE(args != nullptr && count >= 1, "vct_gemm_grouped: empty group");
    const int r = vct::gemm_grouped(args, count, (cudaStream_t)stream);
    if (r <= 0) return r;
    for (int i = 0; i < count; ++i)
        if (int e = vct_gemm(&args[i], stream)) return e;
    return 0;
}
