// Fused attention forward for sm_100a: in-projection + scaled dot product + softmax + value gather in ONE kernel per
// flavour, every contraction on tcgen05 (the reference reaches this through torch/nn/modules/transformer.py _sa_block /
// _mha_block -> F.multi_head_attention_forward: packed in-projection, SDPA with merged masks, attention dropout).
//
//   self  (encoder self / decoder causal self): Q|K|V = x W_in[h]^T + b          (12 k-blocks of 64, TMA ring)
//   cross (decoder cross, K/V of the memory pre-projected): Q = x W_q[h]^T + b;  K, V tiles of head h by TMA
//
// One CTA per (group of nb batch elements, head).  Sequences are padded to LQP / LKP rows (16, 32 or 64) by the
// 3-D tensor maps (out-of-bounds rows are zero-filled), so that a warp's 32 accumulator rows belong to whole
// sequences and every tcgen05.ld address is warp-uniform:
//
//   1. projection     A = x rows [128 x d] (3-D TMA box: 64 K-columns x LQP rows x nb sequences), B = the head's rows of
//                     in_proj_weight (self: ONE 3-D box fetches the Q, K and V slices); fp32 accumulators in TMEM
//   2. epilogue 1     TMEM -> +bias -> bf16 -> Q, K (K-major, SWIZZLE_128B) and V (MN-major) operand tiles in shared
//                     memory (the operand ring is free by then) + q/k/v rows to global for the backward pass
//   3. scores         S[128 x NK] = Q K^T over the whole group (block diagonal is what is used), one tcgen05 chain
//   4. softmax        one THREAD per query row reads its sequence's LKP score columns from TMEM, applies the key-padding /
//                     causal masks, softmax, Philox dropout (same index space as the stand-alone core, so backward
//                     regenerates the mask), and writes the bf16 probability row (zero outside its block) as the A tile
//   5. values         O[128 x dh] = P V on tcgen05, TMEM -> bf16 rows of the attention output (before out_proj)
//
// Roles: warp 0 = A producer, warp 6 = B producer, warp 1 = MMA issuer, warps 2-5 = epilogue / softmax (one TMEM lane
// quadrant each).  Role loops are warp-uniform with one elected lane issuing (see gemm_tc.cu for the measurements).
#include <cstdlib>

#include "attn_device.cuh"
#include "tc_common.cuh"

using namespace vct;

namespace vct {
int get_tensor_map_3d(const void* ptr, const unsigned long long dims[3], const unsigned long long strides_bytes[2],
                      const unsigned int box[3], CUtensorMap* out);   // gemm_tc.cu
long long* gemm_trace_ptr();                                           // gemm_tc.cu (vct_gemm_trace)
}

namespace {

constexpr int kThreads = 224;
constexpr int kStages = 4;
constexpr uint32_t kBlk = 16384;            // one [128 rows x 128 B] swizzle-128B block
constexpr uint32_t kTile = 2 * kBlk;        // Q / K / V / P operand tile: two 64-element column blocks
constexpr uint32_t kColS = 288, kColO = 416;   // TMEM columns: [0, 3 dh) projection, [288, 416) scores, [416, 512) output

struct FusedArgs {
    int B, Lq, Lk, d, H, nb;
    const float* b_in;
    const unsigned char* key_pad;
    int causal;
    float scale;
    __nv_bfloat16* q_out; long long q_ld;      // saved projections: self [B*L, 3d] (q|k|v), cross [B*Lq, d] (q); may be NULL
    __nv_bfloat16* o;
    float* probs;
    float drop_p;
    const unsigned long long* rng_state;
    unsigned int site;
    long long* trace;      // debug (vct_gemm_trace): CTA (0,0) writes clock64 stamps of its phases to [100..110]
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// byte address of 16-byte chunk `cidx` (8 bf16) of row `row` in a two-block SWIZZLE_128B tile (rows 128 B apart inside a
// block, 16-byte chunks XOR-swizzled with the row index, second block of 64 elements 16 KB further)
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int cidx) {
    return base + (uint32_t)(cidx >> 3) * kBlk + (uint32_t)row * 128u + (uint32_t)(((cidx & 7) ^ (row & 7)) << 4);
}

template <int DH, int LQP, int LKP, bool CROSS>
__global__ void __launch_bounds__(kThreads, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmKV, FusedArgs a) {
    constexpr int NSEC = CROSS ? 1 : 3;
    constexpr uint32_t kWBytes = NSEC * DH * 128;
    constexpr uint32_t kStageBytes = kABytes + kWBytes;
    constexpr int NB = (128 / LQP) < (128 / LKP) ? (128 / LQP) : (128 / LKP);    // batch elements per CTA
    constexpr int NK = NB * LKP;                                                   // key rows of the group
    constexpr uint32_t kXBytes = (uint32_t)NB * LQP * 128;                         // bytes of one A box
    constexpr int KSTEPS = DH / 16;
    constexpr int NBLK = (DH + 63) / 64;
    static_assert(!CROSS ? LQP == LKP : true, "self-attention pads queries and keys alike");
    static_assert(DH % 32 == 0 && DH <= 96 && NK % 16 == 0 && NK <= 128, "unsupported head / group shape");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) unsigned long long full_bar[kStages];
    __shared__ __align__(8) unsigned long long empty_bar[kStages];
    __shared__ __align__(8) unsigned long long phase_bar[6];   // 0 proj done, 1 q/k/v tiles ready, 2 scores done, 3 P ready, 4 O done, 5 K/V landed
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float bias_s[3 * DH];

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* const trace = (blockIdx.x == 0 && blockIdx.y == 0) ? a.trace : nullptr;
    if (trace && threadIdx.x == 0) trace[100] = clock64();
    const int h = blockIdx.y;
    const int b0 = blockIdx.x * NB;
    const int d = a.d;
    const int num_kb = (d + BLOCK_K - 1) / BLOCK_K;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]), ph0 = smem_u32(&phase_bar[0]);
    // operand tiles: Q and P alias the (drained) operand ring; self-attention K and V too, cross K / V are TMA targets
    const uint32_t q_tile = smem0, p_tile = smem0 + kTile;
    const uint32_t k_tile = CROSS ? smem0 + kStages * kStageBytes : smem0 + 2 * kTile;
    const uint32_t v_tile = k_tile + kTile;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        if (CROSS) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8u * s, 2);
            mbar_init(empty0 + 8u * s, 1);
        }
        mbar_init(ph0 + 0, 1);
        mbar_init(ph0 + 8, 128);
        mbar_init(ph0 + 16, 1);
        mbar_init(ph0 + 24, 128);
        mbar_init(ph0 + 32, 1);
        mbar_init(ph0 + 40, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    pdl_wait();
    if (trace && threadIdx.x == 0) trace[101] = clock64();

    if (warp == 0 || warp == 6) {
        // ===== TMA producers =====
        const bool is_a = warp == 0;
        if (CROSS && !is_a) {
            if (elect_one()) {
                // K and V tiles of head h for the whole group: 64-column boxes x LKP rows x NB sequences each
                mbar_expect_tx(ph0 + 40, 2u * NBLK * (uint32_t)NK * 128u);
#pragma unroll
                for (int blk = 0; blk < NBLK; ++blk) {
                    tma_load_3d(k_tile + blk * kBlk, &tmKV, h * DH + blk * 64, 0, b0, ph0 + 40);
                    tma_load_3d(v_tile + blk * kBlk, &tmKV, d + h * DH + blk * 64, 0, b0, ph0 + 40);
                }
            }
            __syncwarp();
        }
        uint32_t ph = 1u;
        int k0 = 0;
        for (int kbase = 0; kbase < num_kb; kbase += kStages, ph ^= 1u) {
#pragma unroll
            for (int s = 0; s < kStages; ++s) {
                if (kbase + s < num_kb) {
                    mbar_wait_fast(empty0 + 8u * s, ph);
                    const uint32_t full = full0 + 8u * s;
                    const uint32_t sa = smem0 + (uint32_t)s * kStageBytes, sb = sa + kABytes;
                    if (elect_one()) {
                        if (is_a) {
                            mbar_expect_tx(full, kXBytes);
                            tma_load_3d(sa, &tmX, k0, 0, b0, full);
                        } else {
                            mbar_expect_tx(full, kWBytes);
                            if (CROSS) tma_load_2d(sb, &tmW, k0, h * DH, full);
                            else tma_load_3d(sb, &tmW, k0, h * DH, 0, full);      // Q, K and V rows of head h in one box
                        }
                    }
                    __syncwarp();
                    k0 += BLOCK_K;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t base_idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_M >> 4) << 24);
        const uint32_t idesc_qk = base_idesc | ((uint32_t)((2 * DH) >> 3) << 17);
        const uint32_t idesc_1 = base_idesc | ((uint32_t)(DH >> 3) << 17);
        const uint64_t adesc0 = make_desc(smem0, 16, 1024);
        const uint64_t bdesc0 = make_desc(smem0 + kABytes, 16, 1024);
        uint32_t ph = 0u, accum = 0u;
        for (int kbase = 0; kbase < num_kb; kbase += kStages, ph ^= 1u) {
#pragma unroll
            for (int s = 0; s < kStages; ++s) {
                if (kbase + s < num_kb) {
                    mbar_wait_fast(full0 + 8u * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t adesc = adesc0 + (uint64_t)(((uint32_t)s * kStageBytes) >> 4);
                    const uint64_t bdesc = bdesc0 + (uint64_t)(((uint32_t)s * kStageBytes) >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            const uint32_t acc = (accum | (uint32_t)k) != 0u ? 1u : 0u;
                            if (CROSS) {
                                umma_bf16(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc_1, acc);
                            } else {
                                umma_bf16(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc_qk, acc);
                                umma_bf16(tmem_base + 2 * DH, adesc + 2u * k, bdesc + (uint64_t)((2 * DH * 128) >> 4) + 2u * k, idesc_1, acc);
                            }
                        }
                        umma_commit(empty0 + 8u * s);
                    }
                    __syncwarp();
                    accum = 1u;
                }
            }
        }
        if (elect_one()) umma_commit(ph0 + 0);                  // projection complete
        __syncwarp();
        if (trace && lane == 0) trace[102] = clock64();
        // ---- scores: S = Q K^T over the group (both operands K-major, K = dh) ----
        mbar_wait(ph0 + 8, 0);
        if (CROSS) mbar_wait(ph0 + 40, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
            const uint32_t idesc_s = base_idesc | ((uint32_t)(NK >> 3) << 17);
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const uint32_t off = (uint32_t)(ks >> 2) * kBlk + (uint32_t)(ks & 3) * 32u;
                umma_bf16(tmem_base + kColS, make_desc(q_tile + off, 16, 1024), make_desc(k_tile + off, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
            }
            umma_commit(ph0 + 16);
        }
        __syncwarp();
        // ---- values: O = P V (A = probabilities K-major over the keys, B = V MN-major) ----
        mbar_wait(ph0 + 24, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
            const uint32_t idesc_o = idesc_1 | (1u << 16);
#pragma unroll
            for (int ks = 0; ks < NK / 16; ++ks) {
                const uint32_t aoff = (uint32_t)(ks >> 2) * kBlk + (uint32_t)(ks & 3) * 32u;
                umma_bf16(tmem_base + kColO, make_desc(p_tile + aoff, 16, 1024), make_desc(v_tile + (uint32_t)ks * 2048u, kBlk, 1024), idesc_o,
                          ks > 0 ? 1u : 0u);
            }
            umma_commit(ph0 + 32);
        }
        __syncwarp();
    } else {
        // ===== epilogue / softmax warps: thread = accumulator row r = TMEM lane =====
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int j = r / LQP, i = r % LQP;                 // sequence of the group, position in it
        const int b = b0 + j;
        const bool valid = j < NB && b < a.B && i < a.Lq;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        // ---- while the projection runs (these warps would only spin): stage the head's bias slice in shared memory,
        //      build this row's key mask and draw its dropout decisions ----
        {
            const int t = threadIdx.x - 64;
            for (int c = t; c < NSEC * DH; c += 128) bias_s[c] = a.b_in ? a.b_in[(c / DH) * d + h * DH + (c % DH)] : 0.f;
        }
        const long long prow = ((long long)b * a.H + h) * a.Lq + i;      // probability row index (dropout + need_weights)
        unsigned long long okmask = 0ull, keepmask = ~0ull;
        if (valid) {
            const unsigned char* pad_row = a.key_pad ? a.key_pad + (long long)b * a.Lk : nullptr;
#pragma unroll
            for (int kk = 0; kk < LKP; ++kk) {
                const bool ok = kk < a.Lk && !(a.causal && kk > i) && !(pad_row != nullptr && pad_row[kk]);
                okmask |= (unsigned long long)(ok ? 1u : 0u) << kk;
            }
        }
        const Rng rng = make_rng(a.rng_state, a.drop_p);
        if (rng.p > 0.f && valid) {
            keepmask = 0ull;
#pragma unroll
            for (int g = 0; g < LKP / 8; ++g)
                if (g * 8 < a.Lk)
                    keepmask |= (unsigned long long)dropout_bits8(rng, a.site, (unsigned long long)prow * 8ull + (unsigned long long)g) << (8 * g);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");          // bias_s complete (the four epilogue warps only)
        // ---- epilogue 1: projection accumulators -> bf16 operand tiles (+ saved q/k/v rows) ----
        mbar_wait(ph0 + 0, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (trace && threadIdx.x == 64) trace[103] = clock64();
        __nv_bfloat16* grow = (valid && a.q_out) ? a.q_out + ((long long)b * a.Lq + i) * a.q_ld + h * DH : nullptr;
#pragma unroll 1
        for (int sec = 0; sec < NSEC; ++sec) {
            const uint32_t tile = sec == 0 ? q_tile : (sec == 1 ? k_tile : v_tile);
#pragma unroll 1
            for (int c32 = 0; c32 < DH / 32; ++c32) {
                uint32_t rg[32];
                tmem_ld16(lane_addr + (uint32_t)(sec * DH + c32 * 32), rg);
                tmem_ld16(lane_addr + (uint32_t)(sec * DH + c32 * 32 + 16), rg + 16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const float4* bias4 = reinterpret_cast<const float4*>(bias_s + sec * DH + c32 * 32);
                uint32_t pk[16];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 bv = bias4[g];
                    pk[2 * g] = pack_bf16(__uint_as_float(rg[4 * g]) + bv.x, __uint_as_float(rg[4 * g + 1]) + bv.y);
                    pk[2 * g + 1] = pack_bf16(__uint_as_float(rg[4 * g + 2]) + bv.z, __uint_as_float(rg[4 * g + 3]) + bv.w);
                }
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    sts128(tile_addr(tile, r, c32 * 4 + ch), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
                    if (grow)
                        *reinterpret_cast<uint4*>(grow + (long long)sec * d + c32 * 32 + ch * 8) =
                            make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
                }
            }
        }
        fence_async_smem();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(ph0 + 8);
        if (trace && threadIdx.x == 64) trace[104] = clock64();
        // ---- softmax over this row's sequence ----
        mbar_wait(ph0 + 16, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (trace && threadIdx.x == 64) trace[105] = clock64();
        constexpr int SPW = LQP >= 32 ? 1 : 32 / LQP;         // sequences per warp (1, or 2 when sequences are 16 rows)
        constexpr int NLD = SPW * LKP;                        // score columns the warp loads (16 .. 64)
        float sc[LKP];
        {
            uint32_t rg[NLD];
            const int jw = LQP >= 32 ? (q * 32) / LQP : q * SPW;   // first sequence of this warp's 32 rows
            const int jq = jw < NB ? jw : 0;                   // warps past the group's last sequence read (and discard) block 0
            const uint32_t col0 = kColS + (uint32_t)(jq * LKP);
#pragma unroll
            for (int c = 0; c < NLD / 16; ++c) tmem_ld16(lane_addr + col0 + 16u * c, rg + 16 * c);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const bool upper = SPW == 2 && lane >= 16;
#pragma unroll
            for (int kk = 0; kk < LKP; ++kk) sc[kk] = __uint_as_float(SPW == 2 ? (upper ? rg[(NLD / 2 + kk) % NLD] : rg[kk]) : rg[kk]);
        }
        // softmax in the exp2 domain: exp(scale * (s - max)) = exp2(scale * log2(e) * (s - max)); one MUFU per key
        const float sl2 = a.scale * 1.4426950408889634f;
        float mx = -INFINITY;
#pragma unroll
        for (int kk = 0; kk < LKP; ++kk) {
            sc[kk] = ((okmask >> kk) & 1ull) ? sc[kk] * sl2 : -INFINITY;
            mx = fmaxf(mx, sc[kk]);
        }
        float sum = 0.f;
#pragma unroll
        for (int kk = 0; kk < LKP; ++kk) {
            sc[kk] = ((okmask >> kk) & 1ull) ? exp2f(sc[kk] - mx) : 0.f;
            sum += sc[kk];
        }
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        if (valid && a.probs) {
            float* pr = a.probs + prow * a.Lk;
#pragma unroll
            for (int kk = 0; kk < LKP; ++kk)
                if (kk < a.Lk) pr[kk] = sc[kk] * inv;
        }
        const float inv_keep = inv * rng.inv_keep;
#pragma unroll
        for (int kk = 0; kk < LKP; ++kk) sc[kk] = ((keepmask >> kk) & 1ull) ? sc[kk] * inv_keep : 0.f;
        // probability row -> A tile (K-major over the NK keys of the group): zero outside this sequence's LKP columns
#pragma unroll
        for (int cidx = 0; cidx < NK / 8; ++cidx) sts128(tile_addr(p_tile, r, cidx), 0u, 0u, 0u, 0u);
        if (valid) {
#pragma unroll
            for (int g = 0; g < LKP / 8; ++g)
                sts128(tile_addr(p_tile, r, j * (LKP / 8) + g), pack_bf16(sc[g * 8], sc[g * 8 + 1]), pack_bf16(sc[g * 8 + 2], sc[g * 8 + 3]),
                       pack_bf16(sc[g * 8 + 4], sc[g * 8 + 5]), pack_bf16(sc[g * 8 + 6], sc[g * 8 + 7]));
        }
        fence_async_smem();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(ph0 + 24);
        if (trace && threadIdx.x == 64) trace[106] = clock64();
        // ---- epilogue 2: attention output rows ----
        mbar_wait(ph0 + 32, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (trace && threadIdx.x == 64) trace[107] = clock64();
        __nv_bfloat16* orow = valid ? a.o + ((long long)b * a.Lq + i) * d + h * DH : nullptr;
#pragma unroll 1
        for (int c32 = 0; c32 < DH / 32; ++c32) {
            uint32_t rg[32];
            tmem_ld16(lane_addr + kColO + (uint32_t)(c32 * 32), rg);
            tmem_ld16(lane_addr + kColO + (uint32_t)(c32 * 32 + 16), rg + 16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (orow) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t w[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) w[t] = pack_bf16(__uint_as_float(rg[ch * 8 + 2 * t]), __uint_as_float(rg[ch * 8 + 2 * t + 1]));
                    *reinterpret_cast<uint4*>(orow + c32 * 32 + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (trace && threadIdx.x == 64) trace[108] = clock64();
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
    if (trace && threadIdx.x == 0) trace[109] = clock64();
}

// ---------------------------------------------------------------------------------------------------------------------
// Attention backward on tcgen05.  One CTA per (group of NB sequences, head); Q, K, V, dO tiles of the head come in by
// TMA (3-D boxes pad every sequence to LQP / LKP rows), then
//   S  = Q K^T,  dP = dO V^T                      (K-major operands)         -> TMEM
//   thread per query row: P (same masks, same exp2 softmax, same Philox keep mask as forward), dropped P and
//   dS = P o (dP_dropped - rowsum(P o dP_dropped)) * scale, both as bf16 operand tiles (zero outside the sequence block)
//   dQ = dS K          (A K-major, B = the K tile read MN-major)
//   dK = dS^T Q,  dV = Pd^T dO   (A = the dS / Pd tile read MN-major, i.e. transposed for free)
//   thread per row: dq / dk / dv rows -> bf16 global; per-CTA column sums -> in-projection bias gradient partials
// A tile written with tile_addr() can be consumed K-major (rows = M/N index) or MN-major (rows = K index): the
// SWIZZLE_128B pattern is the same, only the descriptor differs.
// ---------------------------------------------------------------------------------------------------------------------
struct BwdArgs {
    int B, H, Lq, Lk;
    const unsigned char* key_pad;
    int causal;
    float scale;
    __nv_bfloat16* dq; long long dq_ld;
    __nv_bfloat16* dk; long long dk_ld;
    __nv_bfloat16* dv; long long dv_ld;
    float drop_p; const unsigned long long* rng_state; unsigned int site;
    float* dbias; float* dbias_part; unsigned int* dbias_cnt;
};

template <int DH, int LQP, int LKP>
__global__ void __launch_bounds__(192, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, BwdArgs a) {
    constexpr int NB = (128 / LQP) < (128 / LKP) ? (128 / LQP) : (128 / LKP);
    constexpr int NK = NB * LKP, NQ = NB * LQP;
    constexpr int KSTEPS = DH / 16, NBLK = (DH + 63) / 64;
    constexpr uint32_t kColS = 0, kColDP = 128, kColDQ = 256, kColDK = 352, kColDV = 0;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) unsigned long long bars[4];      // 0 tiles landed, 1 S/dP done, 2 P/dS tiles written, 3 gradients done
    __shared__ uint32_t tmem_base_slot;
    __shared__ bool is_last;

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b0 = blockIdx.x * NB;
    const uint32_t smem0 = smem_u32(smem), bar0 = smem_u32(&bars[0]);
    const uint32_t q_tile = smem0, k_tile = smem0 + kTile, v_tile = smem0 + 2 * kTile, do_tile = smem0 + 3 * kTile;
    const uint32_t pd_tile = smem0 + 4 * kTile, ds_tile = smem0 + 5 * kTile;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDO) : "memory");
        mbar_init(bar0 + 0, 1);
        mbar_init(bar0 + 8, 1);
        mbar_init(bar0 + 16, 128);
        mbar_init(bar0 + 24, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    pdl_wait();

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(bar0 + 0, (uint32_t)NBLK * 2u * ((uint32_t)NQ + (uint32_t)NK) * 128u);
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk) {
                tma_load_3d(q_tile + blk * kBlk, &tmQ, h * DH + blk * 64, 0, b0, bar0 + 0);
                tma_load_3d(k_tile + blk * kBlk, &tmK, h * DH + blk * 64, 0, b0, bar0 + 0);
                tma_load_3d(do_tile + blk * kBlk, &tmDO, h * DH + blk * 64, 0, b0, bar0 + 0);
                tma_load_3d(v_tile + blk * kBlk, &tmV, h * DH + blk * 64, 0, b0, bar0 + 0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        const uint32_t base_idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_M >> 4) << 24);
        mbar_wait(bar0 + 0, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
            const uint32_t idesc_s = base_idesc | ((uint32_t)(NK >> 3) << 17);
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const uint32_t off = (uint32_t)(ks >> 2) * kBlk + (uint32_t)(ks & 3) * 32u;
                umma_bf16(tmem_base + kColS, make_desc(q_tile + off, 16, 1024), make_desc(k_tile + off, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
            }
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const uint32_t off = (uint32_t)(ks >> 2) * kBlk + (uint32_t)(ks & 3) * 32u;
                umma_bf16(tmem_base + kColDP, make_desc(do_tile + off, 16, 1024), make_desc(v_tile + off, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
            }
            umma_commit(bar0 + 8);
        }
        __syncwarp();
        mbar_wait(bar0 + 16, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
            const uint32_t idesc_n = base_idesc | ((uint32_t)(DH >> 3) << 17);
            // dQ = dS K : A K-major over the keys, B = K tile MN-major (k-rows = keys)
#pragma unroll
            for (int ks = 0; ks < NK / 16; ++ks) {
                const uint32_t aoff = (uint32_t)(ks >> 2) * kBlk + (uint32_t)(ks & 3) * 32u;
                umma_bf16(tmem_base + kColDQ, make_desc(ds_tile + aoff, 16, 1024), make_desc(k_tile + (uint32_t)ks * 2048u, kBlk, 1024),
                          idesc_n | (1u << 16), ks > 0 ? 1u : 0u);
            }
            // dK = dS^T Q, dV = Pd^T dO : A = the tile read MN-major (k-rows = query rows), B = Q / dO tile MN-major.
            // Only the NQ query rows the TMA boxes filled take part (rows beyond them hold stale shared memory).
#pragma unroll
            for (int ks = 0; ks < NQ / 16; ++ks)
                umma_bf16(tmem_base + kColDK, make_desc(ds_tile + (uint32_t)ks * 2048u, kBlk, 1024), make_desc(q_tile + (uint32_t)ks * 2048u, kBlk, 1024),
                          idesc_n | (1u << 15) | (1u << 16), ks > 0 ? 1u : 0u);
#pragma unroll
            for (int ks = 0; ks < NQ / 16; ++ks)
                umma_bf16(tmem_base + kColDV, make_desc(pd_tile + (uint32_t)ks * 2048u, kBlk, 1024), make_desc(do_tile + (uint32_t)ks * 2048u, kBlk, 1024),
                          idesc_n | (1u << 15) | (1u << 16), ks > 0 ? 1u : 0u);
            umma_commit(bar0 + 24);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int j = r / LQP, i = r % LQP;
        const int b = b0 + j;
        const bool valid = j < NB && b < a.B && i < a.Lq;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const long long prow = ((long long)b * a.H + h) * a.Lq + i;
        unsigned long long okmask = 0ull, keepmask = ~0ull;
        if (valid) {
            const unsigned char* pad_row = a.key_pad ? a.key_pad + (long long)b * a.Lk : nullptr;
#pragma unroll
            for (int kk = 0; kk < LKP; ++kk) {
                const bool ok = kk < a.Lk && !(a.causal && kk > i) && !(pad_row != nullptr && pad_row[kk]);
                okmask |= (unsigned long long)(ok ? 1u : 0u) << kk;
            }
        }
        const Rng rng = make_rng(a.rng_state, a.drop_p);
        if (rng.p > 0.f && valid) {
            keepmask = 0ull;
#pragma unroll
            for (int g = 0; g < LKP / 8; ++g)
                if (g * 8 < a.Lk)
                    keepmask |= (unsigned long long)dropout_bits8(rng, a.site, (unsigned long long)prow * 8ull + (unsigned long long)g) << (8 * g);
        }
        mbar_wait(bar0 + 8, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        constexpr int SPW = LQP >= 32 ? 1 : 32 / LQP, NLD = SPW * LKP;
        float sc[LKP], dp[LKP];
        {
            uint32_t rg[NLD], rd[NLD];
            const int jw = LQP >= 32 ? (q * 32) / LQP : q * SPW;
            const int jq = jw < NB ? jw : 0;
#pragma unroll
            for (int c = 0; c < NLD / 16; ++c) {
                tmem_ld16(lane_addr + kColS + (uint32_t)(jq * LKP) + 16u * c, rg + 16 * c);
                tmem_ld16(lane_addr + kColDP + (uint32_t)(jq * LKP) + 16u * c, rd + 16 * c);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const bool upper = SPW == 2 && lane >= 16;
#pragma unroll
            for (int kk = 0; kk < LKP; ++kk) {
                sc[kk] = __uint_as_float(SPW == 2 ? (upper ? rg[(NLD / 2 + kk) % NLD] : rg[kk]) : rg[kk]);
                dp[kk] = __uint_as_float(SPW == 2 ? (upper ? rd[(NLD / 2 + kk) % NLD] : rd[kk]) : rd[kk]);
            }
        }
        const float sl2 = a.scale * 1.4426950408889634f;
        float mx = -INFINITY;
#pragma unroll
        for (int kk = 0; kk < LKP; ++kk) {
            sc[kk] = ((okmask >> kk) & 1ull) ? sc[kk] * sl2 : -INFINITY;
            mx = fmaxf(mx, sc[kk]);
        }
        float sum = 0.f;
#pragma unroll
        for (int kk = 0; kk < LKP; ++kk) {
            sc[kk] = ((okmask >> kk) & 1ull) ? exp2f(sc[kk] - mx) : 0.f;
            sum += sc[kk];
        }
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        float dsum = 0.f;
#pragma unroll
        for (int kk = 0; kk < LKP; ++kk) {
            const float keep = ((keepmask >> kk) & 1ull) ? rng.inv_keep : 0.f;
            sc[kk] *= inv;                                   // p
            dp[kk] = ((okmask >> kk) & 1ull) ? dp[kk] * keep : 0.f;   // gradient wrt p (through the dropout)
            dsum = fmaf(sc[kk], dp[kk], dsum);
        }
#pragma unroll
        for (int kk = 0; kk < LKP; ++kk) {
            const float keep = ((keepmask >> kk) & 1ull) ? rng.inv_keep : 0.f;
            dp[kk] = sc[kk] * (dp[kk] - dsum) * a.scale;     // dS
            sc[kk] *= keep;                                  // dropped probabilities
        }
#pragma unroll
        for (int cidx = 0; cidx < NK / 8; ++cidx) {
            sts128(tile_addr(pd_tile, r, cidx), 0u, 0u, 0u, 0u);
            sts128(tile_addr(ds_tile, r, cidx), 0u, 0u, 0u, 0u);
        }
        if (valid) {
#pragma unroll
            for (int g = 0; g < LKP / 8; ++g) {
                sts128(tile_addr(pd_tile, r, j * (LKP / 8) + g), pack_bf16(sc[g * 8], sc[g * 8 + 1]), pack_bf16(sc[g * 8 + 2], sc[g * 8 + 3]),
                       pack_bf16(sc[g * 8 + 4], sc[g * 8 + 5]), pack_bf16(sc[g * 8 + 6], sc[g * 8 + 7]));
                sts128(tile_addr(ds_tile, r, j * (LKP / 8) + g), pack_bf16(dp[g * 8], dp[g * 8 + 1]), pack_bf16(dp[g * 8 + 2], dp[g * 8 + 3]),
                       pack_bf16(dp[g * 8 + 4], dp[g * 8 + 5]), pack_bf16(dp[g * 8 + 6], dp[g * 8 + 7]));
            }
        }
        fence_async_smem();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(bar0 + 16);
        // ---- gradients: thread = query row for dq, = key row for dk / dv ----
        mbar_wait(bar0 + 24, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int jk = r / LKP, kk = r % LKP;
        const int bk = b0 + jk;
        const bool kvalid = r < NK && bk < a.B && kk < a.Lk;
        __nv_bfloat16* rows[3] = {
            valid ? a.dq + ((long long)b * a.Lq + i) * a.dq_ld + h * DH : nullptr,
            kvalid ? a.dk + ((long long)bk * a.Lk + kk) * a.dk_ld + h * DH : nullptr,
            kvalid ? a.dv + ((long long)bk * a.Lk + kk) * a.dv_ld + h * DH : nullptr};
        const uint32_t cols[3] = {kColDQ, kColDK, kColDV};
        float* stagef = reinterpret_cast<float*>(smem);      // [128][DH + 1] fp32 column-sum staging (tiles are dead by now)
        const int t = threadIdx.x - 64;
        const int Hd = a.H * DH;
#pragma unroll 1
        for (int sec = 0; sec < 3; ++sec) {
            const bool rowok = sec == 0 ? valid : kvalid;
#pragma unroll 1
            for (int c32 = 0; c32 < DH / 32; ++c32) {
                uint32_t rg[32];
                tmem_ld16(lane_addr + cols[sec] + (uint32_t)(c32 * 32), rg);
                tmem_ld16(lane_addr + cols[sec] + (uint32_t)(c32 * 32 + 16), rg + 16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (rows[sec]) {
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t w[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) w[u] = pack_bf16(__uint_as_float(rg[ch * 8 + 2 * u]), __uint_as_float(rg[ch * 8 + 2 * u + 1]));
                        *reinterpret_cast<uint4*>(rows[sec] + c32 * 32 + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                if (a.dbias) {
#pragma unroll
                    for (int u = 0; u < 32; ++u) stagef[r * (DH + 1) + c32 * 32 + u] = rowok ? __uint_as_float(rg[u]) : 0.f;
                }
            }
            if (a.dbias) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (t < DH) {
                    float s = 0.f;
#pragma unroll 8
                    for (int rr = 0; rr < 128; ++rr) s += stagef[rr * (DH + 1) + t];
                    a.dbias_part[(long long)blockIdx.x * 3 * Hd + sec * Hd + h * DH + t] = s;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (a.dbias) {
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (t == 0) {
                const unsigned int prev = atomicAdd(a.dbias_cnt + h, 1u);
                is_last = (prev == gridDim.x - 1);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (is_last) {
                __threadfence();
                for (int c = t; c < 3 * DH; c += 128) {
                    const long long col = (long long)(c / DH) * Hd + h * DH + (c % DH);
                    float s = 0.f;
                    for (unsigned int g = 0; g < gridDim.x; ++g) s += __ldcg(a.dbias_part + (long long)g * 3 * Hd + col);
                    a.dbias[col] = s;
                }
                if (t == 0) a.dbias_cnt[h] = 0u;
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int DH, int LQP, int LKP>
int launch_bwd_tc(const vct_attn_args* m, cudaStream_t st) {
    constexpr int NB = (128 / LQP) < (128 / LKP) ? (128 / LQP) : (128 / LKP);
    constexpr int smem = 6 * (int)kTile + 1024;
    static_assert(128 * (DH + 1) * 4 <= 6 * (int)kTile, "column-sum staging must fit in the dead tiles");
    const int d = m->H * DH;
    CUtensorMap tmQ, tmK, tmV, tmDO;
    auto make = [&](const void* ptr, long long ld, long long bs, int L, int LP, CUtensorMap* out) {
        const unsigned long long dims[3] = {(unsigned long long)d, (unsigned long long)L, (unsigned long long)m->B};
        const unsigned long long strides[2] = {(unsigned long long)ld * 2ull, (unsigned long long)bs * 2ull};
        const unsigned int box[3] = {64u, (unsigned)LP, (unsigned)NB};
        return get_tensor_map_3d(ptr, dims, strides, box, out);
    };
    const long long q_bs = m->q_bs ? m->q_bs : (long long)m->Lq * m->q_ld, k_bs = m->k_bs ? m->k_bs : (long long)m->Lk * m->k_ld;
    const long long v_bs = m->v_bs ? m->v_bs : (long long)m->Lk * m->v_ld, do_bs = m->do_bs ? m->do_bs : (long long)m->Lq * m->do_ld;
    if (int e = make(m->q, m->q_ld, q_bs, m->Lq, LQP, &tmQ)) return e;
    if (int e = make(m->k, m->k_ld, k_bs, m->Lk, LKP, &tmK)) return e;
    if (int e = make(m->v, m->v_ld, v_bs, m->Lk, LKP, &tmV)) return e;
    if (int e = make(m->d_o, m->do_ld, do_bs, m->Lq, LQP, &tmDO)) return e;
    auto kern = attn_bwd_tc_kernel<DH, LQP, LKP>;
    static bool once = false;
    if (!once) {
        VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        once = true;
    }
    BwdArgs a;
    a.B = m->B; a.H = m->H; a.Lq = m->Lq; a.Lk = m->Lk;
    a.key_pad = m->key_pad; a.causal = m->causal; a.scale = m->scale;
    a.dq = (__nv_bfloat16*)m->dq; a.dq_ld = m->dq_ld;
    a.dk = (__nv_bfloat16*)m->dk; a.dk_ld = m->dk_ld;
    a.dv = (__nv_bfloat16*)m->dv; a.dv_ld = m->dv_ld;
    a.drop_p = m->drop_p; a.rng_state = m->rng_state; a.site = m->site;
    a.dbias = m->dbias; a.dbias_part = m->dbias_partials; a.dbias_cnt = m->dbias_counters;
    dim3 grid((m->B + NB - 1) / NB, m->H);
    vct::launch(kern, grid, dim3(192), smem, st, tmQ, tmK, tmV, tmDO, a);
    return check_launch("vct_attn_bwd(tcgen05)");
}

bool fused_enabled() {
    static const bool on = [] {
        const char* e = getenv("VCT_FUSED_ATTN");
        return e == nullptr || e[0] != '0';
    }();
    return on;
}

template <int DH, int LQP, int LKP, bool CROSS>
int launch_fused(const vct_mha_args* m, int causal, cudaStream_t st) {
    constexpr int NSEC = CROSS ? 1 : 3;
    constexpr int kStage = (int)kABytes + NSEC * DH * 128;
    constexpr int NB = (128 / LQP) < (128 / LKP) ? (128 / LQP) : (128 / LKP);
    constexpr int smem = kStages * kStage + (CROSS ? 2 * (int)kTile : 0) + 1024;
    static_assert(kStages * kStage >= (CROSS ? 2 : 4) * (int)kTile, "operand tiles must fit in the drained ring");
    const int d = m->d;
    const int Lk = CROSS ? m->Lk : m->L;
    CUtensorMap tmX, tmW, tmKV;
    {
        const unsigned long long dims[3] = {(unsigned long long)d, (unsigned long long)m->L, (unsigned long long)m->B};
        const unsigned long long strides[2] = {(unsigned long long)d * 2ull, (unsigned long long)m->L * d * 2ull};
        const unsigned int box[3] = {(unsigned)BLOCK_K, (unsigned)LQP, (unsigned)NB};
        if (int e = get_tensor_map_3d(m->x, dims, strides, box, &tmX)) return e;
    }
    if (CROSS) {
        if (int e = get_tensor_map(m->w_in, d, d, d, BLOCK_K, DH, &tmW)) return e;
        const unsigned long long dims[3] = {2ull * d, (unsigned long long)Lk, (unsigned long long)m->B};
        const unsigned long long strides[2] = {4ull * d, (unsigned long long)Lk * d * 4ull};
        const unsigned int box[3] = {64u, (unsigned)LKP, (unsigned)NB};
        if (int e = get_tensor_map_3d(m->kv, dims, strides, box, &tmKV)) return e;
    } else {
        const unsigned long long dims[3] = {(unsigned long long)d, (unsigned long long)d, 3ull};
        const unsigned long long strides[2] = {(unsigned long long)d * 2ull, (unsigned long long)d * d * 2ull};
        const unsigned int box[3] = {(unsigned)BLOCK_K, (unsigned)DH, 3u};
        if (int e = get_tensor_map_3d(m->w_in, dims, strides, box, &tmW)) return e;
        tmKV = tmW;
    }
    auto kern = attn_fused_kernel<DH, LQP, LKP, CROSS>;
    static bool once = false;
    if (!once) {
        VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        once = true;
    }
    FusedArgs a;
    a.B = m->B; a.Lq = m->L; a.Lk = Lk; a.d = d; a.H = m->H; a.nb = NB;
    a.b_in = m->b_in; a.key_pad = CROSS ? nullptr : m->key_pad; a.causal = causal;
    a.scale = 1.0f / sqrtf((float)DH);
    a.q_out = (__nv_bfloat16*)m->qkv; a.q_ld = CROSS ? d : 3ll * d;
    a.o = (__nv_bfloat16*)m->o; a.probs = m->probs;
    a.drop_p = m->drop_p; a.rng_state = m->rng_state; a.site = m->site;
    a.trace = gemm_trace_ptr();
    dim3 grid((m->B + NB - 1) / NB, m->H);
    vct::launch(kern, grid, dim3(kThreads), smem, st, tmX, tmW, tmKV, a);
    return check_launch(CROSS ? "vct_attn_dec_cross_fwd(fused)" : "vct_attn_self_fwd(fused)");
}

bool common_ok(const vct_mha_args* m) {
    if (!fused_enabled() || m->dtype != VCT_BF16 || m->gemm_impl != VCT_GEMM_TCGEN05) return false;
    if (m->d % m->H != 0 || m->d % 64 != 0 || m->L < 1 || m->B < 1) return false;
    if ((reinterpret_cast<uintptr_t>(m->x) & 15) || (reinterpret_cast<uintptr_t>(m->w_in) & 15) ||
        (reinterpret_cast<uintptr_t>(m->o) & 15) || (m->qkv && (reinterpret_cast<uintptr_t>(m->qkv) & 15)) ||
        (m->b_in && (reinterpret_cast<uintptr_t>(m->b_in) & 15)))
        return false;
    return true;
}

}  // namespace

namespace vct {

// both return 0 when the fused kernel was launched, > 0 when the shape is not covered (the caller composes the
// projection GEMM with the stand-alone core), < 0 on error
int attn_fused_self(const vct_mha_args* m, int causal, cudaStream_t st) {
    if (!common_ok(m) || m->L > 64) return 1;
    const int dh = m->d / m->H;
    // sequences are padded to 16 / 32 / 64 rows: 8 / 4 / 2 of them share one 128-row tile (cfg 5's M = 33 runs on LQP = 64)
    if (dh == 96) {
        if (m->L <= 16) return launch_fused<96, 16, 16, false>(m, causal, st);
        if (m->L <= 32) return launch_fused<96, 32, 32, false>(m, causal, st);
        return launch_fused<96, 64, 64, false>(m, causal, st);
    }
    if (dh == 64) {
        if (m->L <= 16) return launch_fused<64, 16, 16, false>(m, causal, st);
        if (m->L <= 32) return launch_fused<64, 32, 32, false>(m, causal, st);
        return launch_fused<64, 64, 64, false>(m, causal, st);
    }
    return 1;
}

// attention backward on the tensor cores: 0 = launched, > 0 = shape not covered (caller runs the SIMT kernel), < 0 = error
int attn_bwd_tc(const vct_attn_args* m, cudaStream_t st) {
    static const bool on = [] { const char* e = getenv("VCT_ATTN_BWD_TC"); return e == nullptr || e[0] != '0'; }();
    if (!on || m->dtype != VCT_BF16 || m->Lq > 64 || m->Lk > 64 || (m->causal && m->Lq != m->Lk)) return 1;
    if (m->dq_bs || m->dk_bs || m->dv_bs) return 1;                                  // gradients are written densely
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (!al(m->q) || !al(m->k) || !al(m->v) || !al(m->d_o) || !al(m->dq) || !al(m->dk) || !al(m->dv)) return 1;
    if (m->q_ld % 8 || m->k_ld % 8 || m->v_ld % 8 || m->do_ld % 8 || m->dq_ld % 8 || m->dk_ld % 8 || m->dv_ld % 8) return 1;
    if (m->q_bs % 8 || m->k_bs % 8 || m->v_bs % 8 || m->do_bs % 8) return 1;
    // query / key paddings: the smallest covered pair (a 16-row query padding only exists together with 16-row keys)
    int lqp = m->Lq <= 16 ? 16 : (m->Lq <= 32 ? 32 : 64);
    const int lkp = m->Lk <= 16 ? 16 : (m->Lk <= 32 ? 32 : 64);
    if (lqp == 16 && lkp != 16) lqp = 32;
#define VCT_BWD_CASE(DH, A, B) if (m->dh == DH && lqp == A && lkp == B) return launch_bwd_tc<DH, A, B>(m, st);
    VCT_BWD_CASE(96, 16, 16) VCT_BWD_CASE(96, 32, 32) VCT_BWD_CASE(96, 32, 16) VCT_BWD_CASE(96, 32, 64)
    VCT_BWD_CASE(96, 64, 16) VCT_BWD_CASE(96, 64, 32) VCT_BWD_CASE(96, 64, 64)
    VCT_BWD_CASE(64, 16, 16) VCT_BWD_CASE(64, 32, 32) VCT_BWD_CASE(64, 32, 16) VCT_BWD_CASE(64, 32, 64)
    VCT_BWD_CASE(64, 64, 16) VCT_BWD_CASE(64, 64, 32) VCT_BWD_CASE(64, 64, 64)
#undef VCT_BWD_CASE
    return 1;
}

int attn_fused_cross(const vct_mha_args* m, cudaStream_t st) {
    if (!common_ok(m) || !m->kv_ready || m->kv == nullptr || (reinterpret_cast<uintptr_t>(m->kv) & 15)) return 1;
    if (m->L > 64 || m->L < 1 || m->Lk < 1 || m->Lk > 64) return 1;          // query rows are padded to 32 or 64 per sequence
    const int dh = m->d / m->H;
    const bool q64 = m->L > 32;
    if (dh == 96) {
        if (m->Lk <= 16) return q64 ? launch_fused<96, 64, 16, true>(m, 0, st) : launch_fused<96, 32, 16, true>(m, 0, st);
        if (m->Lk <= 32) return q64 ? launch_fused<96, 64, 32, true>(m, 0, st) : launch_fused<96, 32, 32, true>(m, 0, st);
        return q64 ? launch_fused<96, 64, 64, true>(m, 0, st) : launch_fused<96, 32, 64, true>(m, 0, st);
    }
    if (dh == 64) {
        if (m->Lk <= 16) return q64 ? launch_fused<64, 64, 16, true>(m, 0, st) : launch_fused<64, 32, 16, true>(m, 0, st);
        if (m->Lk <= 32) return q64 ? launch_fused<64, 64, 32, true>(m, 0, st) : launch_fused<64, 32, 32, true>(m, 0, st);
        return q64 ? launch_fused<64, 64, 64, true>(m, 0, st) : launch_fused<64, 32, 64, true>(m, 0, st);
    }
    return 1;
}

}  // namespace vct
