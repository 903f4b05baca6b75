// tcgen05 GEMM for sm_100a: bf16 operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a
// multi-stage shared-memory ring, tcgen05.mma (cta_group::1, M = 128, N = 64/128/256, K = 16 per
// instruction) issued by one elected thread, fp32 accumulators in TMEM, epilogue warps read TMEM with
// tcgen05.ld and apply the fused epilogue (gemm_epilogue.cuh).
//
//   warp 0 : TMA producer (one elected lane)
//   warp 1 : TMEM allocation + MMA issuer (one elected lane)
//   warps 2-5 : epilogue, one thread per accumulator row (TMEM lane)
//
// Operand layouts (vct_gemm): "trans = 0" operands are K-major (rows x K, K contiguous) and are loaded
// as one [rows x 64] box per stage; "trans = 1" operands are MN-major (K x rows, rows contiguous) and are
// loaded as rows/64 boxes of [64 K-rows x 64] per stage.  Both land in the canonical UMMA SWIZZLE_128B
// layouts, so forward (x W^T), dgrad (dY W) and wgrad (dY^T X) all run on the same kernel without any
// transposed copies.  Out-of-bounds rows / K tails are zero-filled by TMA.
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "gemm_epilogue.cuh"
#include "tc_common.cuh"

using namespace vct;

namespace {

constexpr int kThreads = 192;
constexpr int kSmemBudget = 200 * 1024;

template <int BLOCK_N, bool A_MN, bool B_MN, int STAGES, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K, Epilogue epi,
               float* __restrict__ splitk_ws, long long ldw) {
    constexpr uint32_t kBBytes = BLOCK_N * BLOCK_K * 2;
    constexpr uint32_t kStageBytes = kABytes + kBBytes;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) unsigned long long full_bar[STAGES];
    __shared__ __align__(8) unsigned long long empty_bar[STAGES];
    __shared__ __align__(8) unsigned long long tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.x * BLOCK_N;
    // split-K: blockIdx.z owns the k-blocks [kb0, kb0 + num_kb)
    const int total_kb = (K + BLOCK_K - 1) / BLOCK_K;
    const int kb_per = (total_kb + (int)gridDim.z - 1) / (int)gridDim.z;
    const int kb0 = (int)blockIdx.z * kb_per;
    const int num_kb = max(0, min(kb_per, total_kb - kb0));

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&tmem_full_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"((uint32_t)BLOCK_N)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    // PDL: barrier init, TMEM allocation and descriptor prefetch above overlapped the predecessor's tail; nothing
    // before this point reads or writes global memory that another kernel of the step produces
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
                const uint32_t full = smem_u32(&full_bar[s]);
                mbar_expect_tx(full, kStageBytes);
                const uint32_t sa = smem_u32(smem + (size_t)s * kStageBytes), sb = sa + kABytes;
                const int k0 = (kb0 + kb) * BLOCK_K;
                if (!A_MN) {
                    tma_load_2d(sa, &tmA, k0, m0, full);
                } else {
#pragma unroll
                    for (int c = 0; c < BLOCK_M / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, m0 + c * 64, k0, full);
                }
                if (!B_MN) {
                    tma_load_2d(sb, &tmB, k0, n0, full);
                } else {
#pragma unroll
                    for (int c = 0; c < BLOCK_N / 64; ++c) tma_load_2d(sb + c * 8192, &tmB, n0 + c * 64, k0, full);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, majors, N>>3, M>>4
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(smem_u32(&full_bar[s]), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + (size_t)s * kStageBytes), sb = sa + kABytes;
                // K-major: 8-row groups 1024 B apart, one MMA K-step = 32 B along the row
                // MN-major: 64-wide MN chunks 8192 B apart (LBO), 8-K-row groups 1024 B apart (SBO), K-step = 2048 B
                const uint64_t adesc = A_MN ? make_desc(sa, 8192, 1024) : make_desc(sa, 16, 1024);
                const uint64_t bdesc = B_MN ? make_desc(sb, 8192, 1024) : make_desc(sb, 16, 1024);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint64_t ak = adesc + (uint64_t)((A_MN ? 2048u : 32u) * k >> 4);
                    const uint64_t bk = bdesc + (uint64_t)((B_MN ? 2048u : 32u) * k >> 4);
                    umma_bf16(tmem_base, ak, bk, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(smem_u32(&empty_bar[s]));      // frees the smem slot when these MMAs retire
            }
            umma_commit(smem_u32(&tmem_full_bar));          // accumulator complete
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
        // phase 1: TMEM -> registers -> shared memory (the operand ring is free once the accumulator is complete)
        // phase 2: shared memory -> fused epilogue -> global, with the threads remapped so that a warp touches
        //          contiguous global memory (the TMEM layout gives each thread a ROW, which would make every
        //          global access a 16-byte piece of a different cache line)
        constexpr int RS = BLOCK_N + 4;                       // padded row stride (floats): conflict-free 16 B stores
        float* stage = reinterpret_cast<float*>(smem);
        const int q = warp & 3;
        const int row = q * 32 + lane;
        mbar_wait(smem_u32(&tmem_full_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 16; ++c) {
            uint32_t r[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16);
            if (num_kb > 0) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = 0u;        // empty K range (split-K tail): contributes zeros
            }
            uint4* dst = reinterpret_cast<uint4*>(stage + row * RS + c * 16);
#pragma unroll
            for (int g = 0; g < 4; ++g) dst[g] = make_uint4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");        // the 4 epilogue warps only
        const Rng rng = make_rng(epi.rng_state, ACT != VCT_ACT_NONE ? epi.drop_p : 0.f);
        const int t = threadIdx.x - 64;
        constexpr int CG = BLOCK_N / 8;                       // 8-column groups per row
#pragma unroll 2
        for (int id = t; id < BLOCK_M * CG; id += 128) {
            const int rr = id / CG, cg = id % CG;
            const int m = m0 + rr, n = n0 + cg * 8;
            if (m >= epi.M || n >= epi.N) continue;
            const float4 a0 = *reinterpret_cast<const float4*>(stage + rr * RS + cg * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(stage + rr * RS + cg * 8 + 4);
            if (splitk_ws != nullptr) {
                // raw partial sums; vct's split-K reduce kernel applies the epilogue
                float* w = splitk_ws + ((long long)blockIdx.z * epi.M + m) * ldw + n;
                *reinterpret_cast<float4*>(w) = a0;
                *reinterpret_cast<float4*>(w + 4) = a1;
            } else {
                epilogue_store8<ACT>(epi, rng, m, n, a0, a1);
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BLOCK_N) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// Persistent variant for GEMMs with several tiles per SM (the generator forward / weight gradient):
// grid = #SMs, every CTA walks tiles round-robin; two 256-column TMEM accumulators let the epilogue of tile i
// (TMEM -> smem staging -> coalesced global stores, in 64-column chunks) overlap the TMA/MMA main loop of
// tile i+1, and the barrier / TMEM / descriptor setup is paid once per CTA instead of once per tile.
// ---------------------------------------------------------------------------------------------------
constexpr int P_BN = 256, P_STAGES = 3, P_CH = 64, P_RS = P_CH + 4;
constexpr uint32_t kPBBytes = P_BN * BLOCK_K * 2, kPStage = kABytes + kPBBytes;
constexpr int kPSmem = P_STAGES * (int)kPStage + BLOCK_M * P_RS * 4 + 1024;

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K,
                          Epilogue epi, int tiles_m, int tiles_n) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* stage = reinterpret_cast<float*>(smem + P_STAGES * kPStage);
    __shared__ __align__(8) unsigned long long full_bar[P_STAGES];
    __shared__ __align__(8) unsigned long long empty_bar[P_STAGES];
    __shared__ __align__(8) unsigned long long tmem_full_bar[2];
    __shared__ __align__(8) unsigned long long tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    const int num_tiles = tiles_m * tiles_n;
    // n-fastest tile order: CTAs running at the same time write adjacent column ranges of the SAME output rows
    // (DRAM-page-friendly stores; the operands are L2-resident either way)
    const bool m_fast = false;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < P_STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&tmem_full_bar[a]), 1);
            mbar_init(smem_u32(&tmem_empty_bar[a]), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    // PDL: barrier init, TMEM allocation and descriptor prefetch above overlapped the predecessor's tail; nothing
    // before this point reads or writes global memory that another kernel of the step produces
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mt = m_fast ? tile % tiles_m : tile / tiles_n, nt = m_fast ? tile / tiles_m : tile % tiles_n;
                const int m0 = mt * BLOCK_M, n0 = nt * P_BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const uint32_t s = it % P_STAGES, ph = (it / P_STAGES) & 1u;
                    mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
                    const uint32_t full = smem_u32(&full_bar[s]);
                    mbar_expect_tx(full, kPStage);
                    const uint32_t sa = smem_u32(smem + (size_t)s * kPStage), sb = sa + kABytes;
                    const int k0 = kb * BLOCK_K;
                    if (!A_MN) {
                        tma_load_2d(sa, &tmA, k0, m0, full);
                    } else {
#pragma unroll
                        for (int c = 0; c < BLOCK_M / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, m0 + c * 64, k0, full);
                    }
                    if (!B_MN) {
                        tma_load_2d(sb, &tmB, k0, n0, full);
                    } else {
#pragma unroll
                        for (int c = 0; c < P_BN / 64; ++c) tma_load_2d(sb + c * 8192, &tmB, n0 + c * 64, k0, full);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(P_BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            uint32_t it = 0, ti = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
                const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
                mbar_wait(smem_u32(&tmem_empty_bar[acc]), aph ^ 1u);     // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)P_BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const uint32_t s = it % P_STAGES, ph = (it / P_STAGES) & 1u;
                    mbar_wait(smem_u32(&full_bar[s]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(smem + (size_t)s * kPStage), sb = sa + kABytes;
                    const uint64_t adesc = A_MN ? make_desc(sa, 8192, 1024) : make_desc(sa, 16, 1024);
                    const uint64_t bdesc = B_MN ? make_desc(sb, 8192, 1024) : make_desc(sb, 16, 1024);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t ak = adesc + (uint64_t)((A_MN ? 2048u : 32u) * k >> 4);
                        const uint64_t bk = bdesc + (uint64_t)((B_MN ? 2048u : 32u) * k >> 4);
                        umma_bf16(d_tmem, ak, bk, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(smem_u32(&empty_bar[s]));
                }
                umma_commit(smem_u32(&tmem_full_bar[acc]));
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int t = threadIdx.x - 64;
        const Rng rng = make_rng(nullptr, 0.f);
        uint32_t ti = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
            const int mt = m_fast ? tile % tiles_m : tile / tiles_n, nt = m_fast ? tile / tiles_m : tile % tiles_n;
            const int m0 = mt * BLOCK_M, n0 = nt * P_BN;
            const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
            mbar_wait(smem_u32(&tmem_full_bar[acc]), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int ch = 0; ch < P_BN / P_CH; ++ch) {
#pragma unroll
                for (int c = 0; c < P_CH / 16; ++c) {
                    uint32_t r[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)P_BN + (uint32_t)(ch * P_CH + c * 16);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr)
                        : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    uint4* dst = reinterpret_cast<uint4*>(stage + row * P_RS + c * 16);
#pragma unroll
                    for (int g = 0; g < 4; ++g) dst[g] = make_uint4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
                }
                if (ch == P_BN / P_CH - 1) {
                    // every TMEM read of this accumulator has completed: hand it back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc])) : "memory");
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                constexpr int CG = P_CH / 8;
#pragma unroll 2
                for (int id = t; id < BLOCK_M * CG; id += 128) {
                    const int rr = id / CG, cg = id % CG;
                    const int m = m0 + rr, n = n0 + ch * P_CH + cg * 8;
                    if (m < epi.M && n < epi.N) {
                        const float4 a0 = *reinterpret_cast<const float4*>(stage + rr * P_RS + cg * 8);
                        const float4 a1 = *reinterpret_cast<const float4*>(stage + rr * P_RS + cg * 8 + 4);
                        epilogue_store8<VCT_ACT_NONE>(epi, rng, m, n, a0, a1);
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");       // staging buffer is reused by the next chunk
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// split-K second pass: sum the partial tiles in a fixed order and apply the fused epilogue
template <int ACT>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long ldw, Epilogue epi) {
    pdl_launch_dependents();
    pdl_wait();
    const int cg_per_row = (epi.N + 7) / 8;
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)epi.M * cg_per_row) return;
    const int m = (int)(id / cg_per_row), n = (int)(id % cg_per_row) * 8;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    for (int z = 0; z < splits; ++z) {
        const float* w = ws + ((long long)z * epi.M + m) * ldw + n;
        const float4 b0 = *reinterpret_cast<const float4*>(w), b1 = *reinterpret_cast<const float4*>(w + 4);
        a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
        a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
    }
    const Rng rng = make_rng(epi.rng_state, ACT != VCT_ACT_NONE ? epi.drop_p : 0.f);
    epilogue_store8<ACT>(epi, rng, m, n, a0, a1);
}

// ---------------------------------------------------------------------------------------------------
// host: tensor maps (cached) and dispatch
// ---------------------------------------------------------------------------------------------------
template <int BLOCK_N, bool A_MN, bool B_MN, int ACT>
int launch_tile(const vct_gemm_args* a, const CUtensorMap& tmA, const CUtensorMap& tmB, int splits, cudaStream_t st) {
    constexpr int kStage = kABytes + BLOCK_N * BLOCK_K * 2;
    constexpr int STAGES = kSmemBudget / kStage > 8 ? 8 : kSmemBudget / kStage;
    constexpr int smem = STAGES * kStage + 1024;
    auto kern = gemm_tc_kernel<BLOCK_N, A_MN, B_MN, STAGES, ACT>;
    static bool once = false;
    if (!once) {
        VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        once = true;
    }
    dim3 grid((a->N + BLOCK_N - 1) / BLOCK_N, (a->M + BLOCK_M - 1) / BLOCK_M, splits);
    const Epilogue epi = make_epilogue(a);
    const long long ldw = ((long long)a->N + 7) / 8 * 8;
    vct::launch(kern, dim3(grid), dim3(kThreads), smem, st, tmA, tmB, a->K, epi, splits > 1 ? (float*)a->splitk_ws : nullptr, ldw);
    if (int e = check_launch("vct_gemm(tcgen05)")) return e;
    if (splits > 1) {
        const long long items = (long long)a->M * ((a->N + 7) / 8);
        vct::launch(splitk_reduce_kernel<ACT>, dim3((unsigned)((items + 255) / 256)), dim3(256), 0, st, (const float*)a->splitk_ws, splits, ldw, epi);
        return check_launch("vct_gemm(tcgen05 split-K reduce)");
    }
    return 0;
}

template <bool A_MN, bool B_MN>
int launch_persistent(const vct_gemm_args* a, const CUtensorMap& tmA, const CUtensorMap& tmB, cudaStream_t st) {
    auto kern = gemm_tc_persistent_kernel<A_MN, B_MN>;
    static bool once = false;
    if (!once) {
        VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
        once = true;
    }
    const int tiles_m = (a->M + BLOCK_M - 1) / BLOCK_M, tiles_n = (a->N + P_BN - 1) / P_BN;
    const int grid = tiles_m * tiles_n < kNumSMs ? tiles_m * tiles_n : kNumSMs;
    vct::launch(kern, dim3(grid), dim3(kThreads), kPSmem, st, tmA, tmB, a->K, make_epilogue(a), tiles_m, tiles_n);
    return check_launch("vct_gemm(tcgen05 persistent)");
}

template <int BLOCK_N>
int dispatch_major(const vct_gemm_args* a, const CUtensorMap& tmA, const CUtensorMap& tmB, int splits, cudaStream_t st) {
    // activation epilogues exist where the path uses them: GELU forward on x W1^T (K-major, K-major) and GELU
    // backward on dY W2 (K-major, MN-major); everything else carries the small plain epilogue
    if (a->act == VCT_ACT_GELU_FWD) {
        VCT_REQUIRE(!a->a_trans && !a->b_trans, "vct_gemm(tcgen05): GELU_FWD is built for a_trans = b_trans = 0");
        return launch_tile<BLOCK_N, false, false, VCT_ACT_GELU_FWD>(a, tmA, tmB, splits, st);
    }
    if (a->act == VCT_ACT_GELU_BWD) {
        VCT_REQUIRE(!a->a_trans, "vct_gemm(tcgen05): GELU_BWD is built for a_trans = 0");
        if (a->b_trans) return launch_tile<BLOCK_N, false, true, VCT_ACT_GELU_BWD>(a, tmA, tmB, splits, st);
        return launch_tile<BLOCK_N, false, false, VCT_ACT_GELU_BWD>(a, tmA, tmB, splits, st);
    }
    if (!a->a_trans && !a->b_trans) return launch_tile<BLOCK_N, false, false, VCT_ACT_NONE>(a, tmA, tmB, splits, st);
    if (!a->a_trans && a->b_trans) return launch_tile<BLOCK_N, false, true, VCT_ACT_NONE>(a, tmA, tmB, splits, st);
    if (a->a_trans && !a->b_trans) return launch_tile<BLOCK_N, true, false, VCT_ACT_NONE>(a, tmA, tmB, splits, st);
    return launch_tile<BLOCK_N, true, true, VCT_ACT_NONE>(a, tmA, tmB, splits, st);
}

}  // namespace

namespace vct {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct MapKey {
    const void* ptr; long long inner, outer, ld; int box_inner, box_outer;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
               box_outer == o.box_outer;
    }
};
struct MapHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        auto mix = [&h](long long v) { h ^= std::hash<long long>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
        mix(k.inner); mix(k.outer); mix(k.ld); mix(k.box_inner); mix(k.box_outer);
        return h;
    }
};

// 2-D bf16 tensor: `inner` contiguous elements, `outer` rows `ld` elements apart; box [box_inner x box_outer]
int get_tensor_map(const void* ptr, long long inner, long long outer, long long ld, int box_inner, int box_outer, CUtensorMap* out) {
    static std::unordered_map<MapKey, CUtensorMap, MapHash> cache;
    static std::mutex mu;
    MapKey key{ptr, inner, outer, ld, box_inner, box_outer};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return 0; }
    }
    EncodeTiledFn enc = get_encode();
    VCT_REQUIRE(enc != nullptr, "vct_gemm(tcgen05): cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VCT_REQUIRE(r == CUDA_SUCCESS, "vct_gemm(tcgen05): cuTensorMapEncodeTiled failed (%d) ptr=%p inner=%lld outer=%lld ld=%lld",
                (int)r, ptr, inner, outer, ld);
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
    return 0;
}


int gemm_tcgen05(const vct_gemm_args* a, cudaStream_t st) {
    VCT_REQUIRE(a->a_dtype == VCT_BF16, "vct_gemm(tcgen05): operands must be bf16");
    VCT_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "vct_gemm(tcgen05): lda/ldb must be multiples of 8 elements (TMA 16-byte strides)");
    VCT_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0,
                "vct_gemm(tcgen05): operands must be 16-byte aligned");
    // Tile width and split-K factor from a small cost model (cycles): these GEMMs are bounded by the fixed cost of
    // a CTA (launch, TMEM alloc, first TMA round trip, epilogue ~ 9k cycles) and by L2 -> SM operand traffic
    // (~42 B/clk/SM when all SMs pull), not by the tensor pipe, so fewer / fuller waves win.
    const long long tiles_m = (a->M + BLOCK_M - 1) / BLOCK_M;
    const int total_kb = (a->K + BLOCK_K - 1) / BLOCK_K;
    const long long ldw = ((long long)a->N + 7) / 8 * 8;
    int bn = 64, splits = 1;
    double best = 1e30;
    // several 128x256 tiles per SM: the persistent kernel (epilogue overlapped with the next tile's main loop)
    const bool persistent = tiles_m * ((a->N + 255) / 256) >= 2 * kNumSMs && a->act == VCT_ACT_NONE &&
                            getenv("VCT_NO_PERSISTENT") == nullptr;
    if (persistent) { bn = 256; splits = 1; best = 0.0; }
    for (int cand : {256, 128, 64}) {
        const long long tiles = tiles_m * ((a->N + cand - 1) / cand);
        const double cyc_kb = fmax(2.0 * cand, (16384.0 + 128.0 * cand) / 42.0);
        for (int sp : {1, 2, 3, 4, 5, 6, 8}) {
            if (sp > 1 && (a->splitk_ws == nullptr || total_kb < 16 * sp ||
                           (long long)sp * a->M * ldw > a->splitk_ws_floats)) continue;
            const long long ctas = tiles * sp;
            const double waves = (double)((ctas + kNumSMs - 1) / kNumSMs);
            const int kbs = (total_kb + sp - 1) / sp;
            double t = waves * (9000.0 + kbs * cyc_kb);
            if (sp > 1) t += 8000.0 + (double)sp * a->M * a->N * 4.0 / 3000.0;
            if (t < best) { best = t; bn = cand; splits = sp; }
        }
    }
    CUtensorMap tmA, tmB;
    if (!a->a_trans) { if (int e = get_tensor_map(a->A, a->K, a->M, a->lda, BLOCK_K, BLOCK_M, &tmA)) return e; }
    else             { if (int e = get_tensor_map(a->A, a->M, a->K, a->lda, 64, BLOCK_K, &tmA)) return e; }
    if (!a->b_trans) { if (int e = get_tensor_map(a->B, a->K, a->N, a->ldb, BLOCK_K, bn, &tmB)) return e; }
    else             { if (int e = get_tensor_map(a->B, a->N, a->K, a->ldb, 64, BLOCK_K, &tmB)) return e; }
    if (bn == 256 && splits == 1 && a->act == VCT_ACT_NONE && persistent) {
        if (!a->a_trans && !a->b_trans) return launch_persistent<false, false>(a, tmA, tmB, st);
        if (!a->a_trans && a->b_trans) return launch_persistent<false, true>(a, tmA, tmB, st);
        if (a->a_trans && !a->b_trans) return launch_persistent<true, false>(a, tmA, tmB, st);
        return launch_persistent<true, true>(a, tmA, tmB, st);
    }
    if (bn == 256) return dispatch_major<256>(a, tmA, tmB, splits, st);
    if (bn == 128) return dispatch_major<128>(a, tmA, tmB, splits, st);
    return dispatch_major<64>(a, tmA, tmB, splits, st);
}

}  // namespace vct
