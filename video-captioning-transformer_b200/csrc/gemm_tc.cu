// tcgen05 GEMM for sm_100a: bf16 operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a
// multi-stage shared-memory ring, tcgen05.mma (cta_group::1, M = 128, N = 64/128/256, K = 16 per
// instruction) issued by one elected thread, fp32 accumulators in TMEM, epilogue warps read TMEM with
// tcgen05.ld and apply the fused epilogue (gemm_epilogue.cuh).
//
//   warp 0 : TMA producer of the A operand (one elected lane)
//   warp 1 : TMEM allocation + MMA issuer (one elected lane)
//   warps 2-5 : epilogue, one thread per accumulator row (TMEM lane)
//   warp 6 : TMA producer of the B operand (one elected lane)
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

constexpr int kThreads = 224;         // warp 0: A producer, 1: MMA issuer, 2-5: epilogue, 6: B producer
constexpr int kSmemBudget = 200 * 1024;
constexpr int kSmemBudgetLow = 100 * 1024;       // two CTAs per SM (227 KB / 2 minus static smem and slack)
int g_tune_bn = 0, g_tune_splits = 0, g_tune_low = 0;   // vct_gemm_tune
long long* g_trace = nullptr;                           // vct_gemm_trace

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
    long long* const trace = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? epi.trace : nullptr;
    if (trace && threadIdx.x == 0) trace[0] = clock64();
    // split-K: blockIdx.z owns the k-blocks [kb0, kb0 + num_kb)
    const int total_kb = (K + BLOCK_K - 1) / BLOCK_K;
    const int kb_per = (total_kb + (int)gridDim.z - 1) / (int)gridDim.z;
    const int kb0 = (int)blockIdx.z * kb_per;
    const int num_kb = max(0, min(kb_per, total_kb - kb0));

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 2);             // one arrive.expect_tx per producer thread
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
    if (trace && threadIdx.x == 0) trace[1] = clock64();
    // PDL: barrier init, TMEM allocation and descriptor prefetch above overlapped the predecessor's tail; nothing
    // before this point reads or writes global memory that another kernel of the step produces
    pdl_wait();
    if (trace && threadIdx.x == 0) trace[2] = clock64();

    // The producer and the MMA issuer are single threads: their loops are bound by instruction latency, not by
    // the tensor pipe or L2 (measured: 536 cycles per k-block with a rolled loop = 4 tcgen05.mma of 32..128 cycles
    // each).  The k-block loop is therefore unrolled by the ring depth: stage addresses, descriptors and barrier
    // addresses become immediates, and the barrier probes take the single-instruction fast path.
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
    if (warp == 0 || warp == 6) {
        {
            // ===== TMA producers: warp 0 loads the A tiles, warp 6 the B tiles =====
            // (a single thread needs ~200 cycles per cp.async.bulk.tensor it issues, measured; with one thread per
            // operand the ring fills twice as fast and the two loads of a k-block are in flight together)
            const bool is_a = warp == 0;
            uint32_t ph = 1u;                                  // parity of "slot free": passes on the fresh barriers
            int k0 = kb0 * BLOCK_K;
            for (int kbase = 0; kbase < num_kb; kbase += STAGES, ph ^= 1u) {
#pragma unroll
                for (int s = 0; s < STAGES; ++s) {
                    if (kbase + s < num_kb) {
                        mbar_wait_fast(empty0 + 8u * s, ph);
#ifdef VCT_GEMM_TRACE_KB
                        if (trace && lane == 0 && kbase + s < 20) trace[(is_a ? 16 : 36) + kbase + s] = clock64();
#endif
                        const uint32_t full = full0 + 8u * s;
                        const uint32_t sa = smem0 + (uint32_t)s * kStageBytes, sb = sa + kABytes;
                        if (elect_one()) {
                            if (is_a) {
                                mbar_expect_tx(full, kABytes);
                                if (!A_MN) {
                                    tma_load_2d(sa, &tmA, k0, m0, full);
                                } else {
#pragma unroll
                                    for (int c = 0; c < BLOCK_M / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, m0 + c * 64, k0, full);
                                }
                            } else {
                                mbar_expect_tx(full, kBBytes);
                                if (!B_MN) {
                                    tma_load_2d(sb, &tmB, k0, n0, full);
                                } else {
#pragma unroll
                                    for (int c = 0; c < BLOCK_N / 64; ++c) tma_load_2d(sb + c * 8192, &tmB, n0 + c * 64, k0, full);
                                }
                            }
                        }
                        __syncwarp();
                        k0 += BLOCK_K;
                    }
                }
            }
        }
    } else if (warp == 1) {
        {
            // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
            // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, majors, N>>3, M>>4
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            // K-major: 8-row groups 1024 B apart, one MMA K-step = 32 B along the row
            // MN-major: 64-wide MN chunks 8192 B apart (LBO), 8-K-row groups 1024 B apart (SBO), K-step = 2048 B
            const uint64_t adesc0 = A_MN ? make_desc(smem0, 8192, 1024) : make_desc(smem0, 16, 1024);
            const uint64_t bdesc0 = B_MN ? make_desc(smem0 + kABytes, 8192, 1024) : make_desc(smem0 + kABytes, 16, 1024);
            uint32_t ph = 0u, accum = 0u;
            for (int kbase = 0; kbase < num_kb; kbase += STAGES, ph ^= 1u) {
#pragma unroll
                for (int s = 0; s < STAGES; ++s) {
                    if (kbase + s < num_kb) {
                        mbar_wait_fast(full0 + 8u * s, ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef VCT_GEMM_TRACE_KB
                        if (trace && lane == 0 && kbase + s < 20) trace[56 + kbase + s] = clock64();
#endif
                        // stage s starts s * kStageBytes further (1024-byte multiple: no carry into the other fields)
                        const uint64_t adesc = adesc0 + (uint64_t)(((uint32_t)s * kStageBytes) >> 4);
                        const uint64_t bdesc = bdesc0 + (uint64_t)(((uint32_t)s * kStageBytes) >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                const uint64_t ak = adesc + (uint64_t)((A_MN ? 2048u : 32u) * k >> 4);
                                const uint64_t bk = bdesc + (uint64_t)((B_MN ? 2048u : 32u) * k >> 4);
                                umma_bf16(tmem_base, ak, bk, idesc, (accum | (uint32_t)k) != 0u ? 1u : 0u);
                            }
                            umma_commit(empty0 + 8u * s);      // frees the smem slot when these MMAs retire
                        }
                        __syncwarp();
                        accum = 1u;
#ifdef VCT_GEMM_TRACE_KB
                        if (trace && lane == 0 && kbase + s < 20) trace[76 + kbase + s] = clock64();
#endif
                    }
                }
            }
            if (elect_one()) {
                umma_commit(smem_u32(&tmem_full_bar));          // accumulator complete
                if (trace) trace[3] = clock64();
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
        // phase 1: TMEM -> registers -> shared memory (the operand ring is free once the accumulator is complete)
        // phase 2: shared memory -> fused epilogue -> global, with the threads remapped so that a warp instruction
        //          touches 512 contiguous bytes (the TMEM layout gives each thread a ROW, which would make every
        //          global access a 16-byte piece of a different cache line)
        constexpr int RS = BLOCK_N + 4;                       // padded row stride (floats): conflict-free 16 B stores
        const uint32_t stage = smem0;                         // shared-space address: explicit ld/st.shared, not generic
        const int q = warp & 3;
        const int row = q * 32 + lane;
        mbar_wait(smem_u32(&tmem_full_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (trace && threadIdx.x == 64) trace[4] = clock64();
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            if (num_kb > 0) {
                tmem_ld16(taddr, r);
                tmem_ld16(taddr + 16, r + 16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = 0u;        // empty K range (split-K tail): contributes zeros
            }
            const uint32_t dst = stage + (uint32_t)(row * RS + c * 32) * 4u;
#pragma unroll
            for (int g = 0; g < 8; ++g) sts128(dst + 16u * g, r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    // phase 2 is run by ALL warps: the producer and MMA warps have nothing left to do, and the elementwise part of the
    // epilogue (GELU, Philox, conversions) is bound by instruction issue of the few warps that execute it
    __syncthreads();
    if (trace && threadIdx.x == 64) trace[5] = clock64();
    {
        constexpr int RS = BLOCK_N + 4;
        const int t = threadIdx.x;
        if (splitk_ws != nullptr) {
            // raw partial sums; vct's split-K reduce kernel applies the epilogue
            epilogue_tile_partial<BLOCK_N, kThreads>(splitk_ws + (long long)blockIdx.z * epi.M * ldw, ldw, epi.M, epi.N, smem0, RS, m0, n0, t);
        } else if (ACT == VCT_ACT_NONE) {
            const Rng rng = make_rng(nullptr, 0.f);
            epilogue_tile_plain<BLOCK_N, kThreads>(epi, rng, smem0, RS, m0, n0, t);
        } else {
            const Rng rng = make_rng(epi.rng_state, epi.drop_p);
            epilogue_tile_act<ACT, BLOCK_N, kThreads>(epi, rng, smem0, RS, m0, n0, t);
        }
        if (trace && threadIdx.x == 64) trace[6] = clock64();
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BLOCK_N) : "memory");
    }
    if (trace && threadIdx.x == 0) trace[7] = clock64();
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
    const uint32_t stage = smem_u32(smem + P_STAGES * kPStage);   // epilogue staging, shared-space address
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
            mbar_init(smem_u32(&full_bar[s]), 2);
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

    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
    if (warp == 0 || warp == 6) {
        {
            const bool is_a = warp == 0;                       // warp 0 loads the A tiles, warp 6 the B tiles
            uint32_t s = 0, ph = 1u;                           // ring slot and its "free" parity
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mt = m_fast ? tile % tiles_m : tile / tiles_n, nt = m_fast ? tile / tiles_m : tile % tiles_n;
                const int m0 = mt * BLOCK_M, n0 = nt * P_BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_fast(empty0 + 8u * s, ph);
                    const uint32_t full = full0 + 8u * s;
                    const uint32_t sa = smem0 + s * kPStage, sb = sa + kABytes;
                    const int k0 = kb * BLOCK_K;
                    if (elect_one()) {
                        if (is_a) {
                            mbar_expect_tx(full, kABytes);
                            if (!A_MN) {
                                tma_load_2d(sa, &tmA, k0, m0, full);
                            } else {
#pragma unroll
                                for (int c = 0; c < BLOCK_M / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, m0 + c * 64, k0, full);
                            }
                        } else {
                            mbar_expect_tx(full, kPBBytes);
                            if (!B_MN) {
                                tma_load_2d(sb, &tmB, k0, n0, full);
                            } else {
#pragma unroll
                                for (int c = 0; c < P_BN / 64; ++c) tma_load_2d(sb + c * 8192, &tmB, n0 + c * 64, k0, full);
                            }
                        }
                    }
                    __syncwarp();
                    if (++s == P_STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(P_BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            const uint64_t adesc0 = A_MN ? make_desc(smem0, 8192, 1024) : make_desc(smem0, 16, 1024);
            const uint64_t bdesc0 = B_MN ? make_desc(smem0 + kABytes, 8192, 1024) : make_desc(smem0 + kABytes, 16, 1024);
            uint32_t s = 0, ph = 0u, ti = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
                const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
                mbar_wait_fast(smem_u32(&tmem_empty_bar[acc]), aph ^ 1u);     // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)P_BN;
                uint32_t accum = 0u;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_fast(full0 + 8u * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t adesc = adesc0 + (uint64_t)((s * kPStage) >> 4);
                    const uint64_t bdesc = bdesc0 + (uint64_t)((s * kPStage) >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            const uint64_t ak = adesc + (uint64_t)((A_MN ? 2048u : 32u) * k >> 4);
                            const uint64_t bk = bdesc + (uint64_t)((B_MN ? 2048u : 32u) * k >> 4);
                            umma_bf16(d_tmem, ak, bk, idesc, (accum | (uint32_t)k) != 0u ? 1u : 0u);
                        }
                        umma_commit(empty0 + 8u * s);
                    }
                    __syncwarp();
                    accum = 1u;
                    if (++s == P_STAGES) { s = 0; ph ^= 1u; }
                }
                if (elect_one()) umma_commit(smem_u32(&tmem_full_bar[acc]));
                __syncwarp();
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
                for (int c = 0; c < P_CH / 32; ++c) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)P_BN + (uint32_t)(ch * P_CH + c * 32);
                    tmem_ld16(taddr, r);
                    tmem_ld16(taddr + 16, r + 16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const uint32_t dst = stage + (uint32_t)(row * P_RS + c * 32) * 4u;
#pragma unroll
                    for (int g = 0; g < 8; ++g) sts128(dst + 16u * g, r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
                }
                if (ch == P_BN / P_CH - 1) {
                    // every TMEM read of this accumulator has completed: hand it back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc])) : "memory");
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                epilogue_tile_plain<P_CH, 128>(epi, rng, stage, P_RS, m0, n0 + ch * P_CH, t);
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

// ---------------------------------------------------------------------------------------------------
// Grouped persistent GEMM: up to kMaxGroup independent problems C_g[M_g, N_g] = A_g^T-stored x B_g^T-stored (both operands
// MN-major: the weight-gradient GEMMs dW = dY^T X of one transformer layer) in ONE launch.  The 128 x 256 tiles of all
// problems form one list that the persistent CTAs walk round-robin, with the same warp roles, operand ring and
// double-buffered TMEM accumulator as gemm_tc_persistent_kernel.  A layer's 5-7 weight gradients were 5-7 launches of
// 18-54 CTAs each (6-9 us apiece, mostly fixed cost); together they are ~240 tiles = 1.6 waves of one launch.
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxGroup = 8;
struct GroupProblem {
    int M, N, K, tiles_n, tile_start, vec_ok;
    float* C;
    long long ldc;
};
struct GroupArgs {
    int count, total_tiles;
    GroupProblem p[kMaxGroup];
    CUtensorMap tmA[kMaxGroup];
    CUtensorMap tmB[kMaxGroup];
};

__device__ __forceinline__ int group_of_tile(const GroupArgs& g, int tile) {
    int q = 0;
#pragma unroll
    for (int i = 1; i < kMaxGroup; ++i)
        if (i < g.count && tile >= g.p[i].tile_start) q = i;
    return q;
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_grouped_kernel(const __grid_constant__ GroupArgs g) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t stage = smem_u32(smem + P_STAGES * kPStage);   // epilogue staging, shared-space address
    __shared__ __align__(8) unsigned long long full_bar[P_STAGES];
    __shared__ __align__(8) unsigned long long empty_bar[P_STAGES];
    __shared__ __align__(8) unsigned long long tmem_full_bar[2];
    __shared__ __align__(8) unsigned long long tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = g.total_tiles;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < g.count; ++i) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmA[i]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmB[i]) : "memory");
        }
        for (int s = 0; s < P_STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 2);
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
    pdl_wait();

    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
    if (warp == 0 || warp == 6) {
        const bool is_a = warp == 0;                           // warp 0 loads the A tiles, warp 6 the B tiles
        uint32_t s = 0, ph = 1u;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int q = group_of_tile(g, tile);
            const GroupProblem& pr = g.p[q];
            const int local = tile - pr.tile_start;
            const int m0 = (local / pr.tiles_n) * BLOCK_M, n0 = (local % pr.tiles_n) * P_BN;
            const int num_kb = (pr.K + BLOCK_K - 1) / BLOCK_K;
            const CUtensorMap* tm = is_a ? &g.tmA[q] : &g.tmB[q];
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_fast(empty0 + 8u * s, ph);
                const uint32_t full = full0 + 8u * s;
                const uint32_t sa = smem0 + s * kPStage, sb = sa + kABytes;
                const int k0 = kb * BLOCK_K;
                if (elect_one()) {
                    if (is_a) {
                        mbar_expect_tx(full, kABytes);
#pragma unroll
                        for (int c = 0; c < BLOCK_M / 64; ++c) tma_load_2d(sa + c * 8192, tm, m0 + c * 64, k0, full);
                    } else {
                        mbar_expect_tx(full, kPBBytes);
#pragma unroll
                        for (int c = 0; c < P_BN / 64; ++c) tma_load_2d(sb + c * 8192, tm, n0 + c * 64, k0, full);
                    }
                }
                __syncwarp();
                if (++s == P_STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(P_BN >> 3) << 17) |
                               ((uint32_t)(BLOCK_M >> 4) << 24);
        const uint64_t adesc0 = make_desc(smem0, 8192, 1024);
        const uint64_t bdesc0 = make_desc(smem0 + kABytes, 8192, 1024);
        uint32_t s = 0, ph = 0u, ti = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
            const int q = group_of_tile(g, tile);
            const int num_kb = (g.p[q].K + BLOCK_K - 1) / BLOCK_K;
            const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
            mbar_wait_fast(smem_u32(&tmem_empty_bar[acc]), aph ^ 1u);     // epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem_base + acc * (uint32_t)P_BN;
            uint32_t accum = 0u;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_fast(full0 + 8u * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t adesc = adesc0 + (uint64_t)((s * kPStage) >> 4);
                const uint64_t bdesc = bdesc0 + (uint64_t)((s * kPStage) >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t ak = adesc + (uint64_t)(2048u * k >> 4);
                        const uint64_t bk = bdesc + (uint64_t)(2048u * k >> 4);
                        umma_bf16(d_tmem, ak, bk, idesc, (accum | (uint32_t)k) != 0u ? 1u : 0u);
                    }
                    umma_commit(empty0 + 8u * s);
                }
                __syncwarp();
                accum = 1u;
                if (++s == P_STAGES) { s = 0; ph ^= 1u; }
            }
            if (elect_one()) umma_commit(smem_u32(&tmem_full_bar[acc]));
            __syncwarp();
        }
    } else {
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const int t = threadIdx.x - 64;
        const Rng rng = make_rng(nullptr, 0.f);
        uint32_t ti = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
            const int q = group_of_tile(g, tile);
            const GroupProblem& pr = g.p[q];
            const int local = tile - pr.tile_start;
            const int m0 = (local / pr.tiles_n) * BLOCK_M, n0 = (local % pr.tiles_n) * P_BN;
            Epilogue epi;
            epi.M = pr.M; epi.N = pr.N;
            epi.C = pr.C; epi.c_dtype = VCT_F32; epi.ldc = pr.ldc;
            epi.C2 = nullptr; epi.c2_dtype = VCT_F32; epi.ldc2 = 0;
            epi.bias = nullptr; epi.row_table = nullptr; epi.row_period = 0;
            epi.addend = nullptr; epi.ld_addend = 0; epi.act = VCT_ACT_NONE;
            epi.aux = nullptr; epi.aux_dtype = VCT_F32; epi.ld_aux = 0;
            epi.drop_p = 0.f; epi.rng_state = nullptr; epi.site = 0u;
            epi.vec_ok = pr.vec_ok; epi.trace = nullptr;
            const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
            mbar_wait(smem_u32(&tmem_full_bar[acc]), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int ch = 0; ch < P_BN / P_CH; ++ch) {
#pragma unroll
                for (int c = 0; c < P_CH / 32; ++c) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + acc * (uint32_t)P_BN + (uint32_t)(ch * P_CH + c * 32);
                    tmem_ld16(taddr, r);
                    tmem_ld16(taddr + 16, r + 16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const uint32_t dst = stage + (uint32_t)(row * P_RS + c * 32) * 4u;
#pragma unroll
                    for (int gg = 0; gg < 8; ++gg) sts128(dst + 16u * gg, r[4 * gg], r[4 * gg + 1], r[4 * gg + 2], r[4 * gg + 3]);
                }
                if (ch == P_BN / P_CH - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc])) : "memory");
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                epilogue_tile_plain<P_CH, 128>(epi, rng, stage, P_RS, m0, n0 + ch * P_CH, t);
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
// LOW: half the shared-memory ring (<= ~100 KB) so that TWO CTAs are resident per SM: one CTA's prologue /
// epilogue (launch, TMEM alloc, first TMA round trip, TMEM -> smem -> global) overlaps the other's main loop.
// These GEMMs are bound by L2 -> SM operand traffic and per-CTA fixed cost, not by the tensor pipe, so the
// shallower ring costs nothing while the co-resident CTA hides ~9k cycles of fixed cost per tile.
template <int BLOCK_N, bool A_MN, bool B_MN, int ACT, bool LOW>
int launch_tile_impl(const vct_gemm_args* a, const CUtensorMap& tmA, const CUtensorMap& tmB, int splits, cudaStream_t st) {
    constexpr int kStage = kABytes + BLOCK_N * BLOCK_K * 2;
    constexpr int kBudget = LOW ? kSmemBudgetLow : kSmemBudget;
    constexpr int kStagesFit = kBudget / kStage > 8 ? 8 : kBudget / kStage;
    // the epilogue stages the [128 x BLOCK_N] fp32 tile (+4 floats of row padding) in the operand ring
    constexpr int kEpiBytes = BLOCK_M * (BLOCK_N + 4) * 4;
    constexpr int STAGES = kStagesFit * kStage >= kEpiBytes ? kStagesFit : (kEpiBytes + kStage - 1) / kStage;
    constexpr int smem = STAGES * kStage + 1024;
    auto kern = gemm_tc_kernel<BLOCK_N, A_MN, B_MN, STAGES, ACT>;
    static bool once = false;
    if (!once) {
        VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        once = true;
    }
    dim3 grid((a->N + BLOCK_N - 1) / BLOCK_N, (a->M + BLOCK_M - 1) / BLOCK_M, splits);
    Epilogue epi = make_epilogue(a);
    epi.trace = g_trace;
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

template <int BLOCK_N, bool A_MN, bool B_MN, int ACT>
int launch_tile(const vct_gemm_args* a, const CUtensorMap& tmA, const CUtensorMap& tmB, int splits, bool low, cudaStream_t st) {
    if constexpr (BLOCK_N <= 128) {
        if (low) return launch_tile_impl<BLOCK_N, A_MN, B_MN, ACT, true>(a, tmA, tmB, splits, st);
    }
    return launch_tile_impl<BLOCK_N, A_MN, B_MN, ACT, false>(a, tmA, tmB, splits, st);
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
int dispatch_major(const vct_gemm_args* a, const CUtensorMap& tmA, const CUtensorMap& tmB, int splits, bool low, cudaStream_t st) {
    // activation epilogues exist where the path uses them: GELU forward on x W1^T (K-major, K-major) and GELU
    // backward on dY W2 (K-major, MN-major); everything else carries the small plain epilogue
    if (a->act == VCT_ACT_GELU_FWD || a->act == VCT_ACT_GELU_FWD_F) {
        VCT_REQUIRE(!a->a_trans && !a->b_trans, "vct_gemm(tcgen05): GELU_FWD is built for a_trans = b_trans = 0");
        if (a->act == VCT_ACT_GELU_FWD_F) return launch_tile<BLOCK_N, false, false, VCT_ACT_GELU_FWD_F>(a, tmA, tmB, splits, low, st);
        return launch_tile<BLOCK_N, false, false, VCT_ACT_GELU_FWD>(a, tmA, tmB, splits, low, st);
    }
    if (a->act == VCT_ACT_MUL_AUX) {
        VCT_REQUIRE(!a->a_trans, "vct_gemm(tcgen05): MUL_AUX is built for a_trans = 0");
        if (a->b_trans) return launch_tile<BLOCK_N, false, true, VCT_ACT_MUL_AUX>(a, tmA, tmB, splits, low, st);
        return launch_tile<BLOCK_N, false, false, VCT_ACT_MUL_AUX>(a, tmA, tmB, splits, low, st);
    }
    if (a->act == VCT_ACT_GELU_BWD) {
        VCT_REQUIRE(!a->a_trans, "vct_gemm(tcgen05): GELU_BWD is built for a_trans = 0");
        if (a->b_trans) return launch_tile<BLOCK_N, false, true, VCT_ACT_GELU_BWD>(a, tmA, tmB, splits, low, st);
        return launch_tile<BLOCK_N, false, false, VCT_ACT_GELU_BWD>(a, tmA, tmB, splits, low, st);
    }
    if (!a->a_trans && !a->b_trans) return launch_tile<BLOCK_N, false, false, VCT_ACT_NONE>(a, tmA, tmB, splits, low, st);
    if (!a->a_trans && a->b_trans) return launch_tile<BLOCK_N, false, true, VCT_ACT_NONE>(a, tmA, tmB, splits, low, st);
    if (a->a_trans && !a->b_trans) return launch_tile<BLOCK_N, true, false, VCT_ACT_NONE>(a, tmA, tmB, splits, low, st);
    return launch_tile<BLOCK_N, true, true, VCT_ACT_NONE>(a, tmA, tmB, splits, low, st);
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


// tile override for tuning sweeps: bn in {64,128,256} (0 = cost model), split-K factor, low: 1 = two-CTA-per-SM ring,
// 0 = full ring, -1 = persistent kernel
void gemm_tune(int bn, int splits, int low) { g_tune_bn = bn; g_tune_splits = splits; g_tune_low = low; }
void gemm_trace(long long* dev_buf) { g_trace = dev_buf; }
long long* gemm_trace_ptr() { return g_trace; }

// 3-D bf16 tensor (dims innermost first, strides of dims 1 and 2 in bytes), SWIZZLE_128B, zero fill out of bounds
int get_tensor_map_3d(const void* ptr, const unsigned long long dims[3], const unsigned long long strides_bytes[2],
                      const unsigned int box[3], CUtensorMap* out) {
    struct Key3 {
        const void* ptr; unsigned long long d[3], s[2]; unsigned int b[3];
        bool operator==(const Key3& o) const {
            return ptr == o.ptr && d[0] == o.d[0] && d[1] == o.d[1] && d[2] == o.d[2] && s[0] == o.s[0] && s[1] == o.s[1] &&
                   b[0] == o.b[0] && b[1] == o.b[1] && b[2] == o.b[2];
        }
    };
    struct Hash3 {
        size_t operator()(const Key3& k) const {
            size_t h = std::hash<const void*>()(k.ptr);
            auto mix = [&h](unsigned long long v) { h ^= std::hash<unsigned long long>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
            for (int i = 0; i < 3; ++i) { mix(k.d[i]); mix(k.b[i]); }
            mix(k.s[0]); mix(k.s[1]);
            return h;
        }
    };
    static std::unordered_map<Key3, CUtensorMap, Hash3> cache;
    static std::mutex mu;
    Key3 key{ptr, {dims[0], dims[1], dims[2]}, {strides_bytes[0], strides_bytes[1]}, {box[0], box[1], box[2]}};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return 0; }
    }
    EncodeTiledFn enc = get_encode();
    VCT_REQUIRE(enc != nullptr, "tensor map: cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t gd[3] = {dims[0], dims[1], dims[2]};
    cuuint64_t gs[2] = {strides_bytes[0], strides_bytes[1]};
    cuuint32_t bx[3] = {box[0], box[1], box[2]};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gd, gs, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VCT_REQUIRE(r == CUDA_SUCCESS, "tensor map: cuTensorMapEncodeTiled(3d) failed (%d) ptr=%p dims=%llu,%llu,%llu box=%u,%u,%u", (int)r,
                ptr, dims[0], dims[1], dims[2], box[0], box[1], box[2]);
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
    return 0;
}

// vct_gemm_grouped: every problem must be a bf16 x bf16 -> fp32 GEMM with both operands MN-major (a_trans = b_trans = 1), no
// epilogue extras.  Returns > 0 when the group is not of that shape (the caller issues the GEMMs one by one).
int gemm_grouped(const vct_gemm_args* args, int count, cudaStream_t st) {
    if (count < 1 || count > kMaxGroup) return 1;
    static const bool on = [] { const char* e = getenv("VCT_GEMM_GROUPED"); return e == nullptr || e[0] != '0'; }();
    if (!on) return 1;
    GroupArgs g;
    memset(&g, 0, sizeof(g));
    g.count = count;
    int tiles = 0;
    for (int i = 0; i < count; ++i) {
        const vct_gemm_args* a = &args[i];
        if (a->impl != VCT_GEMM_TCGEN05 || a->a_dtype != VCT_BF16 || a->b_dtype != VCT_BF16 || !a->a_trans || !a->b_trans ||
            a->c_dtype != VCT_F32 || a->C2 || a->bias || a->row_table || a->addend || a->act != VCT_ACT_NONE)
            return 1;
        if (a->lda % 8 || a->ldb % 8 || (reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->B) & 15) ||
            (reinterpret_cast<uintptr_t>(a->C) & 15))
            return 1;
        GroupProblem& p = g.p[i];
        p.M = a->M; p.N = a->N; p.K = a->K;
        p.tiles_n = (a->N + P_BN - 1) / P_BN;
        p.tile_start = tiles;
        tiles += ((a->M + BLOCK_M - 1) / BLOCK_M) * p.tiles_n;
        p.C = reinterpret_cast<float*>(a->C);
        p.ldc = a->ldc;
        p.vec_ok = (a->ldc % 4 == 0) ? 1 : 0;
        if (int e = get_tensor_map(a->A, a->M, a->K, a->lda, 64, BLOCK_K, &g.tmA[i])) return e;
        if (int e = get_tensor_map(a->B, a->N, a->K, a->ldb, 64, BLOCK_K, &g.tmB[i])) return e;
    }
    g.total_tiles = tiles;
    static bool once = false;
    if (!once) {
        VCT_CUDA(cudaFuncSetAttribute(gemm_tc_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
        once = true;
    }
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
    vct::launch(gemm_tc_grouped_kernel, dim3(grid), dim3(kThreads), kPSmem, st, g);
    return check_launch("vct_gemm_grouped");
}

int gemm_tcgen05(const vct_gemm_args* a, cudaStream_t st) {
    VCT_REQUIRE(a->a_dtype == VCT_BF16, "vct_gemm(tcgen05): operands must be bf16");
    VCT_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "vct_gemm(tcgen05): lda/ldb must be multiples of 8 elements (TMA 16-byte strides)");
    VCT_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0,
                "vct_gemm(tcgen05): operands must be 16-byte aligned");
    // Tile width, ring flavour and split-K factor from a small cost model (cycles), fitted to tools/gemm_sweep.py on
    // B200 (regret 2.7 us over the 62 layer GEMMs of a step).  These GEMMs are bound by the fixed cost of a CTA
    // (launch, TMEM alloc, first TMA round trip, epilogue) and by the per-k-block issue cost of the producer / MMA
    // threads (~320 cycles + 1.2 per tile column), not by the tensor pipe, so fewer / fuller waves win; with the half
    // ring two CTAs share an SM (twice the slots per wave, each k-block ~1.86x slower when both are resident).
    const long long tiles_m = (a->M + BLOCK_M - 1) / BLOCK_M;
    const int total_kb = (a->K + BLOCK_K - 1) / BLOCK_K;
    const long long ldw = ((long long)a->N + 7) / 8 * 8;
    int bn = 64, splits = 1;
    bool low = false;
    double best = 1e30;
    // several 128x256 tiles per SM: the persistent kernel (epilogue overlapped with the next tile's main loop)
    const bool persistent = tiles_m * ((a->N + 255) / 256) >= 2 * kNumSMs && a->act == VCT_ACT_NONE &&
                            getenv("VCT_NO_PERSISTENT") == nullptr;
    if (persistent) { bn = 256; splits = 1; best = 0.0; }
    static const int force_low = getenv("VCT_GEMM_LOW") ? atoi(getenv("VCT_GEMM_LOW")) : 0;
    for (int cand : {256, 128, 64}) {
        if (force_low && cand == 256 && !persistent) continue;
        const long long tiles = tiles_m * ((a->N + cand - 1) / cand);
        const double cyc_kb = fmax(2.0 * cand, 316.5 + 1.215 * cand);
        for (int lo = (force_low && cand <= 128) ? 1 : 0; lo <= (cand <= 128 ? 1 : 0); ++lo) {
            for (int sp : {1, 2, 3, 4, 5, 6, 8}) {
                if (sp > 1 && (a->splitk_ws == nullptr || total_kb < 16 * sp ||
                               (long long)sp * a->M * ldw > a->splitk_ws_floats)) continue;
                const long long ctas = tiles * sp;
                const int slots = lo ? 2 * kNumSMs : kNumSMs;
                const double waves = (double)((ctas + slots - 1) / slots);
                const int kbs = (total_kb + sp - 1) / sp;
                double t = waves * (4716.0 + 26.7 * cand + kbs * cyc_kb * ((lo && ctas > kNumSMs) ? 1.856 : 1.0));
                if (sp > 1) t += 6348.0 + (double)sp * a->M * a->N * 4.0 / 3000.0;
                if (t < best) { best = t; bn = cand; splits = sp; low = lo != 0; }
            }
        }
    }
    bool use_persistent = persistent;
    if (g_tune_bn > 0) {                       // vct_gemm_tune override (tools/gemm_sweep.py)
        bn = g_tune_bn;
        splits = g_tune_splits > 0 ? g_tune_splits : 1;
        if (splits > 1 && (a->splitk_ws == nullptr || (long long)splits * a->M * ldw > a->splitk_ws_floats)) splits = 1;
        low = g_tune_low > 0 && bn <= 128;
        use_persistent = g_tune_low < 0;        // low = -1 selects the persistent kernel (bn must be 256)
    }
    CUtensorMap tmA, tmB;
    if (!a->a_trans) { if (int e = get_tensor_map(a->A, a->K, a->M, a->lda, BLOCK_K, BLOCK_M, &tmA)) return e; }
    else             { if (int e = get_tensor_map(a->A, a->M, a->K, a->lda, 64, BLOCK_K, &tmA)) return e; }
    if (!a->b_trans) { if (int e = get_tensor_map(a->B, a->K, a->N, a->ldb, BLOCK_K, bn, &tmB)) return e; }
    else             { if (int e = get_tensor_map(a->B, a->N, a->K, a->ldb, 64, BLOCK_K, &tmB)) return e; }
    if (bn == 256 && splits == 1 && a->act == VCT_ACT_NONE && use_persistent) {
        if (!a->a_trans && !a->b_trans) return launch_persistent<false, false>(a, tmA, tmB, st);
        if (!a->a_trans && a->b_trans) return launch_persistent<false, true>(a, tmA, tmB, st);
        if (a->a_trans && !a->b_trans) return launch_persistent<true, false>(a, tmA, tmB, st);
        return launch_persistent<true, true>(a, tmA, tmB, st);
    }
    if (bn == 256) return dispatch_major<256>(a, tmA, tmB, splits, low, st);
    if (bn == 128) return dispatch_major<128>(a, tmA, tmB, splits, low, st);
    return dispatch_major<64>(a, tmA, tmB, splits, low, st);
}

}  // namespace vct
