// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (gemm_tc.cu, attn_fused.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace vct {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                 // 64 bf16 = 128 bytes = one swizzle span
constexpr int UMMA_K = 16;
constexpr uint32_t kABytes = BLOCK_M * BLOCK_K * 2;

// cached cuTensorMapEncodeTiled: 2-D bf16 tensor, `inner` contiguous elements, `outer` rows `ld` elements apart,
// box [box_inner x box_outer], SWIZZLE_128B, zero fill out of bounds
int get_tensor_map(const void* ptr, long long inner, long long outer, long long ld, int box_inner, int box_outer, CUtensorMap* out);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000ll) __trap();   // ~2 s: a protocol bug must not hang the GPU
    }
}
// One lane of the (converged) warp: the role loops run warp-uniformly and only the TMA / MMA / commit instructions sit
// under this predicate, so ptxas keeps descriptors and addresses in uniform registers (a loop under `if (lane == 0)`
// is compiled as divergent code: every tcgen05.mma then costs ~10 instructions of R2UR moves and an elect/branch loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}

// hot-loop variant: one try_wait on the fast path (the barrier has usually completed already), the watchdog
// clock is only read once the first probe has failed
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done) mbar_wait(bar, parity);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}

// 16 consecutive fp32 columns of this warp's 32 TMEM lanes (one row per thread); completion: tcgen05.wait::ld
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_128B, sm_100 version bits (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}


}  // namespace vct
