// Shared device/host helpers for the vct_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vct.h"

namespace vct {

// ---------------------------------------------------------------------------------------------
// error handling: every entry point returns 0 or a negative code; message via vct_last_error()
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define VCT_REQUIRE(cond, ...)                         \
    do {                                               \
        if (!(cond)) {                                 \
            ::vct::set_error(__VA_ARGS__);             \
            return VCT_ERR_INVALID;                    \
        }                                              \
    } while (0)

#define VCT_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            ::vct::set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
            return VCT_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

constexpr int kNumSMs = 148;   // B200

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: every kernel is launched with the stream-serialization attribute, lets its
// dependents start early (griddepcontrol.launch_dependents) and waits for its prerequisites
// (griddepcontrol.wait) before touching global memory.  A step is a chain of ~160 short dependent kernels, so
// overlapping each kernel's launch + prologue with its predecessor's tail is worth ~2 us per launch.
// Both instructions are no-ops when the kernel was launched without the attribute.  VCT_PDL=0 disables it.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// dtype helpers (VCT_F32 / VCT_BF16 storage, fp32 arithmetic)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// load / store 4 consecutive elements as fp32 (pointer must be 4-element aligned)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
}

// ---------------------------------------------------------------------------------------------
// warp reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-7 counter RNG for dropout.  The mask of element `idx` at dropout site `site` in training
// step `rng[1]` with seed `rng[0]` is a pure function of those four numbers, so forward and backward
// (and the debug entry vct_dropout_mask) regenerate identical masks without storing them.
// One call yields 128 bits = 16-bit uniforms for the 8 consecutive elements idx8*8 .. idx8*8+7
// (keep iff r16 >= round(p * 65536)): the RNG is ~1/3 of the instruction count of a train step if
// done per 4 elements with 10 rounds, so bits are not wasted.  7 rounds is the minimum the Philox
// authors qualify (Salmon et al., SC11, table 2: Philox4x32-7 passes BigCrush).
// ---------------------------------------------------------------------------------------------
struct Rng {
    uint32_t k0, k1, step_lo;
    uint32_t thresh;   // drop iff r16 < thresh
    float p;           // drop probability; p <= 0 disables
    float inv_keep;
};

__device__ __forceinline__ Rng make_rng(const unsigned long long* rng_state, float p) {
    Rng r;
    r.p = p;
    r.inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
    r.thresh = p > 0.f ? (uint32_t)(p * 65536.f + 0.5f) : 0u;
    if (p > 0.f && rng_state != nullptr) {
        unsigned long long seed = rng_state[0], step = rng_state[1];
        r.k0 = (uint32_t)seed;
        r.k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32);
        r.step_lo = (uint32_t)step;
    } else {
        r.k0 = r.k1 = r.step_lo = 0;
    }
    return r;
}

__device__ __forceinline__ uint4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// keep-multipliers (0 or 1/(1-p)) of elements idx8*8 .. idx8*8+7 of `site`
__device__ __forceinline__ void dropout_scale8(const Rng& r, uint32_t site, unsigned long long idx8, float* out) {
    if (r.p <= 0.f) {
#pragma unroll
        for (int q = 0; q < 8; ++q) out[q] = 1.f;
        return;
    }
    const uint4 u = philox4x32_7((uint32_t)idx8, (uint32_t)(idx8 >> 32), site, r.step_lo, r.k0, r.k1);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        out[2 * q + 0] = (w[q] & 0xFFFFu) >= r.thresh ? r.inv_keep : 0.f;
        out[2 * q + 1] = (w[q] >> 16) >= r.thresh ? r.inv_keep : 0.f;
    }
}
// the same 8 decisions as a bit mask (bit q set = element idx8*8+q is kept)
__device__ __forceinline__ uint32_t dropout_bits8(const Rng& r, uint32_t site, unsigned long long idx8) {
    if (r.p <= 0.f) return 0xFFu;
    const uint4 u = philox4x32_7((uint32_t)idx8, (uint32_t)(idx8 >> 32), site, r.step_lo, r.k0, r.k1);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t bits = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        bits |= ((w[q] & 0xFFFFu) >= r.thresh ? 1u : 0u) << (2 * q);
        bits |= ((w[q] >> 16) >= r.thresh ? 1u : 0u) << (2 * q + 1);
    }
    return bits;
}
// load / store 8 consecutive elements as fp32 (pointer must be 8-element aligned for bf16, 4 for fp32)
__device__ __forceinline__ void ld8(const float* p, float* o) {
    const float4 a = ld4(p), b = ld4(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* o) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[q]));
        o[2 * q] = f.x; o[2 * q + 1] = f.y;
    }
}
__device__ __forceinline__ void st8(float* p, const float* v) {
    st4(p, make_float4(v[0], v[1], v[2], v[3]));
    st4(p + 4, make_float4(v[4], v[5], v[6], v[7]));
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
        w[q] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
// single element (scalar fallback paths)
__device__ __forceinline__ float dropout_scale1(const Rng& r, uint32_t site, unsigned long long idx) {
    if (r.p <= 0.f) return 1.f;
    float v[8];
    dropout_scale8(r, site, idx >> 3, v);
    float o = v[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) o = (int)(idx & 7ull) == q ? v[q] : o;
    return o;
}

// exact-erf GELU and its derivative (activation "gelu" -> F.gelu(approximate='none'))
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_f(float x) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// Fast variants for the bf16 tensor-core epilogues (outputs are rounded to bf16, 2^-9 relative): the normal CDF from
// Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7) in its complementary form, so that the left tail has no cancellation;
// the exponential exp(-x^2/2) is shared between the CDF and the density.  ~15 instructions against ~45 for erff,
// which matters because only the four epilogue warps of a CTA run this code.
__device__ __forceinline__ void normal_cdf_pdf_fast(float x, float& cdf, float& pdf) {
    const float ax = fabsf(x) * 0.70710678118654752f;
    float t;                                                 // 1 / (1 + p |x|): one MUFU (the IEEE-rounded __frcp_rn is ~10 instructions)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = __expf(-ax * ax);                       // exp(-x^2 / 2)
    const float half_erfc = 0.5f * p * t * e;               // 0.5 * erfc(|x| / sqrt 2)
    cdf = x < 0.f ? half_erfc : 1.f - half_erfc;
    pdf = 0.39894228040143268f * e;
}
__device__ __forceinline__ float gelu_fast(float x) {
    float cdf, pdf;
    normal_cdf_pdf_fast(x, cdf, pdf);
    return x * cdf;
}
__device__ __forceinline__ float dgelu_fast(float x) {
    float cdf, pdf;
    normal_cdf_pdf_fast(x, cdf, pdf);
    return fmaf(x, pdf, cdf);
}

}  // namespace vct
