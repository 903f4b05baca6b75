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
// Philox4x32-10 counter RNG for dropout.  The mask of element `idx` at dropout site `site` in
// training step `rng[1]` with seed `rng[0]` is a pure function of those four numbers, so forward
// and backward (and the debug entry vct_dropout_mask) regenerate identical masks without storing.
// One call yields the uniforms of the 4 consecutive elements idx4*4 .. idx4*4+3.
// ---------------------------------------------------------------------------------------------
struct Rng {
    uint32_t k0, k1, step_lo;
    float p;       // drop probability; p <= 0 disables
    float inv_keep;
};

__device__ __forceinline__ Rng make_rng(const unsigned long long* rng_state, float p) {
    Rng r;
    r.p = p;
    r.inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
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

__device__ __forceinline__ uint4 philox4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// keep-multipliers (0 or 1/(1-p)) for elements idx4*4 .. idx4*4+3 of `site`
__device__ __forceinline__ float4 dropout_scale4(const Rng& r, uint32_t site, unsigned long long idx4) {
    if (r.p <= 0.f) return make_float4(1.f, 1.f, 1.f, 1.f);
    uint4 u = philox4((uint32_t)idx4, (uint32_t)(idx4 >> 32), site, r.step_lo, r.k0, r.k1);
    const float s = 1.0f / 16777216.0f;
    float4 o;
    o.x = ((u.x >> 8) * s >= r.p) ? r.inv_keep : 0.f;
    o.y = ((u.y >> 8) * s >= r.p) ? r.inv_keep : 0.f;
    o.z = ((u.z >> 8) * s >= r.p) ? r.inv_keep : 0.f;
    o.w = ((u.w >> 8) * s >= r.p) ? r.inv_keep : 0.f;
    return o;
}
// single element
__device__ __forceinline__ float dropout_scale1(const Rng& r, uint32_t site, unsigned long long idx) {
    if (r.p <= 0.f) return 1.f;
    float4 v = dropout_scale4(r, site, idx >> 2);
    int c = (int)(idx & 3ull);
    return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w;
}

// exact-erf GELU and its derivative (activation "gelu" -> F.gelu(approximate='none'))
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_f(float x) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

}  // namespace vct
