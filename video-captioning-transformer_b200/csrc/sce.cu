// SCE loss (symmetric cross entropy), forward + backward in one pass over the logits.
// One CTA per row: the fp32 logit row (V = 30522 -> 122 KB) is staged into shared memory ONCE
// with bulk asynchronous copies (TMA engine, cp.async.bulk + mbarrier) and every pass of the
// closed form runs out of shared memory, so HBM sees one read of the logits and one
// write of the gradient.
#include <cstdlib>

#include "common.cuh"

using namespace vct;

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr float kRceA = 9.210340371976182f;   // -ln(1e-4)
constexpr float kPMin = 1e-7f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += red[w];
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = -INFINITY;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t = fmaxf(t, red[w]);
    return t;
}

template <typename TD>
__global__ void __launch_bounds__(kThreads, 1)
sce_kernel(const float* __restrict__ logits, long long ld_logits, const long long* __restrict__ ids, long long ids_ld,
           int B, int S, int V, float alpha, float beta, int pad_id, float* __restrict__ loss_out,
           float* __restrict__ row_parts, unsigned int* counter, TD* __restrict__ dlogits, long long ld_dl,
           const float* __restrict__ upstream, int vec) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* zs = reinterpret_cast<float*>(smem_raw);   // [ld_logits]
    __shared__ __align__(8) unsigned long long bar;
    __shared__ float red[kWarps];
    __shared__ bool is_last;

    const int row = blockIdx.x;
    const int N = B * S;
    const int tid = threadIdx.x;
    const uint32_t bar_a = smem_u32(&bar);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t total = (uint32_t)(ld_logits * sizeof(float));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(total) : "memory");
        const char* src = reinterpret_cast<const char*>(logits + (long long)row * ld_logits);
        for (uint32_t off = 0; off < total; off += 32768u) {
            const uint32_t n = total - off < 32768u ? total - off : 32768u;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32(smem_raw + off)),
                "l"(src + off), "r"(n), "r"(bar_a)
                : "memory");
        }
    }
    // overlap with the copy: count non-pad labels (every CTA recounts the N labels; they sit in L2)
    float cnt = 0.f;
    for (int i = tid; i < N; i += kThreads) cnt += ids[(long long)(i / S) * ids_ld + (i % S) + 1] != pad_id ? 1.f : 0.f;
    const float n_valid = block_sum(cnt, red);
    const long long label = ids[(long long)(row / S) * ids_ld + (row % S) + 1];
    const int y = (int)(label < 0 ? 0 : (label >= V ? V - 1 : label));
    const bool valid = label != pad_id;

    {   // wait for the row
        uint32_t done = 0;
        for (long long spin = 0; !done; ++spin) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar_a)
                : "memory");
            if (spin > (1ll << 26)) __trap();
        }
    }

    float ce_i = 0.f, rce_i = 0.f;
    if (vec) {
        // ---- vectorised path (row strides multiples of 8): 16-byte shared-memory accesses, exp through one MUFU
        //      (exp2 of the log2(e)-scaled difference, relative error 2^-22), 16-byte gradient stores.  The kernel is
        //      bound by instruction issue (one CTA per SM: the 122 KB row fills shared memory), so every pass is
        //      written for the fewest instructions per element. ----
        const int n4 = (int)(ld_logits >> 2);
        float4* z4 = reinterpret_cast<float4*>(zs);
        float m = -INFINITY;
        for (int i = tid; i < n4; i += kThreads) {
            float4 v = z4[i];
            const int c = 4 * i;
            if (c + 3 >= V) {                                   // padding columns of the logits buffer are undefined
                if (c + 0 >= V) v.x = -INFINITY;
                if (c + 1 >= V) v.y = -INFINITY;
                if (c + 2 >= V) v.z = -INFINITY;
                v.w = -INFINITY;
                z4[i] = v;
            }
            m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
        }
        m = block_max(m, red);
        const float zy = zs[y];
        __syncthreads();
        const float ml2 = m * 1.4426950408889634f;
        float se = 0.f;
        for (int i = tid; i < n4; i += kThreads) {
            float4 v = z4[i];
            v.x = exp2f(fmaf(v.x, 1.4426950408889634f, -ml2));
            v.y = exp2f(fmaf(v.y, 1.4426950408889634f, -ml2));
            v.z = exp2f(fmaf(v.z, 1.4426950408889634f, -ml2));
            v.w = exp2f(fmaf(v.w, 1.4426950408889634f, -ml2));
            z4[i] = v;
            se += (v.x + v.y) + (v.z + v.w);
        }
        se = block_sum(se, red);
        const float inv = 1.f / se;
        ce_i = valid ? (m + logf(se) - zy) : 0.f;
        float U = 0.f;
        if (alpha != 1.0f) {
            // U = sum_{c != y, p_c >= 1e-7} p_c; the clamped classes are counted instead
            const float ethr = kPMin * se;                       // p >= pmin  <=>  e >= pmin * se
            float u = 0.f, nclamp = 0.f;
            for (int i = tid; i < n4; i += kThreads) {
                const float4 v = z4[i];
                u += (v.x >= ethr ? v.x : 0.f) + (v.y >= ethr ? v.y : 0.f) + (v.z >= ethr ? v.z : 0.f) + (v.w >= ethr ? v.w : 0.f);
                nclamp += (v.x < ethr ? 1.f : 0.f) + (v.y < ethr ? 1.f : 0.f) + (v.z < ethr ? 1.f : 0.f) + (v.w < ethr ? 1.f : 0.f);
            }
            u = block_sum(u, red);
            nclamp = block_sum(nclamp, red);
            // remove the label class and the (zero-probability) padding columns from both tallies
            const float ey = zs[y];
            if (ey >= ethr) u -= ey; else nclamp -= 1.f;
            nclamp -= (float)(4 * n4 - V);
            U = u * inv;
            rce_i = kRceA * (U + nclamp * kPMin);
        }
        if (dlogits != nullptr) {
            const float up = upstream ? upstream[0] : 1.f;
            const float a = (alpha == 1.0f ? 1.f : alpha) * (valid ? up / fmaxf(n_valid, 1.f) : 0.f);
            const float bb = alpha == 1.0f ? 0.f : beta * kRceA * up / (float)N;
            // g = a (p - [c == y]) + bb ([c != y, p >= pmin] p - p U) = e * (ka or kb) with the label fixed up afterwards
            const float ethr = kPMin * se;
            const float k_keep = (a + bb * (1.f - U)) * inv;      // classes at or above the clamp
            const float k_clamp = (a - bb * U) * inv;             // clamped classes (their RCE term has no gradient)
            TD* drow = dlogits + (long long)row * ld_dl;
            const int n8 = (int)(ld_dl >> 3);
            for (int i = tid; i < n8; i += kThreads) {
                float g[8];
                if (8 * i < 4 * n4) {
                    const float4 v0 = z4[2 * i], v1 = (2 * i + 1 < n4) ? z4[2 * i + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                    for (int q = 0; q < 8; ++q) g[q] = e[q] * (e[q] >= ethr ? k_keep : k_clamp);
                    const int rel = y - 8 * i;
                    if (rel >= 0 && rel < 8) {
                        const float py = zs[y] * inv;
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (q == rel) g[q] = a * (py - 1.f) - bb * py * U;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) g[q] = 0.f;
                }
                st8(drow + 8 * i, g);
            }
        }
    } else {
    float m = -INFINITY;
    for (int c = tid; c < V; c += kThreads) m = fmaxf(m, zs[c]);
    m = block_max(m, red);
    const float zy = zs[y];
    __syncthreads();
    float se = 0.f;
    for (int c = tid; c < V; c += kThreads) {
        const float e = expf(zs[c] - m);
        zs[c] = e;
        se += e;
    }
    se = block_sum(se, red);
    const float inv = 1.f / se;
    ce_i = valid ? (m + logf(se) - zy) : 0.f;

    float U = 0.f;
    if (alpha != 1.0f) {
        float u = 0.f, nclamp = 0.f;
        for (int c = tid; c < V; c += kThreads) {
            if (c == y) continue;
            const float p = zs[c] * inv;
            if (p >= kPMin) u += p; else nclamp += 1.f;
        }
        U = block_sum(u, red);
        const float ncl = block_sum(nclamp, red);
        rce_i = kRceA * (U + ncl * kPMin);
    }

    if (dlogits != nullptr) {
        const float up = upstream ? upstream[0] : 1.f;
        const float a = (alpha == 1.0f ? 1.f : alpha) * (valid ? up / fmaxf(n_valid, 1.f) : 0.f);
        const float bb = alpha == 1.0f ? 0.f : beta * kRceA * up / (float)N;
        TD* drow = dlogits + (long long)row * ld_dl;
        for (int c = tid; c < (int)ld_dl; c += kThreads) {
            float g = 0.f;
            if (c < V) {
                const float p = zs[c] * inv;
                g = a * (p - (c == y ? 1.f : 0.f));
                g += bb * (((c != y && p >= kPMin) ? p : 0.f) - p * U);
            }
            drow[c] = from_f32<TD>(g);
        }
    }
    }   // scalar path

    if (loss_out != nullptr) {
        if (tid == 0) {
            row_parts[2 * row + 0] = ce_i;
            row_parts[2 * row + 1] = rce_i;
            __threadfence();
            const unsigned int prev = atomicAdd(counter, 1u);
            is_last = (prev == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            float ce = 0.f, rce = 0.f;
            for (int i = tid; i < N; i += kThreads) {
                ce += __ldcg(row_parts + 2 * i);
                rce += __ldcg(row_parts + 2 * i + 1);
            }
            ce = block_sum(ce, red);
            rce = block_sum(rce, red);
            if (tid == 0) {
                const float ce_mean = ce / fmaxf(n_valid, 1.f);
                loss_out[0] = alpha == 1.0f ? ce_mean : alpha * ce_mean + beta * rce / (float)N;
                *counter = 0u;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Register-resident variant (the product path): 1024 threads, every thread keeps its 32 logits of the row in REGISTERS
// (4 groups of 8 consecutive columns, 16-byte loads), so nothing is staged in shared memory -- the kernel above is bound
// by shared-memory instruction issue with its 122 KB row and one 16-warp CTA per SM (0.38 of the HBM roofline).  Logits
// may be stored in bf16 (TL): the training path then never materialises fp32 logits (156 MB written + read per step at
// B = 64), model/CapDecoder.py:55-59 -> model/loss.py:78-92 discards them anyway (model/MMT4Caption.py:120-121).
// Same closed form and the same reduction structure as above.  Requires ld % 8 == 0, ld <= 32768, ld_dl == ld.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kRThreads = 1024;
constexpr int kRWarps = kRThreads / 32;
constexpr int kRGroups = 4;

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(p)), b = __ldcs(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 r = __ldcs(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

// Block reductions of the register kernel: warp shuffle, one shared-memory slot per warp, and a SECOND shuffle stage in
// which every warp reduces the 32 slots itself (one load + 5 shuffles per thread; a loop over the 32 slots costs ~100
// instructions per thread and made this kernel issue-bound: 40 % of its instructions were reductions).
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* red) {
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = make_float2(a, b);
    __syncthreads();
    const float2 t = red[threadIdx.x & 31];
    return make_float2(warp_sum(t.x), warp_sum(t.y));
}
__device__ __forceinline__ float block_max1(float m, float2* red) {
    m = warp_max(m);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5].x = m;
    __syncthreads();
    return warp_max(red[threadIdx.x & 31].x);
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void lds8(const float* p, float (&v)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void lds8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

// Persistent: one CTA per SM walks the rows blockIdx.x, blockIdx.x + gridDim.x, ...  The NEXT row is fetched into shared
// memory by the TMA engine (cp.async.bulk + mbarrier) while the current row -- already copied into registers -- goes
// through the reductions and its gradient is stored, so HBM sees a continuous stream instead of load / compute / store
// phases of one resident CTA (a non-persistent version of this kernel reached 1.6 TB/s).
template <typename TL, typename TD>
__global__ void __launch_bounds__(kRThreads, 1)
sce_reg_kernel(const TL* __restrict__ logits, long long ld, const long long* __restrict__ ids, long long ids_ld,
               int B, int S, int V, float alpha, float beta, int pad_id, float* __restrict__ loss_out,
               float* __restrict__ row_parts, unsigned int* counter, TD* __restrict__ dlogits,
               const float* __restrict__ upstream) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TL* zs = reinterpret_cast<const TL*>(smem_raw);        // [ld] the row in flight
    __shared__ __align__(8) unsigned long long bar;
    __shared__ float2 red[kRWarps];
    __shared__ float s_zy, s_ey;
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const int N = B * S;
    const int ngroups = (int)(ld >> 3);
    const uint32_t bar_a = smem_u32(&bar);
    const uint32_t row_bytes = (uint32_t)(ld * sizeof(TL));
    auto fetch = [&](int r) {                                     // thread 0 only
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(row_bytes) : "memory");
        const char* src = reinterpret_cast<const char*>(logits + (long long)r * ld);
        for (uint32_t off = 0; off < row_bytes; off += 32768u) {
            const uint32_t n = row_bytes - off < 32768u ? row_bytes - off : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(smem_raw + off)), "l"(src + off), "r"(n), "r"(bar_a) : "memory");
        }
    };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int row = blockIdx.x;
    if (tid == 0 && row < N) fetch(row);
    // while the first row is in flight: count the non-pad labels (every CTA recounts the N labels; they sit in L2)
    float n_valid;
    {
        float cnt = 0.f;
        for (int i = tid; i < N; i += kRThreads) cnt += ids[(long long)(i / S) * ids_ld + (i % S) + 1] != pad_id ? 1.f : 0.f;
        n_valid = block_sum2(cnt, 0.f, red).x;
    }
    const float up = upstream ? upstream[0] : 1.f;
    uint32_t phase = 0;
    for (; row < N; row += gridDim.x) {
        {   // wait for the row
            uint32_t done = 0;
            for (long long spin = 0; !done; ++spin) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(bar_a), "r"(phase)
                    : "memory");
                if (spin > (1ll << 26)) __trap();
            }
            phase ^= 1u;
        }
        float z[kRGroups][8];
#pragma unroll
        for (int g = 0; g < kRGroups; ++g) {
            const int grp = tid + g * kRThreads;
            if (grp < ngroups) lds8(zs + 8 * grp, z[g]);
            else {
#pragma unroll
                for (int q = 0; q < 8; ++q) z[g][q] = -INFINITY;
            }
        }
        __syncthreads();                                          // the buffer has been read by everyone ...
        if (tid == 0 && row + (int)gridDim.x < N) fetch(row + gridDim.x);   // ... and takes the next row
        const long long label = ids[(long long)(row / S) * ids_ld + (row % S) + 1];
        const int y = (int)(label < 0 ? 0 : (label >= V ? V - 1 : label));
        const bool valid = label != pad_id;
        const int ygrp = y >> 3, yq = y & 7;
        const bool mine_y = (ygrp % kRThreads) == tid;
        const int yg = ygrp / kRThreads;
        float m = -INFINITY;
#pragma unroll
        for (int g = 0; g < kRGroups; ++g) {
            const int c0 = 8 * (tid + g * kRThreads);
            if (c0 + 8 > V) {                                    // only the group(s) that straddle V: the padding columns of
#pragma unroll                                                  // the logits buffer are undefined
                for (int q = 0; q < 8; ++q)
                    if (c0 + q >= V) z[g][q] = -INFINITY;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) m = fmaxf(m, z[g][q]);
        }
        if (mine_y) {
#pragma unroll
            for (int g = 0; g < kRGroups; ++g)
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (g == yg && q == yq) s_zy = z[g][q];
        }
        m = block_max1(m, red);
        const float zy = s_zy;
        const float ml2 = m * 1.4426950408889634f;
        float se = 0.f;
#pragma unroll
        for (int g = 0; g < kRGroups; ++g)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                z[g][q] = ex2_approx(fmaf(z[g][q], 1.4426950408889634f, -ml2));   // one MUFU; exp2(-inf) = 0 for the padding
                se += z[g][q];
            }
        if (mine_y) {
#pragma unroll
            for (int g = 0; g < kRGroups; ++g)
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (g == yg && q == yq) s_ey = z[g][q];
        }
        se = block_sum2(se, 0.f, red).x;
        const float inv = 1.f / se;
        const float ce_i = valid ? (m + logf(se) - zy) : 0.f;
        const float ey = s_ey;
        const float ethr = kPMin * se;                           // p >= pmin  <=>  e >= pmin * se
        float U = 0.f, rce_i = 0.f;
        if (alpha != 1.0f) {
            float u = 0.f, nclamp = 0.f;
#pragma unroll
            for (int g = 0; g < kRGroups; ++g)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    u += z[g][q] >= ethr ? z[g][q] : 0.f;
                    nclamp += z[g][q] < ethr ? 1.f : 0.f;
                }
            const float2 t = block_sum2(u, nclamp, red);
            u = t.x;
            nclamp = t.y;
            // remove the label class and the (zero-probability) padding columns / unused register slots from both tallies
            if (ey >= ethr) u -= ey; else nclamp -= 1.f;
            nclamp -= (float)(kRGroups * kRThreads * 8 - V);
            U = u * inv;
            rce_i = kRceA * (U + nclamp * kPMin);
        }
        if (dlogits != nullptr) {
            const float a = (alpha == 1.0f ? 1.f : alpha) * (valid ? up / fmaxf(n_valid, 1.f) : 0.f);
            const float bb = alpha == 1.0f ? 0.f : beta * kRceA * up / (float)N;
            const float k_keep = (a + bb * (1.f - U)) * inv;      // classes at or above the clamp
            const float k_clamp = (a - bb * U) * inv;             // clamped classes (their RCE term has no gradient)
            const float py = ey * inv;
            const float gy = a * (py - 1.f) - bb * py * U;
            TD* drow = dlogits + (long long)row * ld;
#pragma unroll
            for (int g = 0; g < kRGroups; ++g) {
                const int grp = tid + g * kRThreads;
                if (grp < ngroups) {
                    float gr[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) gr[q] = z[g][q] * (z[g][q] >= ethr ? k_keep : k_clamp);
                    if (mine_y && g == yg) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (q == yq) gr[q] = gy;
                    }
                    st8(drow + 8 * grp, gr);
                }
            }
        }
        if (loss_out != nullptr && tid == 0) {
            row_parts[2 * row + 0] = ce_i;
            row_parts[2 * row + 1] = rce_i;
        }
    }
    if (loss_out != nullptr) {
        if (tid == 0) {
            __threadfence();
            const unsigned int prev = atomicAdd(counter, 1u);
            is_last = (prev == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            float ce = 0.f, rce = 0.f;
            for (int i = tid; i < N; i += kRThreads) {
                ce += __ldcg(row_parts + 2 * i);
                rce += __ldcg(row_parts + 2 * i + 1);
            }
            const float2 t = block_sum2(ce, rce, red);
            if (tid == 0) {
                const float ce_mean = t.x / fmaxf(n_valid, 1.f);
                loss_out[0] = alpha == 1.0f ? ce_mean : alpha * ce_mean + beta * t.y / (float)N;
                *counter = 0u;
            }
        }
    }
}

template <typename TL, typename TD>
int launch_sce_reg(const void* logits, long long ld, const long long* ids, long long ids_ld, int B, int S, int V, float alpha,
                   float beta, int pad_id, float* loss_out, float* row_parts, unsigned int* counter, void* dlogits,
                   const float* upstream, cudaStream_t st) {
    auto kern = sce_reg_kernel<TL, TD>;
    const size_t smem = (size_t)ld * sizeof(TL);
    static bool once = false;
    if (!once) {
        VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024));
        once = true;
    }
    const int N = B * S;
    vct::launch(kern, dim3(N < kNumSMs ? N : kNumSMs), dim3(kRThreads), smem, st, (const TL*)logits, ld, ids, ids_ld, B, S, V, alpha, beta,
                pad_id, loss_out, row_parts, counter, (TD*)dlogits, upstream);
    return check_launch("vct_sce");
}

}  // namespace

extern "C" int vct_sce_typed(const void* logits_v, int logits_dtype, long long ld_logits, const long long* ids, long long ids_ld,
                             int B, int S, int V, float alpha, float beta, int pad_id, float* loss_out, float* row_parts,
                             unsigned int* counter, void* dlogits, int dl_dtype, long long ld_dl, const float* upstream,
                             vct_stream_t stream) {
    VCT_REQUIRE(logits_v && ids && B > 0 && S > 0 && V > 0, "vct_sce: bad arguments");
    VCT_REQUIRE(logits_dtype == VCT_F32 || logits_dtype == VCT_BF16, "vct_sce: bad logits dtype");
    VCT_REQUIRE(loss_out == nullptr || (row_parts && counter), "vct_sce: loss needs row_parts and counter");
    VCT_REQUIRE((reinterpret_cast<uintptr_t>(logits_v) & 15) == 0, "vct_sce: logits must be 16-byte aligned");
    {
        // register-resident kernel: rows of up to 32768 columns with 16-byte aligned strides, gradient laid out like the logits
        static const bool reg_on = [] { const char* e = getenv("VCT_SCE_REG"); return e == nullptr || e[0] != '0'; }();
        const bool ok = ld_logits % 8 == 0 && ld_logits >= V && ld_logits <= (long long)kRGroups * kRThreads * 8 &&
                        (dlogits == nullptr || (ld_dl == ld_logits && (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0));
        if (ok && (reg_on || logits_dtype == VCT_BF16)) {
            cudaStream_t st = (cudaStream_t)stream;
#define VCT_SCE_GO(TL, TD) return launch_sce_reg<TL, TD>(logits_v, ld_logits, ids, ids_ld, B, S, V, alpha, beta, pad_id, loss_out, row_parts, counter, dlogits, upstream, st)
            if (logits_dtype == VCT_BF16) { if (dl_dtype == VCT_BF16) VCT_SCE_GO(__nv_bfloat16, __nv_bfloat16); else VCT_SCE_GO(__nv_bfloat16, float); }
            else { if (dl_dtype == VCT_BF16) VCT_SCE_GO(float, __nv_bfloat16); else VCT_SCE_GO(float, float); }
#undef VCT_SCE_GO
        }
    }
    VCT_REQUIRE(logits_dtype == VCT_F32, "vct_sce: bf16 logits need ld %% 8 == 0, ld <= 32768 and ld_dl == ld_logits");
    const float* logits = reinterpret_cast<const float*>(logits_v);
    VCT_REQUIRE(ld_logits % 4 == 0 && ld_logits >= V, "vct_sce: ld_logits must be a multiple of 4 and >= V");
    VCT_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "vct_sce: logits must be 16-byte aligned");
    const size_t smem = (size_t)ld_logits * sizeof(float);
    VCT_REQUIRE(smem <= 220 * 1024, "vct_sce: a logit row of %lld floats does not fit shared memory", ld_logits);
    VCT_REQUIRE(loss_out == nullptr || (row_parts && counter), "vct_sce: loss needs row_parts and counter");
    VCT_REQUIRE(dlogits == nullptr || ld_dl >= V, "vct_sce: ld_dl < V");
    cudaStream_t st = (cudaStream_t)stream;
    const int vec = (ld_logits % 8 == 0) && (dlogits == nullptr || (ld_dl % 8 == 0 && ld_dl <= ld_logits + 8 &&
                                                                    (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0)) ? 1 : 0;
    static bool attr_done[2] = {false, false};
    if (dl_dtype == VCT_BF16) {
        auto kern = sce_kernel<__nv_bfloat16>;
        if (!attr_done[1]) {
            VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            attr_done[1] = true;
        }
        vct::launch(kern, dim3(B * S), dim3(kThreads), smem, st, logits, ld_logits, ids, ids_ld, B, S, V, alpha, beta, pad_id, loss_out,
                                            row_parts, counter, (__nv_bfloat16*)dlogits, ld_dl, upstream, vec);
    } else {
        auto kern = sce_kernel<float>;
        if (!attr_done[0]) {
            VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            attr_done[0] = true;
        }
        vct::launch(kern, dim3(B * S), dim3(kThreads), smem, st, logits, ld_logits, ids, ids_ld, B, S, V, alpha, beta, pad_id, loss_out,
                                            row_parts, counter, (float*)dlogits, ld_dl, upstream, vec);
    }
    return check_launch("vct_sce");
}

extern "C" int vct_sce(const float* logits, long long ld_logits, const long long* ids, long long ids_ld, int B, int S,
                       int V, float alpha, float beta, int pad_id, float* loss_out, float* row_parts,
                       unsigned int* counter, void* dlogits, int dl_dtype, long long ld_dl, const float* upstream,
                       vct_stream_t stream) {
    return vct_sce_typed(logits, VCT_F32, ld_logits, ids, ids_ld, B, S, V, alpha, beta, pad_id, loss_out, row_parts, counter, dlogits,
                         dl_dtype, ld_dl, upstream, stream);
}
