// SCE loss (symmetric cross entropy), forward + backward in one pass over the logits.
// One CTA per row: the fp32 logit row (V = 30522 -> 122 KB) is staged into shared memory ONCE
// with bulk asynchronous copies (TMA engine, cp.async.bulk + mbarrier) and every pass of the
// closed form runs out of shared memory, so HBM sees one read of the logits and one
// write of the gradient.
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
           const float* __restrict__ upstream) {
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
    const float ce_i = valid ? (m + logf(se) - zy) : 0.f;

    float U = 0.f, rce_i = 0.f;
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

}  // namespace

extern "C" int vct_sce(const float* logits, long long ld_logits, const long long* ids, long long ids_ld, int B, int S,
                       int V, float alpha, float beta, int pad_id, float* loss_out, float* row_parts,
                       unsigned int* counter, void* dlogits, int dl_dtype, long long ld_dl, const float* upstream,
                       vct_stream_t stream) {
    VCT_REQUIRE(logits && ids && B > 0 && S > 0 && V > 0, "vct_sce: bad arguments");
    VCT_REQUIRE(ld_logits % 4 == 0 && ld_logits >= V, "vct_sce: ld_logits must be a multiple of 4 and >= V");
    VCT_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "vct_sce: logits must be 16-byte aligned");
    const size_t smem = (size_t)ld_logits * sizeof(float);
    VCT_REQUIRE(smem <= 220 * 1024, "vct_sce: a logit row of %lld floats does not fit shared memory", ld_logits);
    VCT_REQUIRE(loss_out == nullptr || (row_parts && counter), "vct_sce: loss needs row_parts and counter");
    VCT_REQUIRE(dlogits == nullptr || ld_dl >= V, "vct_sce: ld_dl < V");
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr_done[2] = {false, false};
    if (dl_dtype == VCT_BF16) {
        auto kern = sce_kernel<__nv_bfloat16>;
        if (!attr_done[1]) {
            VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            attr_done[1] = true;
        }
        vct::launch(kern, dim3(B * S), dim3(kThreads), smem, st, logits, ld_logits, ids, ids_ld, B, S, V, alpha, beta, pad_id, loss_out,
                                            row_parts, counter, (__nv_bfloat16*)dlogits, ld_dl, upstream);
    } else {
        auto kern = sce_kernel<float>;
        if (!attr_done[0]) {
            VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            attr_done[0] = true;
        }
        vct::launch(kern, dim3(B * S), dim3(kThreads), smem, st, logits, ld_logits, ids, ids_ld, B, S, V, alpha, beta, pad_id, loss_out,
                                            row_parts, counter, (float*)dlogits, ld_dl, upstream);
    }
    return check_launch("vct_sce");
}
