// HBM-bound kernels of the hot path: frame staging, residual+dropout+LayerNorm (fwd/bwd),
// embedding (fwd/bwd), column sums, Adam, casts, greedy argmax, step tick, error plumbing.
#include <stdarg.h>
#include <stdlib.h>

#include <cstdlib>

#include "common.cuh"

namespace vct {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("VCT_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return VCT_ERR_CUDA;
    }
    return 0;
}

}  // namespace vct

using namespace vct;

extern "C" int vct_version(void) { return 100; }
extern "C" const char* vct_last_error(void) { return vct::g_err; }

extern "C" int vct_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    VCT_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    VCT_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    VCT_REQUIRE(prop.major == 10, "libvct_b200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// step tick
// ------------------------------------------------------------------------------------------------
__global__ void step_tick_kernel(unsigned long long* rng_state, float* hyper) {
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (rng_state) rng_state[1] += 1ull;
        if (hyper) {
            float step = hyper[5] + 1.f;
            hyper[5] = step;
            hyper[6] = 1.f - powf(hyper[1], step);
            hyper[7] = 1.f - powf(hyper[2], step);
        }
    }
}

extern "C" int vct_step_tick(unsigned long long* rng_state, float* adam_hyper, vct_stream_t stream) {
    vct::launch(step_tick_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, rng_state, adam_hyper);
    return check_launch("vct_step_tick");
}

// ------------------------------------------------------------------------------------------------
// frame staging: [B,T,Din] fp32 -> [B*(T+1), Din] with the mean row first
// ------------------------------------------------------------------------------------------------
template <typename TO>
__global__ void prep_frames_kernel(const float* __restrict__ feats, TO* __restrict__ out, int B, int T, int Din) {
    pdl_launch_dependents();
    pdl_wait();
    const int nv = Din >> 2;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * nv) return;
    const int b = (int)(gid / nv), c = (int)(gid % nv) * 4;
    const float* src = feats + (long long)b * T * Din + c;
    TO* dst = out + (long long)b * (T + 1) * Din + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
        float4 v = ld4(src + (long long)t * Din);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        st4(dst + (long long)(t + 1) * Din, v);
    }
    const float inv = 1.f / (float)T;
    st4(dst, make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv));
}

extern "C" int vct_prep_frames(const float* feats, void* out, int out_dtype, int B, int T, int Din, vct_stream_t stream) {
    VCT_REQUIRE(Din % 4 == 0 && B > 0 && T > 0, "vct_prep_frames: Din %% 4 != 0 or empty input");
    long long n = (long long)B * (Din / 4);
    int blocks = (int)((n + 255) / 256);
    if (out_dtype == VCT_BF16)
        vct::launch(prep_frames_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, feats, (__nv_bfloat16*)out, B, T, Din);
    else
        vct::launch(prep_frames_kernel<float>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, feats, (float*)out, B, T, Din);
    return check_launch("vct_prep_frames");
}

// ------------------------------------------------------------------------------------------------
// residual + dropout + LayerNorm forward: one warp per row, NV chunks of 8 elements per lane
// ------------------------------------------------------------------------------------------------
constexpr int kLnWarps = 8;

template <int NV, typename TC>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* r, const float* __restrict__ gamma,
              const float* __restrict__ beta, float* __restrict__ y, TC* __restrict__ y_c, float* s_out,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, int R, int d, float drop_p,
              const unsigned long long* __restrict__ rng_state, unsigned int site) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kLnWarps + warp;
    if (row >= R) return;
    const int nv = d >> 3;
    const Rng rng = make_rng(rng_state, x != nullptr ? drop_p : 0.f);
    const long long base = (long long)row * d;
    // every global load of the row (branch output, residual, gamma, beta) is issued before anything consumes one: ONE
    // memory round trip per row.  (Interleaved with the Philox rounds the compiler emitted three batches of loads, each
    // waiting for the previous batch's arithmetic, and fetched gamma / beta only after the statistics.)
    float v[NV][8], xv[NV][8], g[NV][8], bt[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nv) {
            ld8(r + base + c * 8, v[i]);
            if (x != nullptr) ld8(x + base + c * 8, xv[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nv) {
            ld8(gamma + c * 8, g[i]);
            ld8(beta + c * 8, bt[i]);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nv) {
            if (x != nullptr) {
                float sc[8];
                dropout_scale8(rng, site, (unsigned long long)(base >> 3) + c, sc);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[i][q] = xv[i][q] + v[i][q] * sc[q];
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) sum += v[i][q];
        }
    }
    const float mean = warp_sum(sum) / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nv) {
#pragma unroll
            for (int q = 0; q < 8; ++q) { const float a = v[i][q] - mean; sq += a * a; }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)d + 1e-5f);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nv) {
            if (s_out) st8(s_out + base + c * 8, v[i]);
            float o[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = (v[i][q] - mean) * rstd * g[i][q] + bt[i][q];
            if (y) st8(y + base + c * 8, o);
            if (y_c) st8(y_c + base + c * 8, o);
        }
    }
}

template <typename TC>
static int launch_ln_fwd(const float* x, const float* r, const float* gamma, const float* beta, float* y, TC* y_c,
                         float* s_out, float* mean, float* rstd, int R, int d, float drop_p,
                         const unsigned long long* rng_state, unsigned int site, cudaStream_t st) {
    const int blocks = (R + kLnWarps - 1) / kLnWarps;
    const int nvl = (d / 8 + 31) / 32;
#define LN_FWD_CASE(NVV)                                                                                         \
    vct::launch(ln_fwd_kernel<NVV, TC>, dim3(blocks), dim3(kLnWarps * 32), 0, st, x, r, gamma, beta, y, y_c, s_out, mean, rstd, R, d, \
                                                              drop_p, rng_state, site)
    if (nvl <= 1) LN_FWD_CASE(1);
    else if (nvl <= 2) LN_FWD_CASE(2);
    else if (nvl <= 3) LN_FWD_CASE(3);
    else LN_FWD_CASE(4);
#undef LN_FWD_CASE
    return check_launch("vct_ln_residual_fwd");
}

extern "C" int vct_ln_residual_fwd(const float* x, const float* r, const float* gamma, const float* beta, float* y,
                                   void* y_c, int y_c_dtype, float* s_out, float* mean, float* rstd, int R, int d,
                                   float drop_p, const unsigned long long* rng_state, unsigned int site,
                                   vct_stream_t stream) {
    VCT_REQUIRE(d % 8 == 0 && d <= 1024 && R > 0, "vct_ln_residual_fwd: need d %% 8 == 0, d <= 1024 (d=%d)", d);
    VCT_REQUIRE(r && gamma && beta, "vct_ln_residual_fwd: null input");
    if (y_c_dtype == VCT_BF16)
        return launch_ln_fwd<__nv_bfloat16>(x, r, gamma, beta, y, (__nv_bfloat16*)y_c, s_out, mean, rstd, R, d, drop_p,
                                            rng_state, site, (cudaStream_t)stream);
    return launch_ln_fwd<float>(x, r, gamma, beta, y, (float*)y_c, s_out, mean, rstd, R, d, drop_p, rng_state, site,
                                (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// residual + dropout + LayerNorm backward
//   kernel 1: one warp per row (grid ~ 1-2 CTAs per SM), per-CTA column partials [3][d] -> workspace
//   kernel 2: deterministic reduction of the partials over CTAs -> dgamma, dbeta, dbias
// ------------------------------------------------------------------------------------------------
constexpr int kLnBwdWarps = 4;          // ~195 registers/thread: 4 warps -> 2 CTAs per SM

static inline int ln_bwd_rows_per_cta(int R) {
    // one wave: at most 2 CTAs per SM
    int rows = (R + 2 * kNumSMs - 1) / (2 * kNumSMs);
    rows = ((rows + kLnBwdWarps - 1) / kLnBwdWarps) * kLnBwdWarps;
    return rows < kLnBwdWarps ? kLnBwdWarps : rows;
}
static inline int ln_bwd_blocks(int R) {
    const int rows = ln_bwd_rows_per_cta(R);
    return (R + rows - 1) / rows;
}

extern "C" long long vct_ln_bwd_workspace_floats(int R, int d) { return (long long)ln_bwd_blocks(R) * 3 * d; }

// Column partials are kept in a "q-major" order inside each [d] vector: element (chunk c, q) of the 8-wide
// chunking sits at q * (d/8) + c, which makes every shared-memory access of the reduction conflict-free.
// ln_bwd_reduce_kernel undoes the permutation when it writes dgamma / dbeta / dbias.
template <int NV, typename TC>
__global__ void __launch_bounds__(kLnBwdWarps * 32)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ s, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, float* __restrict__ ds,
              TC* __restrict__ dr_c, float* __restrict__ partials, int R, int d, int rows_per_cta, float drop_p,
              const unsigned long long* __restrict__ rng_state, unsigned int site) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];  // [kLnBwdWarps][3][d]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = d >> 3;
    const Rng rng = make_rng(rng_state, drop_p);
    float acc_g[NV][8], acc_b[NV][8], acc_r[NV][8], gam[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc_g[i][q] = acc_b[i][q] = acc_r[i][q] = gam[i][q] = 0.f;
        if (c < nv) ld8(gamma + c * 8, gam[i]);
    }
    const int row_end = min(R, (blockIdx.x + 1) * rows_per_cta);
#pragma unroll 1
    for (int row = blockIdx.x * rows_per_cta + warp; row < row_end; row += kLnBwdWarps) {
        const long long base = (long long)row * d;
        // all loads of the row first (ONE memory round trip per row; written chunk by chunk the compiler emitted one batch
        // of loads per chunk, each behind the previous chunk's arithmetic), the Philox rounds while they are in flight
        float g[NV][8], xh[NV][8];              // raw dy / s, then dy * gamma / x-hat in place
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < nv) {
                ld8(dy + base + c * 8, g[i]);
                ld8(s + base + c * 8, xh[i]);
            }
        }
        const float mu = mean[row], rs = rstd[row];
        uint32_t keep[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            keep[i] = c < nv ? dropout_bits8(rng, site, (unsigned long long)(base >> 3) + c) : 0u;
        }
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < nv) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float dyv = g[i][q];
                    xh[i][q] = (xh[i][q] - mu) * rs;
                    g[i][q] = dyv * gam[i][q];
                    c1 += g[i][q];
                    c2 += g[i][q] * xh[i][q];
                    acc_g[i][q] += dyv * xh[i][q];
                    acc_b[i][q] += dyv;
                }
            }
        }
        c1 = warp_sum(c1) / (float)d;
        c2 = warp_sum(c2) / (float)d;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < nv) {
                float o[8], dr[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    o[q] = rs * (g[i][q] - c1 - xh[i][q] * c2);
                    dr[q] = o[q] * (((keep[i] >> q) & 1u) ? rng.inv_keep : 0.f);     // (p = 0: all bits set, inv_keep = 1)
                    acc_r[i][q] += dr[q];
                }
                if (ds) st8(ds + base + c * 8, o);
                if (dr_c) st8(dr_c + base + c * 8, dr);
            }
        }
    }
    // every warp parks its column sums in its own shared-memory slab (q-major, conflict-free) ...
    float* slab = sm + (size_t)warp * 3 * d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nv) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                slab[q * nv + c] = acc_g[i][q];
                slab[d + q * nv + c] = acc_b[i][q];
                slab[2 * d + q * nv + c] = acc_r[i][q];
            }
        }
    }
    __syncthreads();
    // ... and the CTA adds the slabs in a fixed order
    float* mine = partials + (long long)blockIdx.x * 3 * d;
    for (int i = threadIdx.x; i < 3 * d; i += blockDim.x) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kLnBwdWarps; ++w) a += sm[(size_t)w * 3 * d + i];
        mine[i] = a;
    }
}

// out[k][c*8+q] = sum_b partials[b][k*d + q*(d/8) + c]; 32 entries x 32 block-groups per CTA, fixed summation order
__global__ void __launch_bounds__(1024)
ln_bwd_reduce_kernel(const float* __restrict__ partials, int nblocks, int d, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dbias_r) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[32][33];
    const int cx = threadIdx.x & 31, gy = threadIdx.x >> 5;
    const int idx = blockIdx.x * 32 + cx;
    float a = 0.f;
    if (idx < 3 * d)
        for (int b = gy; b < nblocks; b += 32) a += partials[(long long)b * 3 * d + idx];
    red[gy][cx] = a;
    __syncthreads();
    if (gy == 0 && idx < 3 * d) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 32; ++g) t += red[g][cx];
        const int nv = d >> 3;
        const int k = idx / d, r = idx % d, q = r / nv, c = r % nv;
        const int col = c * 8 + q;
        float* out = k == 0 ? dgamma : (k == 1 ? dbeta : dbias_r);
        if (out) out[col] = t;
    }
}

template <typename TC>
static int launch_ln_bwd(const float* dy, const float* s, const float* mean, const float* rstd, const float* gamma,
                         float* ds, TC* dr_c, float* partials, int R, int d, float drop_p,
                         const unsigned long long* rng_state, unsigned int site, cudaStream_t st) {
    const int rows = ln_bwd_rows_per_cta(R), blocks = ln_bwd_blocks(R);
    const int nvl = (d / 8 + 31) / 32;
    const size_t smem = (size_t)kLnBwdWarps * 3 * d * sizeof(float);
#define LN_BWD_CASE(NVV)                                                                                        \
    {                                                                                                           \
        auto kern = ln_bwd_kernel<NVV, TC>;                                                                     \
        static bool once = false;                                                                               \
        if (!once) { VCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); once = true; } \
        vct::launch(kern, dim3(blocks), dim3(kLnBwdWarps * 32), smem, st, dy, s, mean, rstd, gamma, ds, dr_c, partials, R, d, rows, drop_p, \
                                                     rng_state, site);                                          \
    }
    if (nvl <= 1) LN_BWD_CASE(1)
    else if (nvl <= 2) LN_BWD_CASE(2)
    else if (nvl <= 3) LN_BWD_CASE(3)
    else LN_BWD_CASE(4)
#undef LN_BWD_CASE
    return check_launch("vct_ln_residual_bwd");
}

extern "C" int vct_ln_bwd_reduce(const float* partials, int R, int d, float* dgamma, float* dbeta, float* dbias_r,
                                 vct_stream_t stream) {
    VCT_REQUIRE(partials && d % 8 == 0 && d <= 1024 && R > 0, "vct_ln_bwd_reduce: bad arguments");
    vct::launch(ln_bwd_reduce_kernel, dim3((3 * d + 31) / 32), dim3(1024), 0, (cudaStream_t)stream, partials, ln_bwd_blocks(R), d, dgamma, dbeta,
                                                                              dbias_r);
    return check_launch("vct_ln_bwd_reduce");
}

extern "C" int vct_ln_residual_bwd(const float* dy, const float* s, const float* mean, const float* rstd,
                                   const float* gamma, float* ds, void* dr_c, int dr_dtype, float* dgamma,
                                   float* dbeta, float* dbias_r, float* partials, unsigned int* counter, int R, int d,
                                   float drop_p, const unsigned long long* rng_state, unsigned int site,
                                   vct_stream_t stream) {
    (void)counter;   // kept in the ABI; the two-kernel reduction needs no counter
    VCT_REQUIRE(d % 8 == 0 && d <= 1024 && R > 0, "vct_ln_residual_bwd: need d %% 8 == 0, d <= 1024 (d=%d)", d);
    VCT_REQUIRE(dy && s && mean && rstd && gamma && partials, "vct_ln_residual_bwd: null input");
    int e;
    if (dr_dtype == VCT_BF16)
        e = launch_ln_bwd<__nv_bfloat16>(dy, s, mean, rstd, gamma, ds, (__nv_bfloat16*)dr_c, partials, R, d, drop_p,
                                         rng_state, site, (cudaStream_t)stream);
    else
        e = launch_ln_bwd<float>(dy, s, mean, rstd, gamma, ds, (float*)dr_c, partials, R, d, drop_p, rng_state, site,
                                 (cudaStream_t)stream);
    if (e) return e;
    // all three outputs NULL: the caller runs vct_ln_bwd_reduce itself (e.g. on another stream)
    if (dgamma == nullptr && dbeta == nullptr && dbias_r == nullptr) return 0;
    return vct_ln_bwd_reduce(partials, R, d, dgamma, dbeta, dbias_r, stream);
}

// ------------------------------------------------------------------------------------------------
// embedding (8 elements per thread: one Philox draw covers them)
// ------------------------------------------------------------------------------------------------
template <typename TC>
__global__ void embed_fwd_kernel(const long long* __restrict__ ids, long long ids_ld, const float* __restrict__ E,
                                 const float* __restrict__ pos, float* __restrict__ x, TC* __restrict__ x_c, int B,
                                 int S, int d, int V, int pos_offset, float drop_p,
                                 const unsigned long long* __restrict__ rng_state, unsigned int site) {
    pdl_launch_dependents();
    pdl_wait();
    const int nv = d >> 3;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * S * nv) return;
    const int c = (int)(gid % nv);
    const long long row = gid / nv;
    const int b = (int)(row / S), sidx = (int)(row % S);
    long long id = ids[(long long)b * ids_ld + sidx];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    const Rng rng = make_rng(rng_state, drop_p);
    float e[8], p[8], sc[8], o[8];
    ld8(E + id * d + c * 8, e);
    ld8(pos + (long long)(sidx + pos_offset) * d + c * 8, p);
    dropout_scale8(rng, site, (unsigned long long)gid, sc);
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = (e[q] + p[q]) * sc[q];
    if (x) st8(x + row * d + c * 8, o);
    if (x_c) st8(x_c + row * d + c * 8, o);
}

extern "C" int vct_embed_fwd(const long long* ids, long long ids_ld, const float* E, const float* pos, float* x,
                             void* x_c, int x_c_dtype, int B, int S, int d, int V, int pos_offset, float drop_p,
                             const unsigned long long* rng_state, unsigned int site, vct_stream_t stream) {
    VCT_REQUIRE(d % 8 == 0 && B > 0 && S > 0, "vct_embed_fwd: need d %% 8 == 0 and non-empty input");
    long long n = (long long)B * S * (d / 8);
    int blocks = (int)((n + 255) / 256);
    if (x_c_dtype == VCT_BF16)
        vct::launch(embed_fwd_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, ids, ids_ld, E, pos, x, (__nv_bfloat16*)x_c, B, S, d,
                                                                   V, pos_offset, drop_p, rng_state, site);
    else
        vct::launch(embed_fwd_kernel<float>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, ids, ids_ld, E, pos, x, (float*)x_c, B, S, d, V,
                                                                   pos_offset, drop_p, rng_state, site);
    return check_launch("vct_embed_fwd");
}

__global__ void embed_bwd_kernel(const long long* __restrict__ ids, long long ids_ld, const float* __restrict__ dx,
                                 float* __restrict__ dE, int B, int S, int d, int V, int pad_id, float drop_p,
                                 const unsigned long long* __restrict__ rng_state, unsigned int site) {
    pdl_launch_dependents();
    pdl_wait();
    const int nv = d >> 3;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * S * nv) return;
    const int c = (int)(gid % nv);
    const long long row = gid / nv;
    const int b = (int)(row / S), sidx = (int)(row % S);
    const long long id = ids[(long long)b * ids_ld + sidx];
    if (id == pad_id || id < 0 || id >= V) return;
    const Rng rng = make_rng(rng_state, drop_p);
    float g[8], sc[8];
    ld8(dx + row * d + c * 8, g);
    dropout_scale8(rng, site, (unsigned long long)gid, sc);
    float* dst = dE + id * d + c * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) atomicAdd(dst + q, g[q] * sc[q]);
}

extern "C" int vct_embed_bwd(const long long* ids, long long ids_ld, const float* dx, float* dE, int B, int S, int d,
                             int V, int pad_id, float drop_p, const unsigned long long* rng_state, unsigned int site,
                             vct_stream_t stream) {
    VCT_REQUIRE(d % 8 == 0 && B > 0 && S > 0, "vct_embed_bwd: need d %% 8 == 0 and non-empty input");
    long long n = (long long)B * S * (d / 8);
    int blocks = (int)((n + 255) / 256);
    vct::launch(embed_bwd_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, ids, ids_ld, dx, dE, B, S, d, V, pad_id, drop_p,
                                                               rng_state, site);
    return check_launch("vct_embed_bwd");
}

// rows[r, :] = dx[r, :] * dropmask -- the embedding-table gradient in its SPARSE form (one row per token, no scatter):
// what the data-parallel trainer exchanges (all-gather of B*S rows + ids per rank, < 4 MB) instead of all-reducing the
// dense [V, d] table gradient (94 MB of which at most B*S rows are non-zero).  Same dropout index space as embed_bwd.
__global__ void embed_bwd_rows_kernel(const float* __restrict__ dx, float* __restrict__ rows, long long n8, float drop_p,
                                      const unsigned long long* __restrict__ rng_state, unsigned int site) {
    pdl_launch_dependents();
    pdl_wait();
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n8) return;
    const Rng rng = make_rng(rng_state, drop_p);
    float g[8], sc[8];
    ld8(dx + gid * 8, g);
    dropout_scale8(rng, site, (unsigned long long)gid, sc);
#pragma unroll
    for (int q = 0; q < 8; ++q) g[q] *= sc[q];
    st8(rows + gid * 8, g);
}

extern "C" int vct_embed_bwd_rows(const float* dx, float* rows, int B, int S, int d, float drop_p,
                                  const unsigned long long* rng_state, unsigned int site, vct_stream_t stream) {
    VCT_REQUIRE(d % 8 == 0 && B > 0 && S > 0, "vct_embed_bwd_rows: need d %% 8 == 0 and non-empty input");
    const long long n8 = (long long)B * S * (d / 8);
    vct::launch(embed_bwd_rows_kernel, dim3((unsigned)((n8 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, dx, rows, n8, drop_p,
                rng_state, site);
    return check_launch("vct_embed_bwd_rows");
}

// Deterministic scatter of gathered gradient rows into the table gradient.  Every data-parallel rank runs this on the SAME
// gathered (rows, ids) and must obtain bit-identical sums, otherwise the replicas drift apart (an atomicAdd scatter sums in
// a run-dependent order).  Step 1 (one CTA): keys (id << 16 | token index) of the n = B*S tokens, bitonic sort in shared
// memory (invalid / pad tokens sort to the end) -> tokens of the same id become one contiguous, index-ordered segment.
// Step 2 (one CTA per sorted position, only segment starts work): the segment's rows are summed by four row-chunks
// (chunk c takes segment elements c, c+4, ...; fixed order) and the four partials are added in chunk order; the row of dE
// is WRITTEN (rows no token maps to are left untouched, i.e. zero on the trainer's path).
constexpr int kDetMaxTokens = 16384;

__global__ void __launch_bounds__(1024)
embed_sort_kernel(const long long* __restrict__ ids, long long ids_ld, int B, int S, int V, int pad_id,
                  unsigned int* __restrict__ keys_out, int P) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ unsigned int skeys[];
    const int n = B * S;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        unsigned int key = 0xFFFFFFFFu;
        if (i < n) {
            const long long id = ids[(long long)(i / S) * ids_ld + (i % S)];
            if (id != pad_id && id >= 0 && id < V) key = ((unsigned int)id << 16) | (unsigned int)i;
        }
        skeys[i] = key;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int pos = 2 * t - (t & (stride - 1));
                const unsigned int a = skeys[pos], b = skeys[pos + stride];
                const bool up = (pos & size) == 0;
                if ((a > b) == up) { skeys[pos] = b; skeys[pos + stride] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) keys_out[i] = skeys[i];
}

__global__ void __launch_bounds__(1024)
embed_segment_sum_kernel(const unsigned int* __restrict__ keys, const float* __restrict__ rows, float* __restrict__ dE, int n,
                         int d) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ int s_len;
    extern __shared__ float4 s_part[];                     // [3][d / 4] partials of row-chunks 1..3
    const int i = blockIdx.x;
    const unsigned int key = keys[i];
    if (key == 0xFFFFFFFFu) return;
    const unsigned int id = key >> 16;
    if (i > 0 && (keys[i - 1] >> 16) == id) return;      // not the first token of its segment
    if (threadIdx.x == 0) s_len = n - i;
    __syncthreads();
    // segment length: first position whose id differs (all threads probe a window; the smallest hit wins)
    for (int base = i + 1; base < n; base += blockDim.x) {
        const int j = base + threadIdx.x;
        const bool hit = j < n && (keys[j] >> 16) != id;
        if (__syncthreads_or(hit)) {                       // block-uniform
            if (hit) atomicMin(&s_len, j - i);
            break;
        }
    }
    __syncthreads();
    const int len = s_len;
    const int nv = d >> 2;
    const int chunk = threadIdx.x / nv, c4 = threadIdx.x % nv;      // 4 row-chunks x d/4 float4 columns
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = chunk; j < len; j += 4) {
        const unsigned int tok = keys[i + j] & 0xFFFFu;
        const float4 v = ld4(rows + (long long)tok * d + c4 * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (chunk > 0) s_part[(chunk - 1) * nv + c4] = acc;
    __syncthreads();
    if (chunk == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 v = s_part[c * nv + c4];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        st4(dE + (long long)id * d + c4 * 4, acc);
    }
}

extern "C" int vct_embed_sort(const long long* ids, long long ids_ld, int B, int S, int V, int pad_id, unsigned int* keys_ws,
                              vct_stream_t stream) {
    const long long n = (long long)B * S;
    VCT_REQUIRE(ids && keys_ws && B > 0 && S > 0, "vct_embed_sort: null / empty argument");
    VCT_REQUIRE(n <= kDetMaxTokens && V <= 32768, "vct_embed_sort: at most %d tokens and V <= 32768 (got %lld, %d)", kDetMaxTokens, n, V);
    int P = 2;
    while (P < n) P <<= 1;
    static bool once = false;
    if (!once) {
        VCT_CUDA(cudaFuncSetAttribute(embed_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDetMaxTokens * 4));
        once = true;
    }
    vct::launch(embed_sort_kernel, dim3(1), dim3(1024), (size_t)P * 4, (cudaStream_t)stream, ids, ids_ld, B, S, V, pad_id, keys_ws, P);
    return check_launch("vct_embed_sort");
}

extern "C" int vct_embed_segment_sum(const unsigned int* keys, const float* rows, float* dE, int n, int d, vct_stream_t stream) {
    VCT_REQUIRE(keys && rows && dE && n > 0 && n <= kDetMaxTokens, "vct_embed_segment_sum: bad argument (n <= %d)", kDetMaxTokens);
    VCT_REQUIRE(d % 4 == 0 && d <= 1024, "vct_embed_segment_sum: need d %% 4 == 0 and d <= 1024");
    vct::launch(embed_segment_sum_kernel, dim3((unsigned)n), dim3(d), (size_t)3 * (d / 4) * 16, (cudaStream_t)stream, keys, rows, dE, n, d);
    return check_launch("vct_embed_segment_sum");
}

extern "C" int vct_embed_bwd_det(const long long* ids, long long ids_ld, const float* rows, float* dE, int B, int S, int d, int V,
                                 int pad_id, unsigned int* keys_ws, vct_stream_t stream) {
    VCT_REQUIRE(rows && dE, "vct_embed_bwd_det: null argument");
    if (int e = vct_embed_sort(ids, ids_ld, B, S, V, pad_id, keys_ws, stream)) return e;
    return vct_embed_segment_sum(keys_ws, rows, dE, B * S, d, stream);
}

// stamp[id] = current training step (rng_state[1]) for every token id of the batch: "this step touches row id of the
// embedding table".  No clearing pass: a row is touched in step t iff stamp[row] == t.
__global__ void embed_mark_kernel(const long long* __restrict__ ids, long long ids_ld, int B, int S, int V, int pad_id,
                                  unsigned int* __restrict__ stamp, const unsigned long long* __restrict__ rng_state) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * S) return;
    const long long id = ids[(long long)(i / S) * ids_ld + (i % S)];
    if (id == pad_id || id < 0 || id >= V) return;
    stamp[id] = (unsigned int)rng_state[1];
}

extern "C" int vct_embed_mark(const long long* ids, long long ids_ld, int B, int S, int V, int pad_id, unsigned int* stamp,
                              const unsigned long long* rng_state, vct_stream_t stream) {
    VCT_REQUIRE(ids && stamp && rng_state && B > 0 && S > 0, "vct_embed_mark: null / empty argument");
    vct::launch(embed_mark_kernel, dim3((unsigned)((B * S + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, ids, ids_ld, B, S, V, pad_id,
                stamp, rng_state);
    return check_launch("vct_embed_mark");
}

// dE[ids[b, s], :] = 0: re-zero exactly the rows a step scattered into (after the optimizer consumed them), instead of
// a 94 MB memset of the whole table gradient at the start of every step
__global__ void embed_zero_kernel(const long long* __restrict__ ids, long long ids_ld, float* __restrict__ dE, int B, int S,
                                  int d, int V) {
    pdl_launch_dependents();
    pdl_wait();
    const int nv = d >> 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * S * nv) return;
    const int c = (int)(gid % nv);
    const long long row = gid / nv;
    const int b = (int)(row / S), sidx = (int)(row % S);
    const long long id = ids[(long long)b * ids_ld + sidx];
    if (id < 0 || id >= V) return;
    st4(dE + id * d + c * 4, make_float4(0.f, 0.f, 0.f, 0.f));
}

extern "C" int vct_embed_zero(const long long* ids, long long ids_ld, float* dE, int B, int S, int d, int V,
                              vct_stream_t stream) {
    VCT_REQUIRE(d % 4 == 0 && B > 0 && S > 0, "vct_embed_zero: need d %% 4 == 0 and non-empty input");
    const long long n = (long long)B * S * (d / 4);
    vct::launch(embed_zero_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, ids, ids_ld, dE, B, S, d, V);
    return check_launch("vct_embed_zero");
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients): deterministic two-level reduction in one launch
// ------------------------------------------------------------------------------------------------
constexpr int kCsCols = 256, kCsRows = 64;

extern "C" long long vct_colsum_workspace_floats(int M, int N) {
    return (long long)((M + kCsRows - 1) / kCsRows) * N;
}

// A CTA owns kCsCols columns x kCsRows rows.  Each lane reads 8 consecutive columns (16 bytes of bf16) of a row, a warp
// therefore 512 contiguous bytes; the 8 warps take rows w, w + 8, ... and their partial sums are combined through
// shared memory in warp order (deterministic).  The last CTA of a column block adds the row-slice partials.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ X, long long ld, int M, int N, float* __restrict__ out,
              float* __restrict__ partials, unsigned int* counters, int vec) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ bool is_last;
    __shared__ float part[8][kCsCols + 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * kCsCols;
    const int m0 = blockIdx.y * kCsRows, m1 = min(M, m0 + kCsRows);
    const int nl = n0 + lane * 8;
    float a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = 0.f;
    if (vec && nl + 8 <= N) {
        for (int m = m0 + warp; m < m1; m += 8) {
            float v[8];
            ld8(X + (long long)m * ld + nl, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] += v[q];
        }
    } else {
        for (int m = m0 + warp; m < m1; m += 8)
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (nl + q < N) a[q] += to_f32(X[(long long)m * ld + nl + q]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) part[warp][lane * 8 + q] = a[q];
    __syncthreads();
    const int n = n0 + threadIdx.x;
    if (n < N) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
        partials[(long long)blockIdx.y * N + n] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev = atomicAdd(counters + blockIdx.x, 1u);
        is_last = (prev == gridDim.y - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        if (n < N) {
            float t = 0.f;
            for (unsigned int b = 0; b < gridDim.y; ++b) t += __ldcg(partials + (long long)b * N + n);
            out[n] = t;
        }
        if (threadIdx.x == 0) counters[blockIdx.x] = 0u;
    }
}

extern "C" int vct_colsum(const void* X, int dtype, long long ld, int M, int N, float* out, float* partials,
                          unsigned int* counter, vct_stream_t stream) {
    VCT_REQUIRE(M > 0 && N > 0 && N <= 65536, "vct_colsum: need 0 < N <= 65536 (counter array has 256 entries)");
    dim3 grid((N + kCsCols - 1) / kCsCols, (M + kCsRows - 1) / kCsRows);
    // 16-byte (bf16) / 2 x 16-byte (fp32) row pieces need 8-element aligned rows
    const int vec = (ld % 8 == 0) && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
    if (dtype == VCT_BF16)
        vct::launch(colsum_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)X, ld, M, N, out, partials, counter, vec);
    else
        vct::launch(colsum_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float*)X, ld, M, N, out, partials, counter, vec);
    return check_launch("vct_colsum");
}

// ------------------------------------------------------------------------------------------------
// Adam over the flat arena
// ------------------------------------------------------------------------------------------------
// Two independent 16-byte groups per array in flight per thread.  VCT_ADAM_CTAS_PER_SM caps the grid (default 8 = one
// full wave of 256-thread CTAs).  Measured on B200 inside the train step: 8 -> 1.949 ms/step, 2 -> 1.967, 1 -> 2.079:
// leaving thread slots free for the main lane does not pay, because the kernels of the step are individually
// bound by per-SM operand ingest / HBM and do not speed up when they share an SM.
__device__ __forceinline__ void adam_update4(float4& pv, const float4& gv, float4& mv, float4& vv, float grad_scale, float wd,
                                             float b1, float b2, float eps, float step_size, float inv_sqrt_bc2) {
    float pe[4] = {pv.x, pv.y, pv.z, pv.w}, ge[4] = {gv.x, gv.y, gv.z, gv.w};
    float me[4] = {mv.x, mv.y, mv.z, mv.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float gg = ge[k] * grad_scale;
        if (wd != 0.f) gg += wd * pe[k];
        me[k] = b1 * me[k] + (1.f - b1) * gg;
        ve[k] = b2 * ve[k] + (1.f - b2) * gg * gg;
        const float denom = sqrtf(ve[k]) * inv_sqrt_bc2 + eps;
        pe[k] -= step_size * (me[k] / denom);
    }
    pv = make_float4(pe[0], pe[1], pe[2], pe[3]);
    mv = make_float4(me[0], me[1], me[2], me[3]);
    vv = make_float4(ve[0], ve[1], ve[2], ve[3]);
}

template <typename TG>
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const TG* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            __nv_bfloat16* __restrict__ p_c, long long n4, const float* __restrict__ hyper, float grad_scale) {
    pdl_launch_dependents();
    pdl_wait();
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
    const float step_size = lr / hyper[6], inv_sqrt_bc2 = rsqrtf(hyper[7]);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n4; i += 2 * stride) {
        const long long j = i + stride;
        float4 p0 = ld4(p + i * 4), g0 = ld4(g + i * 4), m0 = ld4(m + i * 4), v0 = ld4(v + i * 4);
        float4 p1 = ld4(p + j * 4), g1 = ld4(g + j * 4), m1 = ld4(m + j * 4), v1 = ld4(v + j * 4);
        adam_update4(p0, g0, m0, v0, grad_scale, wd, b1, b2, eps, step_size, inv_sqrt_bc2);
        adam_update4(p1, g1, m1, v1, grad_scale, wd, b1, b2, eps, step_size, inv_sqrt_bc2);
        st4(p + i * 4, p0); st4(m + i * 4, m0); st4(v + i * 4, v0);
        st4(p + j * 4, p1); st4(m + j * 4, m1); st4(v + j * 4, v1);
        if (p_c) { st4(p_c + i * 4, p0); st4(p_c + j * 4, p1); }
    }
    if (i < n4) {
        float4 p0 = ld4(p + i * 4), g0 = ld4(g + i * 4), m0 = ld4(m + i * 4), v0 = ld4(v + i * 4);
        adam_update4(p0, g0, m0, v0, grad_scale, wd, b1, b2, eps, step_size, inv_sqrt_bc2);
        st4(p + i * 4, p0); st4(m + i * 4, m0); st4(v + i * 4, v0);
        if (p_c) st4(p_c + i * 4, p0);
    }
}

// Adam restricted to the rows of a [R, d] table that this step touched (touched = 1, gradient read from g) or did not
// touch (touched = 0, gradient is identically zero and is not read): the embedding table's 23 M parameters receive
// gradient in at most B*S rows per step, so all the other rows can be updated -- m, v decay, p moves along m -- before
// the step's backward has produced anything, off the critical tail of the step.
__global__ void __launch_bounds__(256)
adam_rows_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 __nv_bfloat16* __restrict__ p_c, long long n4, int d4, const float* __restrict__ hyper, float grad_scale,
                 const unsigned int* __restrict__ stamp, const unsigned long long* __restrict__ rng_state, int touched) {
    pdl_launch_dependents();
    pdl_wait();
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
    const float step_size = lr / hyper[6], inv_sqrt_bc2 = rsqrtf(hyper[7]);
    const unsigned int now = (unsigned int)rng_state[1];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const bool is_touched = stamp[i / d4] == now;
        if ((int)is_touched != touched) continue;
        float4 p0 = ld4(p + i * 4), m0 = ld4(m + i * 4), v0 = ld4(v + i * 4);
        float4 g0 = touched ? ld4(g + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        adam_update4(p0, g0, m0, v0, grad_scale, wd, b1, b2, eps, step_size, inv_sqrt_bc2);
        st4(p + i * 4, p0); st4(m + i * 4, m0); st4(v + i * 4, v0);
        if (p_c) st4(p_c + i * 4, p0);
    }
}

extern "C" int vct_adam_rows(float* p, const float* g, float* m, float* v, void* p_c, int R, int d, const float* hyper,
                             float grad_scale, const unsigned int* stamp, const unsigned long long* rng_state, int touched,
                             vct_stream_t stream) {
    VCT_REQUIRE(p && m && v && hyper && stamp && rng_state && R > 0 && d > 0 && d % 4 == 0, "vct_adam_rows: bad argument");
    VCT_REQUIRE(touched == 0 || g != nullptr, "vct_adam_rows: touched rows need their gradient");
    const long long n4 = (long long)R * (d / 4);
    long long want = (n4 + 255) / 256;
    const long long cap = (long long)kNumSMs * 8;
    const int blocks = (int)(want < cap ? want : cap);
    vct::launch(adam_rows_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, (__nv_bfloat16*)p_c, n4, d / 4, hyper,
                grad_scale, stamp, rng_state, touched);
    return check_launch("vct_adam_rows");
}

extern "C" int vct_adam(float* p, const void* g, int g_dtype, float* m, float* v, void* p_c, long long n, const float* hyper,
                        float grad_scale, vct_stream_t stream) {
    VCT_REQUIRE(n > 0 && n % 4 == 0, "vct_adam: arena length must be a positive multiple of 4 (n=%lld)", n);
    VCT_REQUIRE(g_dtype == VCT_F32 || g_dtype == VCT_BF16, "vct_adam: bad gradient dtype");
    static const int ctas_per_sm = getenv("VCT_ADAM_CTAS_PER_SM") ? atoi(getenv("VCT_ADAM_CTAS_PER_SM")) : 8;
    const long long n4 = n / 4;
    long long want = (n4 + 511) / 512;
    const long long cap = (long long)kNumSMs * (ctas_per_sm > 0 ? ctas_per_sm : 8);
    int blocks = (int)(want < cap ? want : cap);
    if (blocks < 1) blocks = 1;
    if (g_dtype == VCT_BF16)
        vct::launch(adam_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, p, (const __nv_bfloat16*)g, m, v,
                    (__nv_bfloat16*)p_c, n4, hyper, grad_scale);
    else
        vct::launch(adam_kernel<float>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, p, (const float*)g, m, v, (__nv_bfloat16*)p_c,
                    n4, hyper, grad_scale);
    return check_launch("vct_adam");
}

// ------------------------------------------------------------------------------------------------
// cast
// ------------------------------------------------------------------------------------------------
template <typename TO>
__global__ void cast_kernel(const float* __restrict__ src, TO* __restrict__ dst, long long n) {
    pdl_launch_dependents();
    pdl_wait();
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        st4(dst + i * 4, ld4(src + i * 4));
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[n4 * 4 + threadIdx.x] = from_f32<TO>(src[n4 * 4 + threadIdx.x]);
}

extern "C" int vct_cast(const float* src, void* dst, int dst_dtype, long long n, vct_stream_t stream) {
    VCT_REQUIRE(n > 0, "vct_cast: empty");
    long long want = (n / 4 + 255) / 256 + 1;
    int blocks = (int)(want < (long long)kNumSMs * 8 ? want : (long long)kNumSMs * 8);
    if (dst_dtype == VCT_BF16)
        vct::launch(cast_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, src, (__nv_bfloat16*)dst, n);
    else
        vct::launch(cast_kernel<float>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, src, (float*)dst, n);
    return check_launch("vct_cast");
}

// ------------------------------------------------------------------------------------------------
// zero the rows of x [R, d] whose mask byte is set (eval-mode fast path of nn.TransformerEncoder)
// ------------------------------------------------------------------------------------------------
__global__ void zero_rows_kernel(float* __restrict__ x, const unsigned char* __restrict__ mask, int R, int d) {
    pdl_launch_dependents();
    pdl_wait();
    const int nv = d >> 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)R * nv) return;
    const int r = (int)(i / nv);
    if (mask[r]) st4(x + i * 4, make_float4(0.f, 0.f, 0.f, 0.f));
}

extern "C" int vct_zero_rows(float* x, const unsigned char* mask, int R, int d, vct_stream_t stream) {
    VCT_REQUIRE(x && mask && R > 0 && d > 0 && d % 4 == 0, "vct_zero_rows: bad argument (d %% 4 must be 0)");
    const long long n = (long long)R * (d / 4);
    vct::launch(zero_rows_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, x, mask, R, d);
    return check_launch("vct_zero_rows");
}

// ------------------------------------------------------------------------------------------------
// input staging: ONE launch moves a step's device-resident batch into the workspace the launch plans read
//   feats  [B,T,Din] fp32 -> copy                                    (feats_src NULL: skip feats and vid_pad)
//   vid_pad [B,T] bytes (NULL = nothing padded) -> [B,T+1], column 0 (the global token) never padded,
//           model/MMEncoder.py:252-260
//   ids    [B,S1] int64 -> copy; tok_pad [B,S1-1] = the caller's mask, or ids[:, :-1] == pad_id as
//           model/CapPreprocessor.py:35 builds it               (ids_src NULL: skip)
// (five torch copy / fill / compare launches before: ~20 us of host time per step on the e2e path)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stage_inputs_kernel(const float4* __restrict__ feats_src, float4* __restrict__ feats_dst, long long n4,
                    const unsigned char* __restrict__ vid_src, unsigned char* __restrict__ vid_dst, int B, int T,
                    const long long* __restrict__ ids_src, long long* __restrict__ ids_dst,
                    const unsigned char* __restrict__ tok_src, unsigned char* __restrict__ tok_dst, int S1, long long pad_id) {
    // (a plain launch, no programmatic dependent launch: this kernel sits between two steps' CUDA graphs, outside any graph)
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (feats_src != nullptr) {
        for (long long i = tid; i < n4; i += stride) feats_dst[i] = feats_src[i];
        const long long nv = (long long)B * (T + 1);
        for (long long i = tid; i < nv; i += stride) {
            const long long b = i / (T + 1);
            const int m = (int)(i % (T + 1));
            vid_dst[i] = (m == 0 || vid_src == nullptr) ? (unsigned char)0 : (unsigned char)(vid_src[b * T + m - 1] != 0);
        }
    }
    if (ids_src != nullptr) {
        const long long ni = (long long)B * S1;
        for (long long i = tid; i < ni; i += stride) {
            const long long v = ids_src[i];
            ids_dst[i] = v;
            const long long b = i / S1;
            const int sidx = (int)(i % S1);
            if (sidx < S1 - 1) {
                const long long o = b * (S1 - 1) + sidx;
                tok_dst[o] = tok_src != nullptr ? (unsigned char)(tok_src[o] != 0) : (unsigned char)(v == pad_id);
            }
        }
    }
}

extern "C" int vct_stage_inputs(const float* feats_src, float* feats_dst, const unsigned char* vid_src, unsigned char* vid_dst,
                                int B, int T, int Din, const long long* ids_src, long long* ids_dst,
                                const unsigned char* tok_src, unsigned char* tok_dst, int S1, long long pad_id,
                                vct_stream_t stream) {
    VCT_REQUIRE(B > 0 && (feats_src != nullptr || ids_src != nullptr), "vct_stage_inputs: nothing to stage");
    long long n4 = 0, work = 0;
    if (feats_src != nullptr) {
        VCT_REQUIRE(feats_dst && vid_dst && T > 0 && Din > 0 && Din % 4 == 0, "vct_stage_inputs: bad feature arguments (Din %% 4 must be 0)");
        VCT_REQUIRE(((reinterpret_cast<uintptr_t>(feats_src) | reinterpret_cast<uintptr_t>(feats_dst)) & 15) == 0,
                    "vct_stage_inputs: feature buffers must be 16-byte aligned");
        n4 = (long long)B * T * (Din / 4);
        work = n4;
    }
    if (ids_src != nullptr) {
        VCT_REQUIRE(ids_dst && tok_dst && S1 >= 2, "vct_stage_inputs: bad id arguments");
        if ((long long)B * S1 > work) work = (long long)B * S1;
    }
    long long blocks = (work + 255) / 256;
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    stage_inputs_kernel<<<dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream>>>((const float4*)feats_src, (float4*)feats_dst, n4, vid_src,
                                                                                        vid_dst, B, T, ids_src, ids_dst, tok_src, tok_dst, S1,
                                                                                        pad_id);
    return check_launch("vct_stage_inputs");
}

// ------------------------------------------------------------------------------------------------
// greedy argmax + append
// ------------------------------------------------------------------------------------------------
// One CTA (512 threads) per row; the row is read as float4 (ld % 4 == 0, 16-byte aligned base) with four loads in flight
// per thread, ties resolved towards the lowest index at every level (torch.max semantics, model/MMT4Caption.py:165).
__device__ __forceinline__ void argmax_take(float v, int i, float& best, int& bi) {
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
}

__global__ void __launch_bounds__(512)
argmax_append_kernel(const float* __restrict__ logits, long long ld, int V, long long* __restrict__ ys, long long ys_ld,
                     int t, int end_id, int* __restrict__ ended, int* __restrict__ n_ended, int vec) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sv[16];
    __shared__ int si[16];
    const int b = blockIdx.x;
    const float* row = logits + (long long)b * ld;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    if (vec) {
        const float4* row4 = reinterpret_cast<const float4*>(row);
        const int n4 = V >> 2;
        int i = threadIdx.x;
        for (; i + 3 * (int)blockDim.x < n4; i += 4 * blockDim.x) {
            float4 a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = __ldcs(row4 + i + u * blockDim.x);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = 4 * (i + u * blockDim.x);
                argmax_take(a[u].x, c, best, bi); argmax_take(a[u].y, c + 1, best, bi);
                argmax_take(a[u].z, c + 2, best, bi); argmax_take(a[u].w, c + 3, best, bi);
            }
        }
        for (; i < n4; i += blockDim.x) {
            const float4 a = __ldcs(row4 + i);
            argmax_take(a.x, 4 * i, best, bi); argmax_take(a.y, 4 * i + 1, best, bi);
            argmax_take(a.z, 4 * i + 2, best, bi); argmax_take(a.w, 4 * i + 3, best, bi);
        }
        for (int c = (n4 << 2) + threadIdx.x; c < V; c += blockDim.x) argmax_take(row[c], c, best, bi);
    } else {
        for (int c = threadIdx.x; c < V; c += blockDim.x) argmax_take(row[c], c, best, bi);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        argmax_take(ov, oi, best, bi);
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) argmax_take(sv[w], si[w], best, bi);
        if (bi == 0x7fffffff) bi = 0;
        ys[(long long)b * ys_ld + t] = bi;
        if (bi == end_id && ended && !ended[b]) {
            ended[b] = 1;
            if (n_ended) atomicAdd(n_ended, 1);
        }
    }
}

extern "C" int vct_argmax_append(const float* logits, long long ld_logits, int B, int V, long long* ys, long long ys_ld,
                                 int t, int end_id, int* ended, int* n_ended, vct_stream_t stream) {
    VCT_REQUIRE(B > 0 && V > 0 && logits && ys, "vct_argmax_append: bad arguments");
    const int vec = (ld_logits % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) ? 1 : 0;
    vct::launch(argmax_append_kernel, dim3(B), dim3(512), 0, (cudaStream_t)stream, logits, ld_logits, V, ys, ys_ld, t, end_id, ended, n_ended, vec);
    return check_launch("vct_argmax_append");
}

// ------------------------------------------------------------------------------------------------
// debug: dropout keep mask
// ------------------------------------------------------------------------------------------------
__global__ void dropout_mask_kernel(unsigned char* out, long long n, float p, const unsigned long long* rng_state,
                                    unsigned int site) {
    pdl_launch_dependents();
    pdl_wait();
    const Rng rng = make_rng(rng_state, p);
    long long i8 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i8 * 8 >= n) return;
    float sc[8];
    dropout_scale8(rng, site, (unsigned long long)i8, sc);
    for (int k = 0; k < 8; ++k)
        if (i8 * 8 + k < n) out[i8 * 8 + k] = sc[k] != 0.f;
}

extern "C" int vct_dropout_mask(unsigned char* out, long long n, float drop_p, const unsigned long long* rng_state,
                                unsigned int site, vct_stream_t stream) {
    VCT_REQUIRE(n > 0, "vct_dropout_mask: empty");
    long long n8 = (n + 7) / 8;
    vct::launch(dropout_mask_kernel, dim3((int)((n8 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, out, n, drop_p, rng_state, site);
    return check_launch("vct_dropout_mask");
}
