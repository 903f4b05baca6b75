// Data-parallel gradient exchange over NVLink peer memory (replaces DDP's bucketed NCCL all-reduce, train.py:218).
//
// One process per GPU.  Every rank allocates ONE communication region with cudaMalloc, exports it with cudaIpc, and maps
// the regions of all peers (NVSwitch: every pair is P2P-capable).  The collectives are ordinary kernels that read the
// peers' regions with 16-byte loads and synchronise through flag words written straight into the peers' memory, so they
//   * are CUDA-graph capturable -- the whole N-GPU train step is ONE graph like the single-GPU step (NCCL launches had to
//     stay outside the graphs on this stack, which left the multi-GPU step bound by ~35 host operations per step), and
//   * move each byte over NVLink once per direction: two-shot all-reduce = every rank receives the other ranks' values of
//     its 1/N chunk, reduces it (fp32 sum of the N bf16 values in rank order, ONE rounding) and pushes the result to all
//     peers.  The sum is computed once and copied, so all replicas hold bit-identical gradients.
//
// Synchronisation: CTA b of rank r only ever waits for CTA b of the other ranks (same grid size everywhere), through a
// monotone counter per (channel, source rank, CTA) in the destination's flag array: barrier k is passed once every peer's
// counter is >= k.  Counters live in device memory and advance by two per collective, so a captured graph can be
// replayed forever.  A spin that lasts longer than ~20 s sets a status word and gives up (no hung GPU box).
#include "common.cuh"

using namespace vct;

namespace {

constexpr int kMaxWorld = 8;
constexpr int kChannels = 4;
constexpr int kMaxCtas = 64;
constexpr int kThreads = 512;

struct CommDev {
    int world, rank, ctas;
    char* buf[kMaxWorld];           // data region of every rank (peer-mapped; buf[rank] is local)
    unsigned int* flags[kMaxWorld]; // flag array of every rank: [kChannels][kMaxWorld][kMaxCtas]
    unsigned int* epoch;            // local: [kChannels][kMaxCtas] barriers passed so far
    unsigned int* status;           // local: != 0 after a barrier timed out
};

struct CommHost {
    CommDev dev;
    int device;
    size_t bytes;                   // data bytes
    char* base;                     // local allocation: [data | flags | epoch | status]
    void* opened[kMaxWorld];        // cudaIpcOpenMemHandle results to close
    bool connected, in_process;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr size_t kFlagBytes = sizeof(unsigned int) * kChannels * kMaxWorld * kMaxCtas;
constexpr size_t kEpochBytes = sizeof(unsigned int) * kChannels * kMaxCtas;

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// peer data: never through L1 (the local L1 is the only cache that holds peer lines, B300_MICROARCH "NVLink")
__device__ __forceinline__ uint4 ld_peer16(const void* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// every rank's CTA `b` has reached barrier `target` on `channel`
__device__ __forceinline__ void peer_barrier(const CommDev& c, int channel, unsigned int target) {
    __syncthreads();
    const int b = blockIdx.x;
    if ((int)threadIdx.x < c.world) {
        const int p = threadIdx.x;
        __threadfence_system();
        st_release_sys(c.flags[p] + ((size_t)channel * kMaxWorld + c.rank) * kMaxCtas + b, target);
        const unsigned int* mine = c.flags[c.rank] + ((size_t)channel * kMaxWorld + p) * kMaxCtas + b;
        const long long t0 = clock64();
        // (once a barrier has timed out every later one gives up after a short wait: the run is lost, do not hang the box)
        const long long limit = *reinterpret_cast<volatile unsigned int*>(c.status) ? 2000000LL : 40000000000LL;
        while ((int)(ld_acquire_sys(mine) - target) < 0) {
            if (clock64() - t0 > limit) {                  // ~20 s at 1.9 GHz: a peer died or the ranks diverged
                atomicExch(c.status, 1u);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void add8(float (&acc)[8], const uint4& v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        acc[2 * i] += f.x;
        acc[2 * i + 1] += f.y;
    }
}

// local data that a peer has written (staging) or that this GPU wrote earlier: through L2, never a stale L1 line
__device__ __forceinline__ uint4 ld_cg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

// In-place SUM all-reduce of n_vec 16-byte vectors (8 bf16 each) at byte offset `off` of every rank's region, PUSH model:
// all NVLink traffic is stores (posted, bandwidth-bound), all loads are local.  A pull version of this kernel reached
// 190 GB/s with 32 CTAs -- remote loads are latency-bound (~2 us round trip, few requests in flight per SM).
//   1. scatter   for every peer p: my values of chunk p  ->  p's staging area, slot `rank`
//      barrier   (everyone's contributions to my chunk have landed in my staging area)
//   2. reduce    my chunk = sum over ranks in RANK ORDER (own values from my buffer, the others from staging), fp32, one
//                rounding; the result goes to my buffer and is pushed into the same place of every peer's buffer
//      barrier   (every chunk of my buffer has been filled in by its owner)
// Staging (byte offset `stage`): [world][per] vectors, per = ceil(n_vec / world).  CTA b of every rank owns the same
// vectors of every chunk, so CTA b only ever synchronises with the peers' CTA b.
// (No griddepcontrol.launch_dependents: the dependent -- Adam, 8 CTAs per SM -- would become resident at once and sit in
// griddepcontrol.wait for the whole exchange, starving the backward kernels that run beside it.)
__global__ void __launch_bounds__(kThreads)
peer_allreduce_bf16_kernel(CommDev c, long long off, long long stage, long long n_vec, int channel) {
    pdl_wait();
    const int b = blockIdx.x, G = gridDim.x;
    const unsigned int e = c.epoch[channel * kMaxCtas + b];
    const long long per = (n_vec + c.world - 1) / c.world;
    const char* mine = c.buf[c.rank] + off;
    // ---- 1. scatter --------------------------------------------------------------------------------------------------
    for (int q = 1; q < c.world; ++q) {
        const int p = (c.rank + q) % c.world;                // a different first destination on every rank
        const long long lo = (long long)p * per, hi = min(n_vec, lo + per);
        char* dst = c.buf[p] + stage + (long long)c.rank * per * 16;
        for (long long v0 = lo + (long long)b * kThreads; v0 < hi; v0 += (long long)G * kThreads * 4) {
            uint4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long v = v0 + (long long)u * G * kThreads + threadIdx.x;
                if (v < hi) x[u] = ld_cg16(mine + v * 16);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long v = v0 + (long long)u * G * kThreads + threadIdx.x;
                if (v < hi) *reinterpret_cast<uint4*>(dst + (v - lo) * 16) = x[u];
            }
        }
    }
    peer_barrier(c, channel, e + 1);
    // ---- 2. reduce my chunk and broadcast it -------------------------------------------------------------------------
    {
        const long long lo = (long long)c.rank * per, hi = min(n_vec, lo + per);
        const char* st = c.buf[c.rank] + stage;
        for (long long v0 = lo + (long long)b * kThreads; v0 < hi; v0 += (long long)G * kThreads * 2) {
            uint4 x[2][kMaxWorld];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long v = v0 + (long long)u * G * kThreads + threadIdx.x;
#pragma unroll
                for (int p = 0; p < kMaxWorld; ++p)
                    if (p < c.world && v < hi)
                        x[u][p] = ld_cg16(p == c.rank ? mine + v * 16 : st + ((long long)p * per + (v - lo)) * 16);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long v = v0 + (long long)u * G * kThreads + threadIdx.x;
                if (v < hi) {
                    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int p = 0; p < kMaxWorld; ++p)
                        if (p < c.world) add8(acc, x[u][p]);
                    uint4 r;
                    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
                    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
#pragma unroll
                    for (int p = 0; p < kMaxWorld; ++p)
                        if (p < c.world) *reinterpret_cast<uint4*>(c.buf[p] + off + v * 16) = r;
                }
            }
        }
    }
    peer_barrier(c, channel, e + 2);
    if (threadIdx.x == 0) c.epoch[channel * kMaxCtas + b] = e + 2;
}

// All-gather (push): the region at `off` is [world][slot_vec vectors]; rank r has filled slot r of its OWN region and
// writes it into slot r of every peer's region.  The first barrier keeps a fast rank from overwriting a slot whose
// previous contents a slower peer is still consuming.
__global__ void __launch_bounds__(kThreads)
peer_allgather_kernel(CommDev c, long long off, long long slot_vec, int channel) {
    pdl_wait();
    const int b = blockIdx.x, G = gridDim.x;
    const unsigned int e = c.epoch[channel * kMaxCtas + b];
    peer_barrier(c, channel, e + 1);
    const char* src = c.buf[c.rank] + off + (long long)c.rank * slot_vec * 16;
    for (long long v0 = (long long)b * kThreads; v0 < slot_vec; v0 += (long long)G * kThreads * 4) {
        uint4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long v = v0 + (long long)u * G * kThreads + threadIdx.x;
            if (v < slot_vec) x[u] = ld_cg16(src + v * 16);
        }
        for (int q = 1; q < c.world; ++q) {
            char* dst = c.buf[(c.rank + q) % c.world] + off + (long long)c.rank * slot_vec * 16;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long v = v0 + (long long)u * G * kThreads + threadIdx.x;
                if (v < slot_vec) *reinterpret_cast<uint4*>(dst + v * 16) = x[u];
            }
        }
    }
    peer_barrier(c, channel, e + 2);
    if (threadIdx.x == 0) c.epoch[channel * kMaxCtas + b] = e + 2;
}

CommHost* H(void* h) { return reinterpret_cast<CommHost*>(h); }

}  // namespace

extern "C" int vct_comm_create(int rank, int world, long long bytes, int ctas, void** handle_out) {
    VCT_REQUIRE(handle_out && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "vct_comm_create: bad rank / world (<= %d ranks)", kMaxWorld);
    VCT_REQUIRE(bytes > 0 && bytes % 256 == 0, "vct_comm_create: bytes must be a positive multiple of 256");
    VCT_REQUIRE(ctas >= 1 && ctas <= kMaxCtas, "vct_comm_create: ctas must be in [1, %d]", kMaxCtas);
    CommHost* c = new CommHost();
    memset(c, 0, sizeof(*c));
    VCT_CUDA(cudaGetDevice(&c->device));
    c->bytes = (size_t)bytes;
    const size_t total = c->bytes + kFlagBytes + kEpochBytes + 256;
    VCT_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->base), total));
    VCT_CUDA(cudaMemset(c->base, 0, total));
    VCT_CUDA(cudaDeviceSynchronize());
    c->dev.world = world; c->dev.rank = rank; c->dev.ctas = ctas;
    c->dev.buf[rank] = c->base;
    c->dev.flags[rank] = reinterpret_cast<unsigned int*>(c->base + c->bytes);
    c->dev.epoch = reinterpret_cast<unsigned int*>(c->base + c->bytes + kFlagBytes);
    c->dev.status = reinterpret_cast<unsigned int*>(c->base + c->bytes + kFlagBytes + kEpochBytes);
    c->connected = world == 1;
    *handle_out = c;
    return 0;
}

extern "C" void* vct_comm_base(void* handle) { return handle ? H(handle)->base : nullptr; }

extern "C" int vct_comm_ipc_handle(void* handle, unsigned char* out64) {
    VCT_REQUIRE(handle && out64, "vct_comm_ipc_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    VCT_CUDA(cudaIpcGetMemHandle(&h, H(handle)->base));
    memcpy(out64, &h, 64);
    return 0;
}

// all_handles: world x 64 bytes, rank-major (what every rank obtained from vct_comm_ipc_handle, all-gathered by the host)
extern "C" int vct_comm_connect(void* handle, const unsigned char* all_handles) {
    VCT_REQUIRE(handle && all_handles, "vct_comm_connect: null argument");
    CommHost* c = H(handle);
    for (int p = 0; p < c->dev.world; ++p) {
        if (p == c->dev.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + 64 * p, 64);
        void* ptr = nullptr;
        VCT_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->opened[p] = ptr;
        c->dev.buf[p] = reinterpret_cast<char*>(ptr);
        c->dev.flags[p] = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ptr) + c->bytes);
    }
    c->connected = true;
    return 0;
}

// Several ranks inside ONE process (tests): peer_bases[p] = vct_comm_base of rank p's handle; peer access is enabled here.
extern "C" int vct_comm_connect_in_process(void* handle, void* const* peer_bases, const int* peer_devices) {
    VCT_REQUIRE(handle && peer_bases && peer_devices, "vct_comm_connect_in_process: null argument");
    CommHost* c = H(handle);
    VCT_CUDA(cudaSetDevice(c->device));
    for (int p = 0; p < c->dev.world; ++p) {
        if (p == c->dev.rank) continue;
        int can = 0;
        VCT_CUDA(cudaDeviceCanAccessPeer(&can, c->device, peer_devices[p]));
        VCT_REQUIRE(can, "vct_comm_connect_in_process: device %d cannot access device %d", c->device, peer_devices[p]);
        cudaError_t e = cudaDeviceEnablePeerAccess(peer_devices[p], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) VCT_CUDA(e);
        (void)cudaGetLastError();
        c->dev.buf[p] = reinterpret_cast<char*>(peer_bases[p]);
        c->dev.flags[p] = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(peer_bases[p]) + c->bytes);
    }
    c->connected = true;
    c->in_process = true;
    return 0;
}

extern "C" int vct_peer_allreduce_bf16(void* handle, long long byte_off, long long n_elems, long long stage_off, int channel,
                                       vct_stream_t stream) {
    VCT_REQUIRE(handle, "vct_peer_allreduce_bf16: null handle");
    CommHost* c = H(handle);
    VCT_REQUIRE(c->connected, "vct_peer_allreduce_bf16: vct_comm_connect has not been called");
    VCT_REQUIRE(byte_off >= 0 && byte_off % 16 == 0 && n_elems > 0 && n_elems % 8 == 0 && (size_t)(byte_off + n_elems * 2) <= c->bytes,
                "vct_peer_allreduce_bf16: range must be 16-byte aligned, a multiple of 8 elements and inside the region");
    VCT_REQUIRE(channel >= 0 && channel < kChannels, "vct_peer_allreduce_bf16: channel must be in [0, %d)", kChannels);
    {
        const long long per = (n_elems / 8 + c->dev.world - 1) / c->dev.world;
        const long long stage_bytes = per * 16 * c->dev.world;
        VCT_REQUIRE(stage_off >= 0 && stage_off % 16 == 0 && (size_t)(stage_off + stage_bytes) <= c->bytes &&
                        (stage_off + stage_bytes <= byte_off || stage_off >= byte_off + n_elems * 2),
                    "vct_peer_allreduce_bf16: the staging area (%lld bytes at %lld) must lie inside the region and not overlap the data",
                    stage_bytes, stage_off);
    }
    vct::launch(peer_allreduce_bf16_kernel, dim3(c->dev.ctas), dim3(kThreads), 0, (cudaStream_t)stream, c->dev, byte_off, stage_off,
                n_elems / 8, channel);
    return check_launch("vct_peer_allreduce_bf16");
}

extern "C" int vct_peer_allgather(void* handle, long long byte_off, long long slot_bytes, int channel, vct_stream_t stream) {
    VCT_REQUIRE(handle, "vct_peer_allgather: null handle");
    CommHost* c = H(handle);
    VCT_REQUIRE(c->connected, "vct_peer_allgather: vct_comm_connect has not been called");
    VCT_REQUIRE(byte_off >= 0 && byte_off % 16 == 0 && slot_bytes > 0 && slot_bytes % 16 == 0 &&
                    (size_t)(byte_off + slot_bytes * c->dev.world) <= c->bytes,
                "vct_peer_allgather: slots must be 16-byte multiples inside the region");
    VCT_REQUIRE(channel >= 0 && channel < kChannels, "vct_peer_allgather: channel must be in [0, %d)", kChannels);
    vct::launch(peer_allgather_kernel, dim3(c->dev.ctas), dim3(kThreads), 0, (cudaStream_t)stream, c->dev, byte_off, slot_bytes / 16, channel);
    return check_launch("vct_peer_allgather");
}

// 0 = healthy, 1 = a barrier timed out (results are garbage).  Synchronises the device.
extern "C" int vct_comm_status(void* handle) {
    VCT_REQUIRE(handle, "vct_comm_status: null handle");
    CommHost* c = H(handle);
    unsigned int s = 0;
    VCT_CUDA(cudaMemcpy(&s, c->dev.status, sizeof(s), cudaMemcpyDeviceToHost));
    return (int)s;
}

extern "C" int vct_comm_destroy(void* handle) {
    if (!handle) return 0;
    CommHost* c = H(handle);
    if (!c->in_process)
        for (int p = 0; p < c->dev.world; ++p)
            if (c->opened[p]) cudaIpcCloseMemHandle(c->opened[p]);
    cudaFree(c->base);
    delete c;
    (void)cudaGetLastError();
    return 0;
}
