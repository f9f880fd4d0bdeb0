// Gradient all-reduce over NVLink / NVSwitch PEER MEMORY (the one exchange of the path -- the sum of the flat
// parameter-cotangent vector over the ranks, once per optimiser step).
//
// Every rank's flat gradient lives in a buffer that all GPUs of the box have mapped (torch symmetric memory hands out the
// peer pointers; the C ABI only sees raw pointers).  ONE kernel per rank does the whole collective, two-shot, in place:
//   barrier 1   "my vector is complete" -> every peer's flag pad (st.release.sys), wait for all peers' flags;
//   slice r     rank r owns the r-th 1/W of the vector: it loads that slice from ALL W buffers (remote 128-bit loads over
//               NVLink, W loads in flight per thread), adds them in rank order 0..W-1, scales, and stores the result into
//               ALL W buffers (remote 128-bit stores).  Every element is reduced by exactly one rank, so all ranks end up
//               with bit-identical sums and the summation order is fixed (run-to-run deterministic);
//   barrier 2   the last block of the rank to finish signals "my stores are done"; the kernel returns once all peers have.
// Per GPU and direction the links carry 2 (W-1)/W n floats (slice loads served + slice results received), the minimum of
// a reduce-scatter + all-gather; there is no staging copy and no host synchronisation.  The flags are epoch counters
// (never reset): the caller passes a number that grows by one per call, the same on every rank.
#include <stdint.h>
#include "phx_common.cuh"

namespace {

constexpr int PEER_MAXW = 8;
constexpr int PEER_THREADS = 1024;

struct PeerParams {
    float* buf[PEER_MAXW];
    unsigned* flag[PEER_MAXW];   // flag[r]: rank r's pad, 2 * world words: [0, W) "data ready", [W, 2W) "stores done"
    int rank, world;
    unsigned long long n;        // floats
    unsigned epoch;
    float scale;
    unsigned* counter;           // block counter of this rank (device-local, zero between calls)
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// peer data must come from the owner's memory, never from a local cache line of an earlier step
__device__ __forceinline__ float4 ld_peer4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ bool epoch_reached(unsigned seen, unsigned epoch) { return (int)(seen - epoch) >= 0; }

template <int W>
__global__ void __launch_bounds__(PEER_THREADS, 1) peer_allreduce_kernel(const PeerParams p) {
    // ---- barrier 1 ----
    if (blockIdx.x == 0 && threadIdx.x < W) st_release_sys(p.flag[threadIdx.x] + p.rank, p.epoch);
    if (threadIdx.x < W)
        while (!epoch_reached(ld_acquire_sys(p.flag[p.rank] + threadIdx.x), p.epoch)) {}
    __syncthreads();
    // ---- this rank's slice, in float4 units ----
    const unsigned long long n4 = p.n >> 2, per = (n4 + W - 1) / W;
    const unsigned long long lo = per * (unsigned long long)p.rank, hi = lo + per < n4 ? lo + per : n4;
    // U float4 per thread and iteration so that U * W = 8 remote 16-byte loads are in flight per thread (the loads of a
    // small world would otherwise be latency-bound: one NVLink round trip per iteration)
    constexpr int U = 8 / W;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i0 = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += U * stride) {
        float4 v[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned long long i = i0 + u * stride;
            if (i < hi) {
#pragma unroll
                for (int r = 0; r < W; ++r) v[u][r] = ld_peer4(p.buf[r] + 4 * i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned long long i = i0 + u * stride;
            if (i < hi) {
                float4 a = v[u][0];
#pragma unroll
                for (int r = 1; r < W; ++r) {
                    a.x += v[u][r].x;
                    a.y += v[u][r].y;
                    a.z += v[u][r].z;
                    a.w += v[u][r].w;
                }
                a.x *= p.scale; a.y *= p.scale; a.z *= p.scale; a.w *= p.scale;
#pragma unroll
                for (int r = 0; r < W; ++r) *reinterpret_cast<float4*>(p.buf[r] + 4 * i) = a;
            }
        }
    }
    // the n % 4 tail elements: last rank, one thread each
    if (p.rank == W - 1 && blockIdx.x == 0 && threadIdx.x < (unsigned)(p.n & 3)) {
        const unsigned long long i = (n4 << 2) + threadIdx.x;
        float a = ld_peer1(p.buf[0] + i);
#pragma unroll
        for (int r = 1; r < W; ++r) a += ld_peer1(p.buf[r] + i);
        a *= p.scale;
#pragma unroll
        for (int r = 0; r < W; ++r) p.buf[r][i] = a;
    }
    // ---- barrier 2: the last block of this rank announces completion and waits for every peer's announcement ----
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned last;
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(p.counter, 1u);
        last = ticket == gridDim.x - 1;
        if (last) *p.counter = 0u;
    }
    __syncthreads();
    if (last && threadIdx.x < W) {
        __threadfence_system();
        st_release_sys(p.flag[threadIdx.x] + W + p.rank, p.epoch);
        while (!epoch_reached(ld_acquire_sys(p.flag[p.rank] + W + threadIdx.x), p.epoch)) {}
    }
}

// The same collective with the reduction done INSIDE the NVSwitch (NVLS): `mc` is the multicast address of the buffers
// (one address that stands for the same offset in every rank's buffer).  multimem.ld_reduce returns the sum over all
// ranks of an element in ONE load (the switch pulls and adds), multimem.st writes the result to every rank in ONE store:
// per GPU and direction the links carry about (W-1)/W n + n/W floats instead of 2 (W-1)/W n.
template <int W>
__global__ void __launch_bounds__(PEER_THREADS, 1) peer_allreduce_nvls_kernel(const PeerParams p, float* mc) {
    if (blockIdx.x == 0 && threadIdx.x < W) st_release_sys(p.flag[threadIdx.x] + p.rank, p.epoch);
    if (threadIdx.x < W)
        while (!epoch_reached(ld_acquire_sys(p.flag[p.rank] + threadIdx.x), p.epoch)) {}
    __syncthreads();
    const unsigned long long n4 = p.n >> 2, per = (n4 + W - 1) / W;
    const unsigned long long lo = per * (unsigned long long)p.rank, hi = lo + per < n4 ? lo + per : n4;
    constexpr int U = 4;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i0 = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned long long i = i0 + u * stride;
            if (i < hi)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                             : "l"(mc + 4 * i)
                             : "memory");
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned long long i = i0 + u * stride;
            if (i < hi)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc + 4 * i),
                             "f"(v[u].x * p.scale), "f"(v[u].y * p.scale), "f"(v[u].z * p.scale), "f"(v[u].w * p.scale)
                             : "memory");
        }
    }
    if (p.rank == W - 1 && blockIdx.x == 0 && threadIdx.x < (unsigned)(p.n & 3)) {
        const unsigned long long i = (n4 << 2) + threadIdx.x;
        float a;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(a) : "l"(mc + i) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc + i), "f"(a * p.scale) : "memory");
    }
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned last;
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(p.counter, 1u);
        last = ticket == gridDim.x - 1;
        if (last) *p.counter = 0u;
    }
    __syncthreads();
    if (last && threadIdx.x < W) {
        __threadfence_system();
        st_release_sys(p.flag[threadIdx.x] + W + p.rank, p.epoch);
        while (!epoch_reached(ld_acquire_sys(p.flag[p.rank] + W + threadIdx.x), p.epoch)) {}
    }
}

unsigned* g_counter[16] = {};

}  // namespace

static int peer_allreduce_impl(phx_ctx* ctx, void* const* bufs, void* multicast, void* const* flags, int rank, int world,
                              size_t n, unsigned epoch, float scale, void* stream) {
    if (!ctx || !flags || world < 2 || world > PEER_MAXW || (world & (world - 1)) || rank < 0 || rank >= world) {
        phx_set_error("peer all-reduce: world size must be 2, 4 or 8 and 0 <= rank < world (got rank %d of %d)", rank,
                      world);
        return PHX_ERR_INVALID;
    }
    PhxDevGuard dev_guard(ctx);
    const int dev = phx_ctx_device(ctx);
    PeerParams p;
    if (multicast && ((uintptr_t)multicast & 15)) {
        phx_set_error("peer all-reduce: multicast address is not 16-byte aligned");
        return PHX_ERR_INVALID;
    }
    for (int r = 0; r < world; ++r) {
        if (!multicast && !bufs) {
            phx_set_error("peer all-reduce: null buffer list");
            return PHX_ERR_INVALID;
        }
        if (multicast) {
            if (!flags[r]) return PHX_ERR_INVALID;
            p.buf[r] = nullptr;
            p.flag[r] = (unsigned*)flags[r];
            continue;
        }
        if (!bufs[r] || !flags[r] || ((uintptr_t)bufs[r] & 15)) {
            phx_set_error("peer all-reduce: buffer %d is null or not 16-byte aligned", r);
            return PHX_ERR_INVALID;
        }
        p.buf[r] = (float*)bufs[r];
        p.flag[r] = (unsigned*)flags[r];
    }
    if (dev < 0 || dev >= 16) return PHX_ERR_INVALID;
    if (!g_counter[dev]) {
        if (cudaMalloc((void**)&g_counter[dev], sizeof(unsigned)) != cudaSuccess ||
            cudaMemset(g_counter[dev], 0, sizeof(unsigned)) != cudaSuccess) {
            phx_set_error("peer all-reduce: cudaMalloc failed");
            return PHX_ERR_CUDA;
        }
    }
    p.rank = rank; p.world = world; p.n = n; p.epoch = epoch; p.scale = scale; p.counter = g_counter[dev];
    // every block must be resident (blocks spin on the peers' flags): one 1024-thread block per SM
    const int grid = phx_ctx_num_sms(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    if (multicast) {
        float* mc = (float*)multicast;
        if (world == 2) peer_allreduce_nvls_kernel<2><<<grid, PEER_THREADS, 0, st>>>(p, mc);
        else if (world == 4) peer_allreduce_nvls_kernel<4><<<grid, PEER_THREADS, 0, st>>>(p, mc);
        else peer_allreduce_nvls_kernel<8><<<grid, PEER_THREADS, 0, st>>>(p, mc);
    } else if (world == 2) peer_allreduce_kernel<2><<<grid, PEER_THREADS, 0, st>>>(p);
    else if (world == 4) peer_allreduce_kernel<4><<<grid, PEER_THREADS, 0, st>>>(p);
    else peer_allreduce_kernel<8><<<grid, PEER_THREADS, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        phx_set_error("peer all-reduce launch: %s", cudaGetErrorString(e));
        return PHX_ERR_CUDA;
    }
    return PHX_OK;
}

extern "C" int phx_peer_allreduce(phx_ctx* ctx, void* const* bufs, void* const* flags, int rank, int world, size_t n,
                                  unsigned epoch, float scale, void* stream) {
    return peer_allreduce_impl(ctx, bufs, nullptr, flags, rank, world, n, epoch, scale, stream);
}
extern "C" int phx_peer_allreduce_nvls(phx_ctx* ctx, void* multicast_buf, void* const* flags, int rank, int world,
                                       size_t n, unsigned epoch, float scale, void* stream) {
    if (!multicast_buf) {
        phx_set_error("peer all-reduce (NVLS): null multicast address");
        return PHX_ERR_INVALID;
    }
    return peer_allreduce_impl(ctx, nullptr, multicast_buf, flags, rank, world, n, epoch, scale, stream);
}
