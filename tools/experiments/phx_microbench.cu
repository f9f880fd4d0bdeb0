// Micro-benchmarks of the building blocks of the resident kernels (diagnostics only; tools/microbench.py).
#include "../../phoenix_b200/csrc/phx_resident.cuh"

namespace {

// iters x all-reduce of an n-float vector over the grid, nothing else
__global__ void __launch_bounds__(PHX_THREADS, 1) mb_allreduce_kernel(const __grid_constant__ ResParams p, int n,
                                                                      int iters, float* out) {
    Smem s(p);
    s.x.ep = __ldcg(p.ll.epoch);
    s.x.ny = s.x.nd = 0;
    float* vec = s.sp();
    for (int i = threadIdx.x; i < n; i += THREADS) vec[i] = 1.0f + blockIdx.x;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        grid_allreduce_f(p, s, vec, n);
        for (int i = threadIdx.x; i < n; i += THREADS) vec[i] = vec[i] * (1.0f / gridDim.x);
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x < n) out[threadIdx.x] = vec[threadIdx.x];
    epilogue_epoch(p, s);
}

// iters x scalar (double) grid sums
__global__ void __launch_bounds__(PHX_THREADS, 1) mb_sumd_kernel(const __grid_constant__ ResParams p, int nd, int iters,
                                                                 float* out) {
    Smem s(p);
    s.x.ep = __ldcg(p.ll.epoch);
    s.x.ny = s.x.nd = 0;
    double* v = s.ctrl()->dsum;
    for (int it = 0; it < iters; ++it) {
        if (threadIdx.x < nd) v[threadIdx.x] = 1.0;
        __syncthreads();
        grid_sum_d(p, s, v, nd);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (float)v[0];
    epilogue_epoch(p, s);
}

// iters x one streamed pass over this CTA's rows of W1 (column accumulation only)
template <int NV>
__global__ void __launch_bounds__(PHX_THREADS, 1) mb_pass_kernel(const __grid_constant__ ResParams p, int which,
                                                                 int iters, float* out) {
    Smem s(p);
    s.rg.par = 0;
    s.rg.pre = -1;
    ring_init(p, s);
    resident_wait(p, s);
    s.tmem = tmem_setup<NV>(p, s);
    for (int i = threadIdx.x; i < p.B * p.gpc; i += THREADS) { s.acts()[i] = 1.f; s.actl()[i] = 0.5f; }
    for (int i = threadIdx.x; i < p.B * p.K2; i += THREADS) s.sp()[i] = 1.f;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        if (which == MAT_W1) passA<NV, 1>(p, s, s.acts(), s.actl(), s.sp(), MAT_W1);
        else fwd_passBA<NV, 1>(p, s, which == 3, [](int, int, int, float f, bool) { return f; });
    }
    ring_drain(p, s);
    tmem_release(p, s.tmem);
    if (threadIdx.x < p.K2) out[(size_t)blockIdx.x * p.K2 + threadIdx.x] = s.sp()[threadIdx.x];
}

// raw hop latency: CTA 0 and CTA `peer` bounce one tagged slot back and forth
__global__ void mb_pingpong_kernel(unsigned long long* slots, int peer, int iters, long long* out) {
    if (threadIdx.x != 0) return;
    if (blockIdx.x != 0 && blockIdx.x != peer) return;
    long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        if (blockIdx.x == 0) {
            ll_put(slots, 1.f, (unsigned)it);
            ll_get(slots + 64, (unsigned)it);
        } else {
            ll_get(slots, (unsigned)it);
            ll_put(slots + 64, 1.f, (unsigned)it);
        }
    }
    if (blockIdx.x == 0) out[0] = clock64() - t0;
}

// one-hop all-to-all flag barrier: every CTA posts one slot, warp 0 of every CTA polls all of them
__global__ void mb_barrier_kernel(unsigned long long* slots, int iters, long long* out) {
    const int nC = gridDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        unsigned long long* base = slots + (size_t)(it & 1) * 4096;
        if (threadIdx.x == 0) ll_put(base + blockIdx.x * 16, 1.f, (unsigned)it);   // one 128-byte line per CTA
        if (warp == 0)
            for (int c = lane; c < nC; c += 32) ll_get(base + c * 16, (unsigned)it);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

// same, but all flags packed densely (8 bytes apart)
__global__ void mb_barrier_dense_kernel(unsigned long long* slots, int iters, long long* out) {
    const int nC = gridDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        unsigned long long* base = slots + (size_t)(it & 1) * 4096;
        if (threadIdx.x == 0) ll_put(base + blockIdx.x, 1.f, (unsigned)it);
        if (warp == 0)
            for (int c = lane; c < nC; c += 32) ll_get(base + c, (unsigned)it);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

// instrumented copy of the two-phase all-reduce: per-iteration clock stamps of CTA `blockIdx.x` into out[cta][8]
//   0 puts issued | 1 reducer: all partials seen | 2 reducer: result posted | 3 results seen (thread 0) | 4 after barrier
template <int VARIANT>
__global__ void __launch_bounds__(PHX_THREADS, 1) mb_xchg_kernel(const __grid_constant__ ResParams p, int n, int iters,
                                                                 long long* out) {
    Smem s(p);
    s.x.ep = __ldcg(p.ll.epoch);
    s.x.ny = s.x.nd = 0;
    float* vec = s.sp();
    const int nC = gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long acc[5] = {0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < n; i += THREADS) vec[i] = 1.0f;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        const unsigned tag = s.x.next_tag();
        long long t0 = clock64();
        const int nq = n >> 2;
        if (VARIANT < 4) {
            unsigned long long* mine = p.ll.xpart + (size_t)blockIdx.x * PHX_LL_NMAX;
            for (int i = 2 * threadIdx.x; i < n; i += 2 * THREADS) ll_put2(mine + i, vec[i], vec[i + 1], tag);
        } else {   // transposed: [quad][CTA][4] so that the reducer's gather is contiguous
            for (int i = threadIdx.x; i < 2 * nq; i += THREADS) {
                const int q = i >> 1, h = i & 1;
                ll_put2(p.ll.xpart + ((size_t)q * PHX_LL_MAXC + blockIdx.x) * 4 + 2 * h, vec[4 * q + 2 * h],
                        vec[4 * q + 2 * h + 1], tag);
            }
        }
        long long t1 = clock64(), t2 = t1, t3 = t1;
        for (int q = blockIdx.x + nC * warp; q < nq; q += nC * WARPS) {
            unsigned long long w[5][4];
            const unsigned long long* src0 = (VARIANT < 4) ? p.ll.xpart + (size_t)lane * PHX_LL_NMAX + 4 * q
                                                            : p.ll.xpart + ((size_t)q * PHX_LL_MAXC + lane) * 4;
            const size_t ustride = (VARIANT < 4) ? (size_t)32 * PHX_LL_NMAX : (size_t)32 * 4;
            unsigned need = 0;
#pragma unroll
            for (int u = 0; u < 5; ++u)
                if (lane + 32 * u < nC) need |= 3u << (2 * u);
            if (VARIANT == 2 || VARIANT == 3) __nanosleep(VARIANT == 2 ? 200 : 400);
            while (need) {
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    if (need & (1u << (2 * u))) ll_ld2(src0 + u * ustride, w[u][0], w[u][1]);
                    if (need & (2u << (2 * u))) ll_ld2(src0 + u * ustride + 2, w[u][2], w[u][3]);
                }
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    if ((need & (1u << (2 * u))) && (unsigned)(w[u][0] >> 32) == tag && (unsigned)(w[u][1] >> 32) == tag)
                        need &= ~(1u << (2 * u));
                    if ((need & (2u << (2 * u))) && (unsigned)(w[u][2] >> 32) == tag && (unsigned)(w[u][3] >> 32) == tag)
                        need &= ~(2u << (2 * u));
                }
            }
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int u = 0; u < 5; ++u) {
                if (lane + 32 * u < nC) {
                    a0 += __uint_as_float((unsigned)w[u][0]);
                    a1 += __uint_as_float((unsigned)w[u][1]);
                    a2 += __uint_as_float((unsigned)w[u][2]);
                    a3 += __uint_as_float((unsigned)w[u][3]);
                }
            }
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
            t2 = clock64();
            const int e = lane & 3;
            if (VARIANT == 0) {
                ll_put(p.ll.xres + (size_t)(lane >> 2) * PHX_LL_NMAX + 4 * q + e,
                       e == 0 ? a0 : (e == 1 ? a1 : (e == 2 ? a2 : a3)), tag);
            } else {   // 16-byte posts: lanes 0..15 cover 8 replicas x 2 halves
                if (lane < 16)
                    ll_put2(p.ll.xres + (size_t)(lane >> 1) * PHX_LL_NMAX + 4 * q + 2 * (lane & 1),
                            (lane & 1) ? a2 : a0, (lane & 1) ? a3 : a1, tag);
            }
            t3 = clock64();
        }
        const unsigned long long* res = p.ll.xres + (size_t)(blockIdx.x % PHX_LL_RCOPIES) * PHX_LL_NMAX;
        if (VARIANT == 2 || VARIANT == 3) __nanosleep(VARIANT == 2 ? 400 : 800);
        for (int i = 2 * threadIdx.x; i < n; i += 2 * THREADS) ll_get2(res + i, tag, vec[i], vec[i + 1]);
        long long t4 = clock64();
        __syncthreads();
        long long t5 = clock64();
        if (threadIdx.x == 0) {
            acc[0] += t1 - t0; acc[1] += t2 - t0; acc[2] += t3 - t0; acc[3] += t4 - t0; acc[4] += t5 - t0;
        }
        for (int i = threadIdx.x; i < n; i += THREADS) vec[i] = 1.0f;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < 5; ++k) out[blockIdx.x * 8 + k] = acc[k];
    epilogue_epoch(p, s);
}

}  // namespace

extern "C" int phx_microbench_xchg(phx_ctx* ctx, int G, int H, int variant, int n, int iters, void* workspace,
                                   long long* out, void* stream) {
    ResLaunchPlan plan;
    int rc = phx_resident_plan(phx_ctx_num_sms(ctx), G, H, 1, 0, &plan);
    if (rc != PHX_OK) return rc;
    ResParams p;
    memset(&p, 0, sizeof(p));
    p.G = G; p.H = H; p.Hp = phx_Hp(H); p.K2 = 2 * p.Hp; p.K2q = p.K2 / 4; p.B = 1; p.gpc = plan.gpc;
    p.ring_rows = plan.ring_rows; p.ring_stages = plan.ring_stages; p.so = plan.so;
    p.ll = phx_ll_view(workspace);
    void* args[] = {&p, &n, &iters, &out};
    const void* fn = variant == 0 ? (const void*)mb_xchg_kernel<0>
                     : (variant == 1 ? (const void*)mb_xchg_kernel<1>
                                     : (variant == 2 ? (const void*)mb_xchg_kernel<2>
                                                     : (variant == 3 ? (const void*)mb_xchg_kernel<3>
                                                                     : (const void*)mb_xchg_kernel<4>)));
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(plan.nCTA), dim3(PHX_THREADS), args, plan.smem_bytes,
                                                (cudaStream_t)stream);
    if (e != cudaSuccess) {
        phx_set_error("microbench launch: %s", cudaGetErrorString(e));
        return PHX_ERR_CUDA;
    }
    return plan.nCTA;
}

extern "C" int phx_microbench_sync(int what, int nCTA, int peer, int iters, void* slots, long long* out, void* stream) {
    if (what == 0) mb_pingpong_kernel<<<nCTA, 32, 0, (cudaStream_t)stream>>>((unsigned long long*)slots, peer, iters, out);
    else {
        void* args[] = {&slots, &iters, &out};
        cudaLaunchCooperativeKernel(what == 1 ? (const void*)mb_barrier_kernel : (const void*)mb_barrier_dense_kernel,
                                    dim3(nCTA), dim3(128), args, 0, (cudaStream_t)stream);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int phx_microbench(phx_ctx* ctx, int G, int H, int B, int what, int n, int iters, const float* packed,
                              void* workspace, size_t workspace_bytes, float* out, void* stream) {
    ResLaunchPlan plan;
    int rc = phx_resident_plan(phx_ctx_num_sms(ctx), G, H, B, 0, &plan);
    if (rc != PHX_OK) return rc;
    ResParams p;
    memset(&p, 0, sizeof(p));
    p.G = G; p.H = H; p.Hp = phx_Hp(H); p.K2 = 2 * p.Hp; p.K2q = p.K2 / 4; p.B = B; p.gpc = plan.gpc;
    p.ring_rows = plan.ring_rows; p.ring_stages = plan.ring_stages; p.so = plan.so;
    p.w = phx_packed_view(packed, G, H);
    p.ll = phx_ll_view(workspace);
    void* args3[] = {&p, &n, &iters, &out};
    int which = (what == 2) ? MAT_W1 : (what == 3 ? MAT_WA : 3);
    void* args2[] = {&p, &which, &iters, &out};
    const void* fn;
    void** args;
    if (what == 0) { fn = (const void*)mb_allreduce_kernel; args = args3; }
    else if (what == 1) { fn = (const void*)mb_sumd_kernel; args = args3; }
    else {
        fn = plan.NV == 1 ? (const void*)mb_pass_kernel<1> : (plan.NV == 2 ? (const void*)mb_pass_kernel<2>
                                                                             : (const void*)mb_pass_kernel<4>);
        args = args2;
    }
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(plan.nCTA), dim3(PHX_THREADS), args, plan.smem_bytes,
                                                (cudaStream_t)stream);
    if (e != cudaSuccess) {
        phx_set_error("microbench launch: %s", cudaGetErrorString(e));
        return PHX_ERR_CUDA;
    }
    return plan.nCTA;
}
