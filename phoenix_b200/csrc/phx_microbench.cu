// Micro-benchmarks of the building blocks of the resident kernels (diagnostics only; tools/microbench.py).
#include "phx_resident.cuh"

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
__global__ void __launch_bounds__(PHX_THREADS, 1) mb_pass_kernel(const __grid_constant__ ResParams p, int iters,
                                                                 float* out) {
    Smem s(p);
    s.rg.par = 0;
    s.rg.pre_mat = nullptr;
    ring_init(p, s);
    for (int i = threadIdx.x; i < p.B * p.gpc; i += THREADS) { s.acts()[i] = 1.f; s.actl()[i] = 0.5f; }
    __syncthreads();
    for (int it = 0; it < iters; ++it) passA<NV, 1>(p, s);
    if (threadIdx.x < p.K2) out[(size_t)blockIdx.x * p.K2 + threadIdx.x] = s.sp()[threadIdx.x];
}

}  // namespace

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
    void* args2[] = {&p, &iters, &out};
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
