// Micro-benchmark (diagnostic, not part of the library): issue rate of tcgen05.mma.cta_group::1 on sm_100a for the operand
// forms the branch contraction uses.  One warp per CTA, one CTA per SM, NREP accumulating MMAs back to back, then one
// tcgen05.commit and an mbarrier wait; prints SM cycles per MMA.
//   kind::tf32, M = 128, K = 8, N in {64, 128, 208, 256}, A from shared memory (SS) or tensor memory (TS)
//   kind::f16 (bf16 operands), K = 16, the same N
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// D fp32; A/B format code `fmt` (tf32 = 2 under kind::tf32, bf16 = 1 under kind::f16); K-major; M x N
__device__ __forceinline__ unsigned idesc(int fmt, int M, int N) {
    return (1u << 4) | ((unsigned)fmt << 7) | ((unsigned)fmt << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

template <int KIND, int TS>   // KIND 0: tf32, 1: f16(bf16)
__global__ void __launch_bounds__(32, 1) rate_kernel(int N, int nrep, int nbuf, int nacc, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long bar;
    __shared__ unsigned slot;
    const unsigned bar_a = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 32) reinterpret_cast<float*>(smem)[i] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = slot;
    // operand tiles: A 128 rows x 32 bytes (two core-matrix columns, LBO = 16*128), B N rows x 32 bytes (LBO = N/8*128);
    // `nbuf` different tiles are cycled through (a ring, like the real kernel)
    const unsigned a_bytes = 128 * 32, b_bytes = (unsigned)N * 32;
    const unsigned a0 = smem_u32(smem), b0 = a0 + 8 * a_bytes;
    const unsigned id = idesc(KIND == 0 ? 2 : 1, 128, N);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        // descriptors of the 4 ring tiles precomputed; the loop body is 4 (x NACC accumulators) straight-line MMAs
        uint64_t ad[4], bd[4];
        unsigned ta[4];
        for (int s = 0; s < 4; ++s) {
            const int t = s % nbuf;
            ad[s] = smem_desc(a0 + t * a_bytes, 16 * 128, 128);
            bd[s] = smem_desc(b0 + t * b_bytes, (unsigned)(N / 8) * 128, 128);
            ta[s] = tmem + 2 * 256 - 64 + 8 * t;
        }
        const unsigned d1 = tmem + (nacc > 1 ? 256 : 0);   // second accumulator (independent chain) when nacc == 2
        t0 = clock64();
        for (int i = 0; i < nrep; i += 4) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const unsigned acc = (i + s) > 1;
                const unsigned d = (s & 1) ? d1 : tmem;
                if (TS) {
                    if (KIND == 0)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
                                     "r"(ta[s]), "l"(bd[s]), "r"(id), "r"(acc));
                    else
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
                                     "r"(ta[s]), "l"(bd[s]), "r"(id), "r"(acc));
                } else {
                    if (KIND == 0)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                                     "l"(ad[s]), "l"(bd[s]), "r"(id), "r"(acc));
                    else
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                                     "l"(ad[s]), "l"(bd[s]), "r"(id), "r"(acc));
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
        unsigned done = 0;
        while (!done)
            asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\nselp.u32 %0, 1, 0, P1;\n}\n"
                         : "=r"(done)
                         : "r"(bar_a)
                         : "memory");
        t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int KIND, int TS>
void run(const char* name, long long* dout) {
    cudaFuncSetAttribute(rate_kernel<KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int Ns[] = {64, 128, 208, 256};
    for (int nacc : {1, 2})
    for (int nbuf : {1, 4}) {
        for (int N : Ns) {
            if (nacc == 2 && N > 192) continue;   // two accumulators + the A columns must fit 512 columns
            const int nrep = 2000;
            rate_kernel<KIND, TS><<<148, 32, 200 * 1024>>>(N, nrep, nbuf, nacc, dout);
            rate_kernel<KIND, TS><<<148, 32, 200 * 1024>>>(N, nrep, nbuf, nacc, dout);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("%s N=%d: %s\n", name, N, cudaGetErrorString(e));
                return;
            }
            long long h[148];
            cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0;
            for (int i = 0; i < 148; ++i) s += (double)h[i];
            printf("%-22s nacc=%d nbuf=%d N=%3d: %7.1f cycles / MMA  (floor N/2 = %d)\n", name, nacc, nbuf, N, s / 148 / nrep, N / 2);
        }
    }
}

int main() {
    long long* dout;
    cudaMalloc(&dout, 148 * sizeof(long long));
    run<0, 0>("tf32 K=8  A smem (SS)", dout);
    run<0, 1>("tf32 K=8  A tmem (TS)", dout);
    run<1, 0>("bf16 K=16 A smem (SS)", dout);
    run<1, 1>("bf16 K=16 A tmem (TS)", dout);
    return 0;
}
