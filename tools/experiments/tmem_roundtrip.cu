#include <stdio.h>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void tm_ld16(uint32_t taddr, float (&r)[16]) {
    uint32_t u[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]),"=r"(u[1]),"=r"(u[2]),"=r"(u[3]),"=r"(u[4]),"=r"(u[5]),"=r"(u[6]),"=r"(u[7]),
          "=r"(u[8]),"=r"(u[9]),"=r"(u[10]),"=r"(u[11]),"=r"(u[12]),"=r"(u[13]),"=r"(u[14]),"=r"(u[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const float (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        :: "r"(taddr), "r"(__float_as_uint(r[0])),"r"(__float_as_uint(r[1])),"r"(__float_as_uint(r[2])),"r"(__float_as_uint(r[3])),
           "r"(__float_as_uint(r[4])),"r"(__float_as_uint(r[5])),"r"(__float_as_uint(r[6])),"r"(__float_as_uint(r[7])),
           "r"(__float_as_uint(r[8])),"r"(__float_as_uint(r[9])),"r"(__float_as_uint(r[10])),"r"(__float_as_uint(r[11])),
           "r"(__float_as_uint(r[12])),"r"(__float_as_uint(r[13])),"r"(__float_as_uint(r[14])),"r"(__float_as_uint(r[15])) : "memory");
}
__global__ void k(float* out, const float* in) {
    __shared__ uint32_t base_s;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t base = base_s;
    uint32_t taddr = base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 128);
    float r[16];
    for (int i = 0; i < 16; ++i) r[i] = in[threadIdx.x * 16 + i];
    tm_st16(taddr, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    float q[16];
    tm_ld16(taddr, q);
    for (int i = 0; i < 16; ++i) out[threadIdx.x * 16 + i] = q[i];
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(512u) : "memory");
}
int main() {
    float *in, *out; cudaMalloc(&in, 512*16*4); cudaMalloc(&out, 512*16*4);
    float h[512*16]; for (int i = 0; i < 512*16; ++i) h[i] = i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    k<<<1, 512>>>(out, in);
    float g[512*16]; cudaError_t e = cudaMemcpy(g, out, sizeof(g), cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < 512*16; ++i) bad += g[i] != h[i];
    printf("err=%s bad=%d\n", cudaGetErrorString(e), bad);
    return 0;
}
