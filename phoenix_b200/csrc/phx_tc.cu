// Tensor-core path of the batched RHS (ODENet.forward / prior_only_forward, odenet.py:85-98) for B >= 128 rows:
// hand-written tcgen05.mma (kind::tf32, 3xTF32 split for fp32 parity) with accumulators in tensor memory, operands
// staged by 1-D TMA bulk copies and -- for the Hill activations -- by producer warps that compute them on the fly.
// Layouts: phx_tc.cuh.  Three launches per RHS evaluation:
//
//   tc_branch_kernel   [S|P]partial = act(y) W1        M = 128 batch rows per CTA, N = 2 x Hn (both branches, so each
//                      y element's soft-sign / log1p is computed exactly once), K = a slice of the genes (K-split so
//                      that ~148 CTAs are busy).  Warps 0-7 load y, evaluate s(y), l(y), split hi/lo and write the four
//                      A tiles of a stage; warp 8 streams the matching w1img chunk with cp.async.bulk; warp 9 issues
//                      the MMAs; warps 0-7 then read the accumulators back (tcgen05.ld) and store the partial sums.
//   tc_spfinish_kernel sums the K-split partials in fixed order, adds the bias, exponentiates the prods half, writes the
//                      plain [B][K2] copy (VJP / parity) and the hi|lo operand image of the next contraction.
//   tc_joint_kernel    f^T tile = WA [S|P]^T           persistent, M = 128 genes, N = 256 batch rows, K = 2*Hn, double-
//                      buffered TMEM accumulators; both operands arrive by bulk copy (warp 0), warp 1 issues MMAs,
//                      warps 2-5 run the epilogue f = fscale * relu(m) * (J - y) straight out of tensor memory: lanes
//                      are genes, so every global access of the epilogue is a coalesced 128-byte row segment.
//
// Every accumulation order is fixed (MMA issue order, K-split reduction order) => results are run-to-run identical.
#include <stdint.h>
#include <stdlib.h>
#include "phx_common.cuh"

namespace {

constexpr int BK = PHX_TC_BK;

__device__ __forceinline__ unsigned smem_u32(const void* ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor bit layout; version 1 = sm_100)
__device__ __forceinline__ uint64_t smem_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D fp32, A/B tf32, both K-major, M x N  (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ inline unsigned idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(unsigned tmem_d, uint64_t adesc, uint64_t bdesc, unsigned idesc,
                                         unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc512(unsigned slot_saddr) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_saddr), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free512(unsigned base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(u[i]);
}

__device__ __forceinline__ float tf32_rna(float x) {
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = tf32_rna(x);
    lo = tf32_rna(x - hi);
}

// ---- operand images of the weights -------------------------------------------------------------------------------------
__global__ void tc_pack_kernel(int G, int H, int Hp, int K2, int Hn, const float* __restrict__ W1,
                               const float* __restrict__ WA, float* __restrict__ w1img, float* __restrict__ waimg) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // w1img: one thread per (gene g, branch, hidden unit n): reads of W1 rows are coalesced over n
    const int KB1 = phx_tc_KB1(G);
    const size_t n1 = (size_t)KB1 * BK * 2 * Hn;
    for (size_t i = tid0; i < n1; i += stride) {
        const int n = (int)(i % Hn);
        const int br = (int)((i / Hn) & 1);
        const int g = (int)(i / (2 * (size_t)Hn));
        const float v = (g < G && n < H) ? W1[(size_t)g * K2 + br * Hp + n] : 0.f;
        float hi, lo;
        split_tf32(v, hi, lo);
        const int kb = g / BK, k = g % BK;
        float* chunk = w1img + (size_t)kb * 4 * Hn * BK + (size_t)br * 2 * Hn * BK;
        const int off = phx_tc_tile_off(Hn, n, k);
        chunk[off] = hi;
        chunk[Hn * BK + off] = lo;
    }
    // waimg: one thread per (gene g, column c of [S|P] in Hn-padded numbering): reads of WA rows coalesced over c
    const int KB2 = 2 * Hn / BK, GT = phx_tc_GT(G);
    const size_t n2 = (size_t)GT * 128 * 2 * Hn;
    for (size_t i = tid0; i < n2; i += stride) {
        const int c = (int)(i % (2 * Hn));
        const int g = (int)(i / (2 * (size_t)Hn));
        const int br = c / Hn, n = c % Hn;
        const float v = (g < G && n < H) ? WA[(size_t)g * K2 + br * Hp + n] : 0.f;
        float hi, lo;
        split_tf32(v, hi, lo);
        const int gt = g >> 7, r = g & 127, kb = c / BK, k = c % BK;
        float* chunk = waimg + ((size_t)gt * KB2 + kb) * 2 * 128 * BK;
        const int off = phx_tc_tile_off(128, r, k);
        chunk[off] = hi;
        chunk[128 * BK + off] = lo;
    }
}

__device__ __forceinline__ void hill(float y, float& s, float& l) {
    float z = y - 0.5f;
    float den = 1.0f + fabsf(z);
    s = z / den;
    l = log1pf(s);
}

// ---- kernel 1: branch contraction ----------------------------------------------------------------------------------------
struct BranchParams {
    int G, B, Bpad, Hn, KB1, kb_per_split, stages, nterms;
    unsigned a_lbo, a_sbo, b_lbo, b_sbo;   // descriptor fields
    unsigned a_kadv, b_kadv;               // byte advance of the start address per K = 8 MMA (two core matrices)
    const float* y;       // [B][G]
    const float* w1img;
    float* spart;         // [ks][Bpad][2*Hn]
};
constexpr int K1_THREADS = 320;
constexpr unsigned K1_A_TILE = 128 * BK * 4;   // bytes of one 128 x 16 tile
constexpr unsigned K1_A_BYTES = 4 * K1_A_TILE; // s_hi, s_lo, l_hi, l_lo

__global__ void __launch_bounds__(K1_THREADS, 1) tc_branch_kernel(BranchParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * 128, ks = blockIdx.y;
    const int kb0 = ks * p.kb_per_split;
    const int nkb = min(p.KB1, kb0 + p.kb_per_split) - kb0;   // >= 1 by construction (phx_tc_ksplit)
    const int S = p.stages, Hn = p.Hn;
    const unsigned b_tile = (unsigned)Hn * BK * 4;
    const unsigned b_bytes = 4 * b_tile;
    const unsigned stage_bytes = K1_A_BYTES + b_bytes;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)S * stage_bytes);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + S), done = smem_u32(bars + 2 * S);
    unsigned* slot = reinterpret_cast<unsigned*>(bars + 2 * S + 1);
    const unsigned stage0 = smem_u32(smem);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 9);    // 8 producer warps + the bulk-copy issuer (expect_tx)
            mbar_init(empty0 + 8 * s, 1);   // tcgen05.commit
        }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc512(smem_u32(slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *slot;

    if (warp < 8) {
        // ---- producers: Hill activations of this CTA's 128 x 16 slab of y, hi/lo split, core-matrix layout ----
        const int rl = warp * 16 + (lane & 7), kc = lane >> 3;   // rows rl and rl + 8, k-chunk kc (4 floats)
        float cur[8], nxt[8];
        auto load = [&](int kb, float (&v)[8]) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int row = m0 + rl + 8 * it;
                const float* src = p.y + (size_t)row * p.G;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int g = kb * BK + kc * 4 + j;
                    v[it * 4 + j] = (row < p.B && g < p.G) ? __ldg(src + g) : 0.5f;   // s(0.5) = l(0.5) = 0
                }
            }
        };
        load(kb0, cur);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % S;
            const unsigned ph = (unsigned)(i / S) & 1u;
            if (i + 1 < nkb) load(kb0 + i + 1, nxt);
            float4 t[2][4];
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                float sv[4], lv[4], shi[4], slo[4], lhi[4], llo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    hill(cur[it * 4 + j], sv[j], lv[j]);
                    split_tf32(sv[j], shi[j], slo[j]);
                    split_tf32(lv[j], lhi[j], llo[j]);
                }
                t[it][0] = make_float4(shi[0], shi[1], shi[2], shi[3]);
                t[it][1] = make_float4(slo[0], slo[1], slo[2], slo[3]);
                t[it][2] = make_float4(lhi[0], lhi[1], lhi[2], lhi[3]);
                t[it][3] = make_float4(llo[0], llo[1], llo[2], llo[3]);
            }
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            unsigned char* a = smem + (size_t)s * stage_bytes;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int r = rl + 8 * it;
                const unsigned off = (unsigned)(((kc * 16 + (r >> 3)) * 8 + (r & 7)) * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(a + q * K1_A_TILE + off) = t[it][q];
            }
            fence_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
#pragma unroll
            for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
        }
        // ---- epilogue: accumulators -> K-split partial sums ----
        mbar_wait(done, 0u);
        tc_fence_after();
        const int q = warp & 3, br = warp >> 2;
        const int row = m0 + q * 32 + lane;
        float* dst = p.spart + ((size_t)ks * p.Bpad + row) * (2 * Hn) + br * Hn;
        for (int c0 = 0; c0 < Hn; c0 += 16) {
            float v[16];
            tmem_ld16(tmem + ((unsigned)(q * 32) << 16) + (unsigned)(br * 256 + c0), v);
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    } else if (warp == 8) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % S;
                const unsigned ph = (unsigned)(i / S) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                mbar_expect_tx(full0 + 8 * s, b_bytes);
                bulk_g2s(stage0 + s * stage_bytes + K1_A_BYTES, p.w1img + (size_t)(kb0 + i) * 4 * Hn * BK, b_bytes,
                         full0 + 8 * s);
            }
        }
    } else {
        if (lane == 0) {
            const unsigned idesc = idesc_tf32(128, Hn);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % S;
                const unsigned ph = (unsigned)(i / S) & 1u;
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const unsigned a_base = stage0 + s * stage_bytes, b_base = a_base + K1_A_BYTES;
#pragma unroll
                for (int k8 = 0; k8 < BK / 8; ++k8) {
#pragma unroll
                    for (int br = 0; br < 2; ++br) {
                        const uint64_t a_hi = smem_desc(a_base + (2 * br) * K1_A_TILE + k8 * p.a_kadv, p.a_lbo, p.a_sbo);
                        const uint64_t a_lo =
                            smem_desc(a_base + (2 * br + 1) * K1_A_TILE + k8 * p.a_kadv, p.a_lbo, p.a_sbo);
                        const uint64_t b_hi = smem_desc(b_base + (2 * br) * b_tile + k8 * p.b_kadv, p.b_lbo, p.b_sbo);
                        const uint64_t b_lo =
                            smem_desc(b_base + (2 * br + 1) * b_tile + k8 * p.b_kadv, p.b_lbo, p.b_sbo);
                        const unsigned d = tmem + (unsigned)(br * 256);
                        const unsigned acc = (i > 0 || k8 > 0) ? 1u : 0u;
                        if (p.nterms == 3) {
                            mma_tf32(d, a_lo, b_hi, idesc, acc);
                            mma_tf32(d, a_hi, b_lo, idesc, 1u);
                            mma_tf32(d, a_hi, b_hi, idesc, 1u);
                        } else {
                            mma_tf32(d, a_hi, b_hi, idesc, acc);
                        }
                    }
                }
                mma_commit(empty0 + 8 * s);
            }
            mma_commit(done);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_free512(tmem);
}

// ---- K-split reduction + bias + exp + operand image of [S|P] ------------------------------------------------------------
// one thread per (batch row b in [0, BT*256), 4 consecutive Hn-numbered columns)
__global__ void tc_spfinish_kernel(int B, int Bpad, int H, int Hp, int Hn, int K2, int ks, const float* __restrict__ spart,
                                   const float* __restrict__ bias, float* __restrict__ SP, float* __restrict__ spimg) {
    const int BT = phx_tc_BT(B), KB2 = 2 * Hn / BK, C4 = 2 * Hn / 4;
    const size_t total = (size_t)BT * 256 * C4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i % ((size_t)BT * 256));   // rows fastest: image stores of a warp are contiguous
        const int c4 = (int)(i / ((size_t)BT * 256));
        const int c = c4 * 4, br = c / Hn, n0 = c % Hn;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (b < B) {
            for (int s = 0; s < ks; ++s) {
                const float4 t = *reinterpret_cast<const float4*>(spart + ((size_t)s * Bpad + b) * (2 * Hn) + c);
                v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + j;
                if (n < H) {
                    float x = v[j] + bias[br * Hp + n];
                    v[j] = br ? expf(x) : x;
                } else {
                    v[j] = 0.f;
                }
                if (n < Hp) SP[(size_t)b * K2 + br * Hp + n] = v[j];
            }
        }
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_tf32(v[j], hi[j], lo[j]);
        const int bt = b >> 8, r = b & 255, kb = c / BK, k = c % BK;
        float* chunk = spimg + ((size_t)bt * KB2 + kb) * 2 * 256 * BK;
        const int off = phx_tc_tile_off(256, r, k);
        *reinterpret_cast<float4*>(chunk + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(chunk + 256 * BK + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// ---- kernel 2: joint contraction + decay epilogue -------------------------------------------------------------------------
struct JointParams {
    int G, B, KB2, GT, BT, nterms, decay;
    unsigned a_lbo, a_sbo, b_lbo, b_sbo;   // descriptor fields
    unsigned a_kadv, b_kadv;               // byte advance of the start address per K = 8 MMA (two core matrices)
    float fscale;
    const float* waimg;
    const float* spimg;
    const float* y;       // [B][G]
    const float* relum;   // [G]
    float* f;             // [B][G]
};
constexpr int K2_THREADS = 192;
constexpr int K2_STAGES = 4;
constexpr unsigned K2_A_TILE = 128 * BK * 4, K2_B_TILE = 256 * BK * 4;
constexpr unsigned K2_STAGE_BYTES = 2 * K2_A_TILE + 2 * K2_B_TILE;   // 48 KB

__global__ void __launch_bounds__(K2_THREADS, 1) tc_joint_kernel(JointParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int S = K2_STAGES;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)S * K2_STAGE_BYTES);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + S), tfull0 = smem_u32(bars + 2 * S),
                   tempty0 = smem_u32(bars + 2 * S + 2);
    unsigned* slot = reinterpret_cast<unsigned*>(bars + 2 * S + 4);
    const unsigned stage0 = smem_u32(smem);
    const int ntiles = p.GT * p.BT;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull0 + 8 * b, 1);
            mbar_init(tempty0 + 8 * b, 4);   // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc512(smem_u32(slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int gt = t / p.BT, bt = t % p.BT;
                for (int kb = 0; kb < p.KB2; ++kb, ++it) {
                    const int s = it % S;
                    const unsigned ph = (unsigned)(it / S) & 1u;
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);
                    mbar_expect_tx(full0 + 8 * s, K2_STAGE_BYTES);
                    const unsigned dst = stage0 + s * K2_STAGE_BYTES;
                    bulk_g2s(dst, p.waimg + ((size_t)gt * p.KB2 + kb) * 2 * 128 * BK, 2 * K2_A_TILE, full0 + 8 * s);
                    bulk_g2s(dst + 2 * K2_A_TILE, p.spimg + ((size_t)bt * p.KB2 + kb) * 2 * 256 * BK, 2 * K2_B_TILE,
                             full0 + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const unsigned idesc = idesc_tf32(128, 256);
            int it = 0, j = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
                const int buf = j & 1;
                const unsigned tph = (unsigned)(j >> 1) & 1u;
                mbar_wait(tempty0 + 8 * buf, tph ^ 1u);   // epilogue has drained this accumulator
                tc_fence_after();
                const unsigned d = tmem + (unsigned)(buf * 256);
                for (int kb = 0; kb < p.KB2; ++kb, ++it) {
                    const int s = it % S;
                    const unsigned ph = (unsigned)(it / S) & 1u;
                    mbar_wait(full0 + 8 * s, ph);
                    tc_fence_after();
                    const unsigned a_base = stage0 + s * K2_STAGE_BYTES, b_base = a_base + 2 * K2_A_TILE;
#pragma unroll
                    for (int k8 = 0; k8 < BK / 8; ++k8) {
                        const uint64_t a_hi = smem_desc(a_base + k8 * p.a_kadv, p.a_lbo, p.a_sbo);
                        const uint64_t a_lo = smem_desc(a_base + K2_A_TILE + k8 * p.a_kadv, p.a_lbo, p.a_sbo);
                        const uint64_t b_hi = smem_desc(b_base + k8 * p.b_kadv, p.b_lbo, p.b_sbo);
                        const uint64_t b_lo = smem_desc(b_base + K2_B_TILE + k8 * p.b_kadv, p.b_lbo, p.b_sbo);
                        const unsigned acc = (kb > 0 || k8 > 0) ? 1u : 0u;
                        if (p.nterms == 3) {
                            mma_tf32(d, a_lo, b_hi, idesc, acc);
                            mma_tf32(d, a_hi, b_lo, idesc, 1u);
                            mma_tf32(d, a_hi, b_hi, idesc, 1u);
                        } else {
                            mma_tf32(d, a_hi, b_hi, idesc, acc);
                        }
                    }
                    mma_commit(empty0 + 8 * s);
                }
                mma_commit(tfull0 + 8 * buf);
            }
        }
    } else {
        // ---- epilogue warps 2..5: lane quarter q = warp % 4 of the accumulator = 32 consecutive genes ----
        const int q = warp & 3;
        int j = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
            const int gt = t / p.BT, bt = t % p.BT;
            const int buf = j & 1;
            const unsigned tph = (unsigned)(j >> 1) & 1u;
            const int g = gt * 128 + q * 32 + lane;
            const bool gok = g < p.G;
            const float rm = (gok && p.decay) ? p.relum[g] : 1.f;
            mbar_wait(tfull0 + 8 * buf, tph);
            tc_fence_after();
            const int b0 = bt * 256;
            for (int c0 = 0; c0 < 256 && b0 + c0 < p.B; c0 += 16) {
                float v[16], yv[16];
                tmem_ld16(tmem + ((unsigned)(q * 32) << 16) + (unsigned)(buf * 256 + c0), v);
                if (p.decay) {
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) {
                        const int row = b0 + c0 + jj;
                        yv[jj] = (gok && row < p.B) ? p.y[(size_t)row * p.G + g] : 0.f;
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const int row = b0 + c0 + jj;
                    if (gok && row < p.B) {
                        const float r = p.decay ? rm * (v[jj] - yv[jj]) : v[jj];
                        p.f[(size_t)row * p.G + g] = p.fscale * r;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_free512(tmem);
}

int debug_flags() {
    static int flags = -1;
    if (flags < 0) {
        const char* e = getenv("PHX_TC_DEBUG");
        flags = e ? atoi(e) : 0;
    }
    return flags;
}

}  // namespace

int phx_tc_pack_launch(int G, int H, const PhxPacked& w, cudaStream_t st) {
    const int Hp = phx_Hp(H), Hn = phx_tc_Hn(H);
    tc_pack_kernel<<<PHX_TC_SMS * 8, 256, 0, st>>>(G, H, Hp, 2 * Hp, Hn, reinterpret_cast<const float*>(w.W1),
                                                    reinterpret_cast<const float*>(w.WA), const_cast<float*>(w.w1img),
                                                    const_cast<float*>(w.waimg));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        phx_set_error("tc_pack launch: %s", cudaGetErrorString(e));
        return PHX_ERR_CUDA;
    }
    return PHX_OK;
}

// SP (plain [B][K2], bias/exp applied) and f; tcws = scratch of phx_tc_scratch_floats(G, H, B) floats
int phx_tc_rhs_forward_launch(int G, int H, int B, const PhxPacked& w, const float* y, float* f, int decay,
                              float fscale, float* SP, float* tcws, cudaStream_t st) {
    const int Hp = phx_Hp(H), K2 = 2 * Hp, Hn = phx_tc_Hn(H);
    const int Bpad = phx_round_up(B, 128);
    int ks, per;
    phx_tc_ksplit(G, B, &ks, &per);
    float* base = reinterpret_cast<float*>(((uintptr_t)tcws + 127) & ~(uintptr_t)127);
    float* spart = base;
    float* spimg = spart + (size_t)ks * Bpad * 2 * Hn;
    const int dbg = debug_flags();
    const int nterms = (w.tc == 1) ? 1 : 3;

    BranchParams bp;
    bp.G = G; bp.B = B; bp.Bpad = Bpad; bp.Hn = Hn; bp.KB1 = phx_tc_KB1(G); bp.kb_per_split = per; bp.nterms = nterms;
    bp.a_lbo = 16 * 128; bp.a_sbo = 128; bp.b_lbo = (unsigned)(Hn / 8) * 128; bp.b_sbo = 128;
    bp.a_kadv = 2 * bp.a_lbo; bp.b_kadv = 2 * bp.b_lbo;
    if (dbg & 1) {   // diagnostic: swapped meaning of the two descriptor offsets
        unsigned t = bp.a_lbo; bp.a_lbo = bp.a_sbo; bp.a_sbo = t;
        t = bp.b_lbo; bp.b_lbo = bp.b_sbo; bp.b_sbo = t;
    }
    bp.y = y; bp.w1img = w.w1img; bp.spart = spart;
    const size_t stage1 = K1_A_BYTES + (size_t)4 * Hn * BK * 4;
    int S1 = (int)((PHX_SMEM_LIMIT - 256) / stage1);
    if (S1 > 4) S1 = 4;
    if (S1 < 2) {
        phx_set_error("tc branch kernel: stage of %zu bytes does not fit twice", stage1);
        return PHX_ERR_UNSUPPORTED;
    }
    bp.stages = S1;
    const size_t smem1 = (size_t)S1 * stage1 + 256;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(tc_branch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PHX_SMEM_LIMIT);
        cudaFuncSetAttribute(tc_joint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PHX_SMEM_LIMIT);
        attr_done = true;
    }
    tc_branch_kernel<<<dim3(Bpad / 128, ks), K1_THREADS, smem1, st>>>(bp);

    const int BT = phx_tc_BT(B);
    {
        const size_t total = (size_t)BT * 256 * (2 * Hn / 4);
        int blocks = (int)((total + 255) / 256);
        if (blocks > PHX_TC_SMS * 16) blocks = PHX_TC_SMS * 16;
        tc_spfinish_kernel<<<blocks, 256, 0, st>>>(B, Bpad, H, Hp, Hn, K2, ks, spart, w.bias, SP, spimg);
    }
    if (f) {
        JointParams jp;
        jp.G = G; jp.B = B; jp.KB2 = phx_tc_KB2(H); jp.GT = phx_tc_GT(G); jp.BT = BT; jp.nterms = nterms;
        jp.decay = decay; jp.fscale = fscale;
        jp.a_lbo = 16 * 128; jp.a_sbo = 128; jp.b_lbo = 32 * 128; jp.b_sbo = 128;
        jp.a_kadv = 2 * jp.a_lbo; jp.b_kadv = 2 * jp.b_lbo;
        if (dbg & 1) {
            unsigned t = jp.a_lbo; jp.a_lbo = jp.a_sbo; jp.a_sbo = t;
            t = jp.b_lbo; jp.b_lbo = jp.b_sbo; jp.b_sbo = t;
        }
        jp.waimg = w.waimg; jp.spimg = spimg; jp.y = y; jp.relum = w.relum; jp.f = f;
        const int ntiles = jp.GT * jp.BT;
        const int grid = ntiles < PHX_TC_SMS ? ntiles : PHX_TC_SMS;
        tc_joint_kernel<<<grid, K2_THREADS, (size_t)K2_STAGES * K2_STAGE_BYTES + 256, st>>>(jp);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        phx_set_error("tc rhs_forward launch: %s", cudaGetErrorString(e));
        return PHX_ERR_CUDA;
    }
    return PHX_OK;
}
