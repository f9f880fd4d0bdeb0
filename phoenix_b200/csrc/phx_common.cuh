// Shared definitions of the phoenix_b200 CUDA library (sm_100a).
//
// Data layout in HBM (all fp32):
//   packed weights  (built once per optimiser step by phx_pack_weights, constant during a solve)
//     W1[G][K2]   row g = [ Ws[0:H, g] | 0-pad to Hp | Wp[0:H, g] | 0-pad to Hp ]   (gene-major transposes of the
//                 two branch matrices, so the weights of one gene are one contiguous K2*4-byte row)
//     WA[G][K2]   row g = [ Wa[g, 0:H] | 0-pad | Wa[g, H:2H] | 0-pad ]
//     bias[K2]    [ bs | 0 | bp | 0 ]
//     relum[G]    relu(gene_multipliers)            maskm[G]  (gene_multipliers > 0)
//   Hp = round_up(H, 4), K2 = 2*Hp, so every row is a whole number of float4 and each float4 lies in one half.
//   solver state [slot][B][G]: every (row b, gene g) element is owned by exactly one thread of the CTA that owns
//   gene g, so state never needs cross-CTA synchronisation; only the K2-long branch vectors do.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/phoenix_b200.h"

#define PHX_THREADS 512
#define PHX_WARPS 16
#define PHX_MAX_B_FWD 8
#define PHX_MAX_B_ADJ 4
#define PHX_SMEM_LIMIT (227 * 1024)

static inline __host__ __device__ int phx_round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline __host__ __device__ int phx_Hp(int H) { return phx_round_up(H, 4); }
static inline __host__ __device__ int phx_K2(int H) { return 2 * phx_Hp(H); }

struct PhxPacked {
    const float4* W1;
    const float4* WA;
    const float* bias;
    const float* relum;
    const float* maskm;
};

static inline __host__ __device__ size_t phx_packed_floats(int G, int H) {
    size_t K2 = (size_t)phx_K2(H);
    return 2 * (size_t)G * K2 + K2 + 2 * (size_t)phx_round_up(G, 4);
}

static inline __host__ __device__ PhxPacked phx_packed_view(const float* base, int G, int H) {
    size_t K2 = (size_t)phx_K2(H);
    PhxPacked v;
    v.W1 = reinterpret_cast<const float4*>(base);
    v.WA = reinterpret_cast<const float4*>(base + (size_t)G * K2);
    v.bias = base + 2 * (size_t)G * K2;
    v.relum = v.bias + K2;
    v.maskm = v.relum + phx_round_up(G, 4);
    return v;
}

// flat-gradient offsets, reference parameter order
struct PhxGradOff {
    size_t m, Wp, bp, Ws, bs, Wa, total;
};
static inline __host__ __device__ PhxGradOff phx_grad_offsets(int G, int H) {
    PhxGradOff o;
    o.m = 0;
    o.Wp = (size_t)G;
    o.bp = o.Wp + (size_t)H * G;
    o.Ws = o.bp + (size_t)H;
    o.bs = o.Ws + (size_t)H * G;
    o.Wa = o.bs + (size_t)H;
    o.total = o.Wa + (size_t)G * 2 * H;
    return o;
}

// Parameters of the persistent ("resident") solver kernels.
struct ResParams {
    int G, H, Hp, K2, K2q, B, T, method, gpc, t_is_f32, adjoint;
    float rtol_f, atol_f;
    float fsign;           // +1, or -1 when the caller's t was decreasing (f -> -f(-t, y), misc.py:159-162)
    long long max_steps;
    PhxPacked w;
    const double* t;       // [T] device copy of the output times
    // forward
    const float* y0;       // [B][G]
    float* yout;           // [T][B][G]
    // adjoint
    const float* ysaved;   // [T][B][G]
    const float* grad_y;   // [T][B][G]
    float* adj_y0;         // [B][G]
    float* theta0;         // [P] caller's flat grads (result lands here)
    float* theta1;         // [P] scratch twin
    // workspace
    float* st;             // state slots [nslots][B][G]
    float* part;           // [nCTA][B*K2] all-reduce partials
    float* redout;         // [B*K2]
    double* partd;         // [2][nCTA][8]
    phx_status* status;
    double* steplog;
    int steplog_cap;
    long long* prof;       // optional [PHX_PROF_SLOTS] phase-timer accumulators (phx_ctx_set_profile), else nullptr
};
#define PHX_PROF_SLOTS 32

struct ResLaunchPlan {
    int nCTA, gpc, NV;
    size_t smem_bytes;
};

// host helpers implemented in phx_resident.cu
int phx_resident_plan(int num_sms, int G, int H, int B, int adjoint, ResLaunchPlan* plan);
size_t phx_resident_workspace_floats(int nCTA, int G, int H, int B, int T, int adjoint, size_t* off_st,
                                     size_t* off_part, size_t* off_redout, size_t* off_partd, size_t* off_t,
                                     size_t* off_theta1);
int phx_resident_launch(const ResParams& p, const ResLaunchPlan& plan, cudaStream_t stream);

// host helpers implemented in phx_rhs.cu
int phx_rhs_forward_launch(int G, int H, int B, const PhxPacked& w, const float* y, float* f, int decay, float fscale,
                           float* ws, cudaStream_t stream);
int phx_rhs_vjp_launch(int G, int H, int B, const PhxPacked& w, const float* y, const float* g, int decay,
                       float* ybar, float* grads_flat, int accumulate, float* f_out, float fscale, float* ws,
                       cudaStream_t stream);
size_t phx_rhs_workspace_floats(int G, int H, int B);

void phx_set_error(const char* fmt, ...);
