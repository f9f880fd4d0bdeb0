// Tensor-core (tcgen05 / TMEM) path of the batched RHS: shared layout definitions.
//
// When a call carries more rows than the resident solver kernels take (B >= PHX_TC_MIN_ROWS: the 60-row gene-influence
// scan, the 10 000-row prior batch of train_insilico.py:134, the 4 096-row synthetic sweep), the contractions run on the 5th-generation tensor cores as
// TF32 MMAs with fp32 accumulators in tensor memory.  fp32 parity with the reference (odenet.py:85-91 evaluated by ATen
// in fp32) is kept by the 3xTF32 split: every operand x is stored as hi = rna_tf32(x), lo = rna_tf32(x - hi) and a
// product is accumulated as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (PHX_PREC_3XTF32).  PHX_PREC_TF32 issues only hi*hi and is
// reported separately with its own tolerance.
//
// Operand images.  tcgen05.mma reads its operands from shared memory through matrix descriptors; we use the K-major
// NO-SWIZZLE canonical layout, whose unit is the 8-row x 16-byte "core matrix" stored as 128 contiguous bytes:
//     tile of R rows x 16 floats (one k-block):  float offset(r, k) = (((k/4) * (R/8) + r/8) * 8 + r%8) * 4 + k%4
//     descriptor: leading byte offset (K-adjacent core matrices) = (R/8)*128, stride byte offset (row groups) = 128.
// The constant operands (weights) are re-laid into exactly this image once per weight update (tc_pack_kernel), k-block
// by k-block and hi|lo side by side, so one k-block of an operand is ONE contiguous chunk of global memory and is
// fetched with a single 1-D TMA bulk copy (cp.async.bulk + mbarrier complete_tx) -- no tensor maps, no swizzle.
//     w1img [KB1][branch 2][hi|lo][Hn x 16]   branch contraction S|P = act(y) W1      (N = Hn hidden units, K = genes)
//     waimg [GT][KB2][hi|lo][128 x 16]        joint contraction J = [S|P] WA^T         (M = genes, K = 2*Hn)
//     spimg [BT][KB2][hi|lo][256 x 16]        [S|P] rows, written by the K-split reduction kernel (N = batch rows)
//     watimg = w1img layout built from WA     cotangent contraction gSP = gJ WA       (N = Hn columns of a half, K = genes)
//     w1kimg = waimg layout built from W1     state cotangent u|v = gSP W1^T          (M = genes, K = Hn of one half)
//     gsimg  = spimg layout holding gSP
// Hn = round_up(H, 16) (UMMA N granularity at M = 128); pads are zero in every image.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <mutex>
// (included from phx_common.cuh after phx_round_up)

#define PHX_TC_BK 16
#define PHX_TC_MIN_ROWS 5   /* more rows than the resident adjoint kernel takes: measured 40-55x faster than the
                             fp32 SGEMM path already at 17-60 rows (128-row tiles mostly padding, still far ahead) */
#define PHX_TC_MAX_HN 256
#define PHX_TC_SMS 148
#define PHX_TC_MAX_KSPLIT 64

static inline __host__ __device__ int phx_tc_Hn(int H) { return phx_round_up(H, 16); }
static inline __host__ __device__ int phx_tc_KB1(int G) { return (G + PHX_TC_BK - 1) / PHX_TC_BK; }
static inline __host__ __device__ int phx_tc_KB2(int H) { return 2 * phx_tc_Hn(H) / PHX_TC_BK; }
static inline __host__ __device__ int phx_tc_GT(int G) { return (G + 127) / 128; }
static inline __host__ __device__ int phx_tc_BT(int B) { return (B + 255) / 256; }
static inline int phx_tc_min_rows_rt() {
    static int v = 0;
    if (!v) {
        const char* e = getenv("PHX_TC_MIN_ROWS");   // experiments only
        v = e && atoi(e) > 0 ? atoi(e) : PHX_TC_MIN_ROWS;
    }
    return v;
}
static inline bool phx_tc_shape_ok(int H, int B) { return B >= phx_tc_min_rows_rt() && phx_tc_Hn(H) <= PHX_TC_MAX_HN; }

static inline __host__ __device__ size_t phx_tc_w1img_floats(int G, int H) {
    return (size_t)phx_tc_KB1(G) * 4 * phx_tc_Hn(H) * PHX_TC_BK;
}
static inline __host__ __device__ size_t phx_tc_waimg_floats(int G, int H) {
    return (size_t)phx_tc_GT(G) * phx_tc_KB2(H) * 2 * 128 * PHX_TC_BK;
}
static inline __host__ __device__ size_t phx_tc_spimg_floats(int H, int B) {
    return (size_t)phx_tc_BT(B) * phx_tc_KB2(H) * 2 * 256 * PHX_TC_BK;
}
// [S|P] / gSP as the B operand of the K = batch contractions: [ceil(B/16)][branch 2][hi|lo][Hn x 16]
static inline __host__ __device__ size_t phx_tc_timg_floats(int H, int B) {
    return (size_t)((B + PHX_TC_BK - 1) / PHX_TC_BK) * 4 * phx_tc_Hn(H) * PHX_TC_BK;
}
// float offset of element (r, k) inside an R-row x 16-float tile image
static inline __host__ __device__ int phx_tc_tile_off(int R, int r, int k) {
    return ((((k >> 2) * (R >> 3) + (r >> 3)) * 8 + (r & 7)) << 2) + (k & 3);
}
// Operands of the branch-type kernel (tc_branch_kernel): -DPHX_TC_BRANCH_SW64=1 selects the K-major SWIZZLE_64B
// canonical layout instead (rows of 16 floats = 64 bytes, 8-row x 64-byte atoms, the 16-byte chunk index XORed with
// bits 1-2 of the row).  Measured on B200: identical results and identical speed (0.79 ms per RHS evaluation at
// 20 000 x 200 x 4 096 either way), so the default stays the no-swizzle core-matrix layout above.
#ifndef PHX_TC_BRANCH_SW64
#define PHX_TC_BRANCH_SW64 0
#endif
static inline __host__ __device__ int phx_tc_btile_off(int R, int r, int k) {
#if PHX_TC_BRANCH_SW64
    return r * 16 + ((((k >> 2) ^ ((r >> 1) & 3))) << 2) + (k & 3);
#else
    return phx_tc_tile_off(R, r, k);
#endif
}

// Accumulation-chain limit of the branch contraction: k-blocks summed inside tensor memory before the chunk sum is
// drained to its own partial-sum slot (phx_tc.cu, "Accumulation chains").  16 k-blocks = 96 accumulating MMAs.
static inline int phx_tc_chunk() {
    static int chunk = 0;
    if (!chunk) {
        const char* e = getenv("PHX_TC_CHUNK");
        int v = e ? atoi(e) : 16;
        chunk = v > 0 ? v : 16;
    }
    return chunk;
}
// Work split of the branch contraction (tc_branch_kernel): every CTA owns one branch, one 128-row tile and a K range of
// whole chunks; one CTA runs per SM at a time and the hardware hands the next CTA (prods CTAs are launched first) to the
// first SM that frees up.  The numbers of K ranges per branch (ks_p, ks_s) are chosen by SIMULATING that dispatch with the
// measured cost per k-block (log1p CTAs ~1 245 cycles, soft-sign ~1 160, cotangent operand ~1 100; ~30 000 cycles of
// prologue + final fold per CTA; each further partial-sum slot costs the finishing pass a read of the slot) and taking
// the split with the shortest makespan -- with few row tiles "two CTAs per SM" quantises badly: 79 tiles (the 10 000-row
// prior batch) gave 2 + 1 ranges and a 1.6-wave schedule 28 % above the balanced time.
struct PhxTcBranchPlan {
    int mtiles, ks_p, per_p, ks_s, per_s, slots;
    int ks_e, per_e;   // equal-cost halves (MODE 1: the cotangent operand): ranges per half
};
static inline double phx_tc_makespan(int mt, int n_first, double d_first, int n_second, double d_second) {
    // n_first CTAs of duration d_first, then n_second of d_second, dispatched in order to PHX_TC_SMS SMs
    double t[PHX_TC_SMS];
    for (int i = 0; i < PHX_TC_SMS; ++i) t[i] = 0.0;
    // equal durations within a class: fill in rounds -- SM free times stay sorted ascending if we always take the minimum;
    // keep a simple binary heap
    auto sift = [&](int i) {
        for (;;) {
            int l = 2 * i + 1, r = l + 1, m = i;
            if (l < PHX_TC_SMS && t[l] < t[m]) m = l;
            if (r < PHX_TC_SMS && t[r] < t[m]) m = r;
            if (m == i) break;
            double x = t[i]; t[i] = t[m]; t[m] = x;
            i = m;
        }
    };
    (void)mt;
    for (int c = 0; c < n_first + n_second; ++c) {
        t[0] += c < n_first ? d_first : d_second;
        sift(0);
    }
    double mx = 0.0;
    for (int i = 0; i < PHX_TC_SMS; ++i) mx = t[i] > mx ? t[i] : mx;
    return mx;
}
static inline PhxTcBranchPlan phx_tc_branch_plan_compute(int G, int B);
// the simulation costs milliseconds: plans are cached per (K, M) (a handful of shapes per process; per translation unit)
static inline PhxTcBranchPlan phx_tc_branch_plan(int G, int B) {
    struct Entry { int G, B; PhxTcBranchPlan pl; };
    static Entry cache[64];
    static int n = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n; ++i)
        if (cache[i].G == G && cache[i].B == B) return cache[i].pl;
    PhxTcBranchPlan pl = phx_tc_branch_plan_compute(G, B);
    if (n < 64) {
        cache[n].G = G; cache[n].B = B; cache[n].pl = pl;
        ++n;
    }
    return pl;
}
static inline PhxTcBranchPlan phx_tc_branch_plan_compute(int G, int B) {
    PhxTcBranchPlan pl;
    const int KB1 = phx_tc_KB1(G), chunk = phx_tc_chunk();
    const int nch = (KB1 + chunk - 1) / chunk;   // chunks along K
    pl.mtiles = (B + 127) / 128;
    const double c_p = 1245.0, c_s = 1160.0, c_e = 1100.0, ovh = 30000.0;
    const int kmax = nch < 24 ? nch : 24;
    const double slot_cost = 6.0 * pl.mtiles * 128.0 / PHX_TC_SMS * 2.0;   // finishing pass: cycles per extra slot (rough)
    double best = -1.0;
    int bp = 1, bs = 1;
    for (int kp = 1; kp <= kmax; ++kp) {
        const int per_p = (nch + kp - 1) / kp;   // chunks per prods CTA
        if ((nch + per_p - 1) / per_p != kp) continue;   // same split as a smaller kp
        for (int ks = 1; ks <= kmax; ++ks) {
            const int per_s = (nch + ks - 1) / ks;
            if ((nch + per_s - 1) / per_s != ks) continue;
            if ((long long)pl.mtiles * (kp + ks) > 16 * PHX_TC_SMS) continue;
            const double m = phx_tc_makespan(pl.mtiles, pl.mtiles * kp, ovh + per_p * chunk * c_p, pl.mtiles * ks,
                                             ovh + per_s * chunk * c_s) + slot_cost * (kp > ks ? kp : ks);
            if (best < 0.0 || m < best) { best = m; bp = kp; bs = ks; }
        }
    }
    if (const char* e = getenv("PHX_TC_KS")) {   // experiments: "ks_p,ks_s" for every shape of the process
        int a = 0, b2 = 0;
        if (sscanf(e, "%d,%d", &a, &b2) == 2 && a >= 1 && b2 >= 1) {
            bp = a < nch ? a : nch;
            bs = b2 < nch ? b2 : nch;
        }
    }
    pl.per_p = (nch + bp - 1) / bp * chunk;
    pl.per_s = (nch + bs - 1) / bs * chunk;
    pl.ks_p = (KB1 + pl.per_p - 1) / pl.per_p;
    pl.ks_s = (KB1 + pl.per_s - 1) / pl.per_s;
    best = -1.0;
    int be = 1;
    for (int ke = 1; ke <= kmax; ++ke) {
        const int per_e = (nch + ke - 1) / ke;
        if ((nch + per_e - 1) / per_e != ke) continue;
        if ((long long)pl.mtiles * 2 * ke > 16 * PHX_TC_SMS) continue;
        const double m = phx_tc_makespan(pl.mtiles, pl.mtiles * 2 * ke, ovh + per_e * chunk * c_e, 0, 0.0) + slot_cost * ke;
        if (best < 0.0 || m < best) { best = m; be = ke; }
    }
    pl.per_e = (nch + be - 1) / be * chunk;
    pl.ks_e = (KB1 + pl.per_e - 1) / pl.per_e;
    pl.slots = pl.ks_p > pl.ks_s ? pl.ks_p : pl.ks_s;
    if (pl.ks_e > pl.slots) pl.slots = pl.ks_e;
    return pl;
}
// scratch of the tensor-core RHS in floats (partial-sum slots + [S|P] operand image), 128-byte aligned inside
static inline size_t phx_tc_scratch_floats(int G, int H, int B) {
    const PhxTcBranchPlan pl = phx_tc_branch_plan(G, B);
    const size_t Bpad = (size_t)phx_round_up(B, 256);   // 128-row tiles, paired
    const PhxTcBranchPlan pg = phx_tc_branch_plan(B, G);   // K = batch contractions (parameter cotangents)
    return (size_t)pl.slots * Bpad * 2 * phx_tc_Hn(H) + 2 * phx_tc_spimg_floats(H, B) + 2 * phx_tc_timg_floats(H, B) +
           (size_t)pg.slots * phx_round_up(pg.mtiles, 2) * 128 * 2 * phx_tc_Hn(H) + 64;
}
