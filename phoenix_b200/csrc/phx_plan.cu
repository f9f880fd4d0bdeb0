// Host-side planning of the resident solver launches (grid size, shared-memory and workspace layout).
#include <stdlib.h>
#include "phx_common.cuh"

int phx_launch_fwd_nv1(const ResParams&, const ResLaunchPlan&, cudaStream_t);
int phx_launch_fwd_nv2(const ResParams&, const ResLaunchPlan&, cudaStream_t);
int phx_launch_fwd_nv4(const ResParams&, const ResLaunchPlan&, cudaStream_t);
int phx_launch_adj_nv1(const ResParams&, const ResLaunchPlan&, cudaStream_t);
int phx_launch_adj_nv2(const ResParams&, const ResLaunchPlan&, cudaStream_t);
int phx_launch_adj_nv4(const ResParams&, const ResLaunchPlan&, cudaStream_t);

// ---- host side -------------------------------------------------------------------------------------------------------
int phx_resident_plan(int num_sms, int G, int H, int B, int adjoint, ResLaunchPlan* plan) {
    const int K2 = phx_K2(H);
    const int K2q = K2 / 4;
    int NV = (K2q + 31) / 32;
    if (NV > 4) {
        phx_set_error("resident solver supports neurons <= 256 (got H=%d)", H);
        return PHX_ERR_UNSUPPORTED;
    }
    NV = (NV <= 1) ? 1 : (NV == 2 ? 2 : 4);
    if (B < 1 || B > (adjoint ? PHX_MAX_B_ADJ : PHX_MAX_B_FWD)) {
        phx_set_error("resident solver supports 1 <= B <= %d rows here (got %d)",
                      adjoint ? PHX_MAX_B_ADJ : PHX_MAX_B_FWD, B);
        return PHX_ERR_UNSUPPORTED;
    }
    if (num_sms > PHX_LL_MAXC) num_sms = PHX_LL_MAXC;
    // one CTA per SM; at least 16 genes (one per warp) per CTA so small problems use few CTAs (cheaper exchanges).
    // Models of 256 .. 2048 genes take 32 genes per CTA: measured on B200 the latency of one solve is the same (SIM690
    // training step 5.2 ms either way) while twice as many independent solves fit on the GPU side by side
    // (odeint_adjoint_many: 3.0 -> 2.0 ms per SIM690 step).
    int gpc = (G + num_sms - 1) / num_sms;
    int gmin = (G >= 256 && G <= 2048) ? 32 : 16;
    if (const char* e = getenv("PHX_MIN_GENES_PER_CTA")) gmin = atoi(e) > 0 ? atoi(e) : gmin;   // experiments
    if (gpc < gmin) gpc = gmin;
    int nCTA = (G + gpc - 1) / gpc;
    plan->nCTA = nCTA;
    plan->gpc = gpc;
    plan->NV = NV;
    // Where the CTA's two weight slices live.  Preference: both resident in shared memory for the whole solve (every
    // pass is a shared-memory pass; all shipped configs up to the yeast shape) -> W1 resident (it is read twice per
    // VJP evaluation) and WA streamed from L2 through the ring (the 11k-gene breast shape) -> both streamed.
    // Ring geometry: every warp owns a private ring of S row-sized slots (S = 4..1, as many as fit).
    auto fits = [&](int w1r, int war, int stages) {
        return phx_smem_layout(nCTA, B, K2, gpc, adjoint, w1r, war, PHX_WARPS, stages, nullptr) <= PHX_SMEM_LIMIT;
    };
    int w1r = 0, war = 0, best_rows = 0, best_stages = 0;
    bool ok = false;
    if (fits(1, 1, 0)) {
        w1r = war = 1;
        ok = true;
    }
    // Tensor memory as a second on-chip weight store: 512 columns x 128 lanes x 32 bit per SM.  A warp can only touch
    // the 32 lanes of its quarter (warp % 4), the four warps of a quarter split the 512 columns, and a row costs
    // 4 NV columns per lane (the lane's own float4 columns of the row): 32 / NV rows per warp.
    const int rows_per_warp = (gpc + PHX_WARPS - 1) / PHX_WARPS;
    const bool tmem_ok = rows_per_warp * 4 * NV <= 128 && !getenv("PHX_NO_TMEM");
    if (!ok && tmem_ok) {
        if (fits(1, 2, 0)) {
            w1r = 1;
            war = 2;
            ok = true;
        } else {
            for (int stages = 4; stages >= 1 && !ok; --stages) {
                if (!fits(0, 2, stages)) continue;
                w1r = 0;
                war = 2;
                best_rows = PHX_WARPS;
                best_stages = stages;
                ok = true;
            }
        }
    }
    for (int res = 1; res >= 0 && !ok; --res) {
        for (int stages = 4; stages >= 1 && !ok; --stages) {
            if (!fits(res, 0, stages)) continue;
            w1r = res;
            war = 0;
            best_rows = PHX_WARPS;
            best_stages = stages;
            ok = true;
        }
    }
    if (!ok) {
        phx_set_error("resident solver does not fit in %d B shared memory for G=%d H=%d B=%d", PHX_SMEM_LIMIT, G, H, B);
        return PHX_ERR_UNSUPPORTED;
    }
    plan->w1_res = w1r;
    plan->wa_res = war;
    plan->ring_rows = best_rows;
    plan->ring_stages = best_stages;
    plan->smem_bytes = phx_smem_layout(nCTA, B, K2, gpc, adjoint, w1r, war, best_rows, best_stages, &plan->so);
    return PHX_OK;
}

// workspace: [ LL exchange area (zeroed once by phx_solve_workspace_init) | t[T] doubles | theta twin (adjoint) ]
size_t phx_resident_workspace_floats(int G, int H, int B, int T, int adjoint, size_t* off_t, size_t* off_theta1) {
    size_t off = phx_ll_words() * 2;
    auto take = [&](size_t nfloats) {
        size_t o = off;
        off += (nfloats + 3) & ~size_t(3);
        return o;
    };
    size_t o_t = take(2 * (size_t)T);  // doubles
    size_t o_th = adjoint ? take(phx_grad_offsets(G, H).total) : 0;
    if (off_t) *off_t = o_t;
    if (off_theta1) *off_theta1 = o_th;
    return off;
}

int phx_resident_launch(const ResParams& p, const ResLaunchPlan& plan, cudaStream_t stream) {
    if (p.adjoint) {
        if (plan.NV == 1) return phx_launch_adj_nv1(p, plan, stream);
        if (plan.NV == 2) return phx_launch_adj_nv2(p, plan, stream);
        return phx_launch_adj_nv4(p, plan, stream);
    }
    if (plan.NV == 1) return phx_launch_fwd_nv1(p, plan, stream);
    if (plan.NV == 2) return phx_launch_fwd_nv2(p, plan, stream);
    return phx_launch_fwd_nv4(p, plan, stream);
}


// ---- rows kernels ------------------------------------------------------------------------------------------------------
int phx_launch_rows_fwd(const ResParams&, const RowsPlan&, cudaStream_t);
int phx_launch_rows_adj(const ResParams&, const RowsPlan&, cudaStream_t);

int phx_rows_plan(int num_sms, int G, int H, int adjoint, RowsPlan* plan) {
    const int K2 = phx_K2(H), K2q = K2 / 4;
    if (K2q > 128) {
        phx_set_error("rows solver supports neurons <= 256 (got H=%d)", H);
        return PHX_ERR_UNSUPPORTED;
    }
    if (getenv("PHX_NO_ROWS")) {
        phx_set_error("rows solver disabled (PHX_NO_ROWS)");
        return PHX_ERR_UNSUPPORTED;
    }
    if (num_sms > PHX_LL_MAXC) num_sms = PHX_LL_MAXC;
    int gpc = (G + num_sms - 1) / num_sms;
    int gmin = (G >= 256 && G <= 2048) ? 32 : 16;
    if (const char* e = getenv("PHX_MIN_GENES_PER_CTA")) gmin = atoi(e) > 0 ? atoi(e) : gmin;
    if (gpc < gmin) gpc = gmin;
    const int nCTA = (G + gpc - 1) / gpc;
    // a quad-warp covers 32 float4 columns of the K2-long rows; the 16 warps are nqw quad-warps x ngg gene groups
    int nqw = K2q <= 32 ? 1 : (K2q <= 64 ? 2 : 4);
    // W1 rows are padded in shared memory so that two consecutive rows are 64 bytes apart modulo 128: the
    // (gene, 4 k-chunks) threads of the state-cotangent pass then hit distinct banks
    int w1s = K2q;
    while ((w1s & 7) != 4) ++w1s;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int ngg = PHX_WARPS / nqw;
        const int gpg = phx_round_up((gpc + ngg - 1) / ngg, 4);
        for (int wa_res = 1; wa_res <= 2; ++wa_res) {
            if (wa_res == 2 && nqw != 4) continue;   // tensor-memory lanes are tied to the warp's quarter: lane = quad
            const int tm_wa = wa_res == 2 ? 4 * ngg * gpg : 0;
            int rows = PHX_ROWS_MAX;
            if (adjoint)
                while (rows > 0 && tm_wa + 2 * PHX_ROWS_NFS * 4 * rows > 512) --rows;
            else if (tm_wa > 512) rows = 0;
            if (rows == 0) continue;
            SmemOff so;
            size_t bytes = phx_rows_smem_layout(nCTA, K2, gpc, adjoint, wa_res, nqw, w1s, &so);
            if (bytes > PHX_SMEM_LIMIT) continue;
            plan->nCTA = nCTA; plan->gpc = gpc; plan->nqw = nqw; plan->ngg = ngg; plan->gpg = gpg;
            plan->wa_res = wa_res; plan->rows = rows; plan->w1_stride_q = w1s; plan->tm_wa = tm_wa;
            plan->tm_fac = tm_wa; plan->smem_bytes = bytes; plan->so = so;
            return PHX_OK;
        }
        if (nqw == 4) break;
        nqw = 4;   // second attempt: the tensor-memory placement needs the 4 x 4 warp grid
    }
    phx_set_error("rows solver: the weight slices of G=%d H=%d do not fit on chip", G, H);
    return PHX_ERR_UNSUPPORTED;
}

// workspace: [ LL exchange area | t[N*T] doubles | per-row packed cotangent double buffers (adjoint) ]
size_t phx_rows_workspace_floats(int G, int H, int N, int T, int rows, int adjoint, size_t* off_t, size_t* off_theta) {
    size_t off = phx_ll_words() * 2;
    auto take = [&](size_t nfloats) {
        size_t o = off;
        off += (nfloats + 31) & ~size_t(31);
        return o;
    };
    size_t o_t = take(2 * (size_t)T * N);
    size_t o_th = adjoint ? take(2 * (size_t)rows * phx_packed_grad_offsets(G, H).total) : 0;
    if (off_t) *off_t = o_t;
    if (off_theta) *off_theta = o_th;
    return off;
}

int phx_rows_launch(const ResParams& p, const RowsPlan& plan, cudaStream_t stream) {
    return p.adjoint ? phx_launch_rows_adj(p, plan, stream) : phx_launch_rows_fwd(p, plan, stream);
}
