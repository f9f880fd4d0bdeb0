// Host-side planning of the resident solver launches (grid size, shared-memory and workspace layout).
#include "phx_resident.cuh"

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
    // one CTA per SM; at least 16 genes (one per warp) per CTA so small problems use few CTAs (cheaper all-reduce)
    int gpc = (G + num_sms - 1) / num_sms;
    if (gpc < 16) gpc = 16;
    int nCTA = (G + gpc - 1) / gpc;
    plan->nCTA = nCTA;
    plan->gpc = gpc;
    plan->NV = NV;
    plan->smem_bytes = smem_layout(B, K2, gpc, adjoint, nullptr, nullptr);
    if (plan->smem_bytes > PHX_SMEM_LIMIT) {
        phx_set_error("resident solver needs %zu B shared memory (> %d) for G=%d H=%d B=%d", plan->smem_bytes,
                      PHX_SMEM_LIMIT, G, H, B);
        return PHX_ERR_UNSUPPORTED;
    }
    return PHX_OK;
}

size_t phx_resident_workspace_floats(int nCTA, int G, int H, int B, int T, int adjoint, size_t* off_st,
                                     size_t* off_part, size_t* off_redout, size_t* off_partd, size_t* off_t,
                                     size_t* off_theta1) {
    const size_t K2 = (size_t)phx_K2(H);
    size_t off = 0;
    auto take = [&](size_t nfloats) {
        size_t o = off;
        off += (nfloats + 3) & ~size_t(3);
        return o;
    };
    size_t nslots = adjoint ? 18 : 9;
    size_t o_t = take(2 * (size_t)T);  // doubles
    size_t o_partd = take(2 * 2 * (size_t)nCTA * 8);
    size_t o_st = take(nslots * (size_t)B * G);
    size_t o_part = take((size_t)nCTA * B * K2);
    size_t o_red = take((size_t)B * K2);
    size_t o_th = adjoint ? take(phx_grad_offsets(G, H).total) : 0;
    if (off_st) *off_st = o_st;
    if (off_part) *off_part = o_part;
    if (off_redout) *off_redout = o_red;
    if (off_partd) *off_partd = o_partd;
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
