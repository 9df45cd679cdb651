// gj_islands_vrp_chain.cu -- translation unit of the VRP LateAcceptance / SimulatedAnnealing chains
// (kernels: gj_islands_vrp_chain.cuh).
#include <algorithm>
#include <cstdlib>

#include "gj_islands_dev.cuh"
#include "gj_islands_vrp_chain.cuh"

gj_status gj_launch_vrp_gindex(gj_islands* g, cudaStream_t st) {
    const GjProblemDev& P = g->p->dev;
    const size_t smem = gj_vrp_smem_bytes(P.n_entities, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
    if (smem <= 200 * 1024) {
        gj_status rc;
        if ((rc = opt_in_smem(k_vrp_chain_gindex_cta, smem))) return rc;
        k_vrp_chain_gindex_cta<<<1, kVrpWarps * 32, smem, st>>>(P, g->I, g->gbest, g->gver, g->vcs);
        GJ_LAUNCH_CHECK();
        return GJ_OK;
    }
    k_vrp_chain_gindex<<<1, kGindexWarps * 32, 0, st>>>(g->p->dev, g->I, g->gbest, g->gver, g->vcs);
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

gj_status gj_launch_vrp_chains(gj_islands* g, const GjChainArgs& A, cudaStream_t st) {
    const GjProblemDev& P = g->p->dev;
    gj_status rc;
    // the global top's route index, if a newer top was published since it was built (else: one CTA, two loads)
    if (g->gver && (rc = gj_launch_vrp_gindex(g, st))) return rc;
    const size_t esmem = gj_vrp_smem_bytes(P.n_entities, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
    const int cta_rebuild = esmem <= 200 * 1024 ? 1 : 0;
    k_vrp_chain_prepare<<<(unsigned)((g->I + kVrpChainWarps - 1) / kVrpChainWarps), kVrpChainWarps * 32, 0, st>>>(P, A, g->vcs, cta_rebuild);
    GJ_LAUNCH_CHECK();
    if (cta_rebuild) {
        if ((rc = opt_in_smem(k_vrp_chain_rebuild_cta, esmem))) return rc;
        k_vrp_chain_rebuild_cta<<<(unsigned)g->I, kVrpWarps * 32, esmem, st>>>(P, g->stride, g->cur, g->vcs);
        GJ_LAUNCH_CHECK();
    }
    // warps (chains) per CTA: as many as share an SM anyway, so that every SM gets chains and the
    // warps of a CTA -- which re-align every step -- are the ones that share its instruction cache
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g->p->device);
    // (not a power of two: 4096 chains are 28 per SM on 147 SMs, where 32 per CTA left 20 SMs idle)
    int vw = (int)std::min<int64_t>(kVrpStepWarps, std::max<int64_t>(4, (g->I + sms - 1) / sms));
    if (const char* e = getenv("GJ_VRPC_CTA_WARPS")) vw = std::max(1, std::min(kVrpStepWarps, atoi(e)));   // development knob
    const unsigned vgrid = (unsigned)((g->I + vw - 1) / vw);
    const size_t vsmem = sizeof(GjVrpcScratch) * vw;
    if (g->prm.agent == GJ_AGENT_LATE_ACCEPTANCE) {
        if ((rc = opt_in_smem(k_vrp_chains<GJ_AGENT_LATE_ACCEPTANCE>, sizeof(GjVrpcScratch) * kVrpStepWarps))) return rc;
        k_vrp_chains<GJ_AGENT_LATE_ACCEPTANCE><<<vgrid, vw * 32, vsmem, st>>>(P, g->groups, A, g->vcs);
    } else {
        if ((rc = opt_in_smem(k_vrp_chains<GJ_AGENT_SIMULATED_ANNEALING>, sizeof(GjVrpcScratch) * kVrpStepWarps))) return rc;
        k_vrp_chains<GJ_AGENT_SIMULATED_ANNEALING><<<vgrid, vw * 32, vsmem, st>>>(P, g->groups, A, g->vcs);
    }
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}
