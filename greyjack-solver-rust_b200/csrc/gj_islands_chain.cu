// gj_islands_chain.cu -- translation unit of the N-Queens / TSP LateAcceptance and
// SimulatedAnnealing chains (kernel: gj_islands_chain.cuh).
#include "gj_islands_dev.cuh"
#include "gj_islands_chain.cuh"

template <int KIND>
static gj_status launch_la(gj_islands* g, const GjChainArgs& A, cudaStream_t st) {
    const GjProblemDev& P = g->p->dev;
    gj_status rc;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g->p->device);
    // more chains per SM than the narrow variant keeps resident (5 CTAs of 4 at 96 registers): one wide CTA
    // per SM, as many chains as the SM's share -- when their shared-memory slices fit
    const int per_sm = (int)((g->I + sms - 1) / sms);
    const size_t static_smem = sizeof(GjMove) * kChainWarpsWide + 64;
    if (per_sm > 20 && per_sm <= kChainWarpsWide && g->chain_bytes * per_sm + static_smem <= 227 * 1024) {
        const size_t smem = g->chain_bytes * per_sm;
        const unsigned grid = (unsigned)((g->I + per_sm - 1) / per_sm);
        if ((rc = opt_in_smem(k_la_chains<KIND, kChainWarpsWide>, smem))) return rc;
        k_la_chains<KIND, kChainWarpsWide><<<grid, per_sm * 32, smem, st>>>(P, g->groups, A, g->chain_bytes);
    } else {
        const size_t smem = g->chain_bytes * kChainWarps;
        const unsigned grid = (unsigned)((g->I + kChainWarps - 1) / kChainWarps);
        if ((rc = opt_in_smem(k_la_chains<KIND, kChainWarps>, smem))) return rc;
        k_la_chains<KIND, kChainWarps><<<grid, kChainWarps * 32, smem, st>>>(P, g->groups, A, g->chain_bytes);
    }
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

gj_status gj_launch_la_chains(gj_islands* g, const GjChainArgs& A, cudaStream_t st) {
    return g->p->dev.kind == GJ_NQUEENS ? launch_la<GJ_NQUEENS>(g, A, st) : launch_la<GJ_TSP>(g, A, st);
}
