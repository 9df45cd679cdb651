// gj_islands_chain.cu -- translation unit of the N-Queens / TSP LateAcceptance and
// SimulatedAnnealing chains (kernel: gj_islands_chain.cuh).
#include "gj_islands_dev.cuh"
#include "gj_islands_chain.cuh"

gj_status gj_launch_la_chains(gj_islands* g, const GjChainArgs& A, cudaStream_t st) {
    const GjProblemDev& P = g->p->dev;
    const size_t smem = g->chain_bytes * kChainWarps;
    const unsigned grid = (unsigned)((g->I + kChainWarps - 1) / kChainWarps);
    gj_status rc;
    if (P.kind == GJ_NQUEENS) {
        if ((rc = opt_in_smem(k_la_chains<GJ_NQUEENS>, smem))) return rc;
        k_la_chains<GJ_NQUEENS><<<grid, kChainWarps * 32, smem, st>>>(P, g->groups, A, g->chain_bytes);
    } else {
        if ((rc = opt_in_smem(k_la_chains<GJ_TSP>, smem))) return rc;
        k_la_chains<GJ_TSP><<<grid, kChainWarps * 32, smem, st>>>(P, g->groups, A, g->chain_bytes);
    }
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}
