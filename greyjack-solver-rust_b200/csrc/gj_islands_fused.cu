// gj_islands_fused.cu -- translation unit of the fused TabuSearch / LateAcceptance island step
// (kernel: gj_islands_fused.cuh).  Separate from gj_islands.cu so that the eight template
// instantiations compile in parallel with the rest of the library.
#include "gj_islands_dev.cuh"
#include "gj_islands_fused.cuh"

gj_status gj_launch_fused_step(gj_islands* g, cudaStream_t st, bool trace) {
    const GjProblemDev& P = g->p->dev;
    GjFusedArgs F{};
    F.A = gj_make_select_args(g, trace, false);
    F.S = g->ds;
    F.symmetric = g->p->symmetric_D ? 1 : 0;
    F.n_clone = g->fused_clones;
    F.lean = g->fused_lean ? 1 : 0;
    F.scores_out = trace ? g->cand_scores : nullptr;
    F.moves_out = trace ? g->moves : nullptr;
    F.worklist = g->worklist;
    F.phase_clocks = g->phase_clocks;
    F.top_is_cur = g->top_is_cur; F.gedge = g->gedge; F.graw = g->graw; F.gedge_ver = g->gedge_ver;
    gj_status rc;
    // registers per thread are capped by the CTA size (64 K registers per SM): 1024 threads -> 64,
    // 512 -> 128.  The kernel keeps a neighbour's move, its score keys and the RNG in registers.
#define GJ_LAUNCH_FUSED(KIND, NT)                                                                   \
    do {                                                                                            \
        if ((rc = opt_in_smem(k_ls_step_fused<KIND, NT>, g->fused_smem))) return rc;               \
        k_ls_step_fused<KIND, NT><<<g->I, g->fused_threads, g->fused_smem, st>>>(P, g->groups, F);  \
    } while (0)
    const int nt = g->fused_threads;
    if (P.kind == GJ_NQUEENS) {
        if (nt > 512) GJ_LAUNCH_FUSED(GJ_NQUEENS, 1024);
        else if (nt > 256) GJ_LAUNCH_FUSED(GJ_NQUEENS, 512);
        else if (nt > 128) GJ_LAUNCH_FUSED(GJ_NQUEENS, 256);
        else GJ_LAUNCH_FUSED(GJ_NQUEENS, 128);
    } else {
        if (nt > 512) GJ_LAUNCH_FUSED(GJ_TSP, 1024);
        else if (nt > 256) GJ_LAUNCH_FUSED(GJ_TSP, 512);
        else if (nt > 128) GJ_LAUNCH_FUSED(GJ_TSP, 256);
        else GJ_LAUNCH_FUSED(GJ_TSP, 128);
    }
#undef GJ_LAUNCH_FUSED
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

