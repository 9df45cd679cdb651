// gj_islands_tsfast.cu -- translation unit of the fixed-point TabuSearch step for TSP
// (kernel: gj_islands_tsfast.cuh).
#include "gj_islands_dev.cuh"
#include "gj_islands_tsfast.cuh"

// Mover::do_move picks the first move whose cumulative probability is >= u (mover.rs:105-121); the
// generator draws u = x / 2^32 from a 32-bit x and counts the thresholds below it (gj_generate_move).
// For an integer x:  x / 2^32 > t  <=>  x > floor(t * 2^32)  (t * 2^32 is exact in f64), so the same
// decision can be made on the raw draw.
void gj_kind_thresholds_u32(const double* thr, uint32_t* out) {
    for (int i = 0; i < 5; ++i) {
        const double y = thr[i] * 4294967296.0;
        out[i] = (y >= 4294967295.0) ? 0xffffffffu : (y <= 0.0 ? 0u : (uint32_t)std::floor(y));
    }
}

size_t gj_tsfast_smem(const gj_islands* g) {
    return gj_tsfast_smem_bytes(g->n_vars, g->tabu_words, 32 * g->p->dev.bm_words);
}

gj_status gj_launch_tsfast_step(gj_islands* g, cudaStream_t st, bool trace) {
    const GjProblemDev& P = g->p->dev;
    GjTsFastArgs F{};
    F.A = gj_make_select_args(g, trace, false);
    F.edge = g->ts_edge;
    F.edge_stride = g->ts_edge_stride;
    F.stale = g->ds.stale;
    gj_kind_thresholds_u32(g->mover.thresholds, F.kind_thr);
    F.first = g->p->groups[0][0];
    F.glen = (int)g->p->groups[0].size();
    F.cnt_stride = 32 * P.bm_words;
    F.scores_out = trace ? g->cand_scores : nullptr;
    F.moves_out = trace ? g->moves : nullptr;
    F.worklist = g->worklist;
    F.phase_clocks = g->phase_clocks;
    F.done_counter = trace ? nullptr : g->done_counter;   // a trace bypasses the global-top bookkeeping
    F.pub = (trace || g->I > 4096) ? nullptr : g->ts_pub;
    F.gedge = g->gedge; F.gedge_ver = g->gedge_ver;
    gj_status rc;
#define GJ_LAUNCH_TSFAST(NT, MB, TR)                                                         \
    do {                                                                                     \
        if ((rc = opt_in_smem(k_ts_step_fast<NT, MB, TR>, g->fused_smem))) return rc;       \
        k_ts_step_fast<NT, MB, TR><<<g->I, NT, g->fused_smem, st>>>(P, g->groups, F);       \
    } while (0)
    // The step's outcome does not depend on the CTA shape (neighbours are ordered by index), so the
    // trace always runs one instantiation; production picks (threads, resident CTAs per SM) from the
    // number of islands per SM -- fused_threads / fused_mb, set at creation.
    const int nt = g->fused_threads, mb = g->fused_mb;
    if (trace) GJ_LAUNCH_TSFAST(256, 4, true);
    else if (nt > 512) GJ_LAUNCH_TSFAST(1024, 1, false);
    else if (nt > 256) { if (mb >= 3) GJ_LAUNCH_TSFAST(512, 3, false); else GJ_LAUNCH_TSFAST(512, 2, false); }
    else if (nt > 128) { if (mb >= 6) GJ_LAUNCH_TSFAST(256, 6, false); else GJ_LAUNCH_TSFAST(256, 4, false); }
    else GJ_LAUNCH_TSFAST(128, 8, false);
#undef GJ_LAUNCH_TSFAST
    if (F.pub) {
        GJ_LAUNCH_CHECK();
        k_ts_publish<<<1, 256, 0, st>>>(g->ts_pub, g->stride, g->n_vars, g->best, g->best_score, g->gbest, g->gbest_score, g->gver,
                                        g->cur_score, g->ds.stale, g->dirty, g->ts_edge, g->ts_edge_stride, g->gedge, g->gedge_ver);
    }
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}
