// gj_islands_chain.cuh -- LateAcceptance chains: MANY steps of one island per launch
// (included by gj_islands.cu).
//
// A LateAcceptance agent scores ONE neighbour per step (late_acceptance_base.rs:116-141), so
// throughput only comes from running thousands of independent chains, and a kernel launch per
// step would be all overhead.  Here one WARP owns one chain for `n_steps` consecutive steps:
//
//   stage   the chain's solution, value counts, edge lengths (TSP), tabu bitmap + deque and
//           late-score list live in the warp's slice of shared memory for the whole launch
//   step    move from the counter RNG (same generator as every other path; tabu ids by
//           rejection against the bitmap, Mover::select_non_tabu_ids :75-96), delta score
//           (gj_delta.cuh) or -- for moves it does not cover -- apply to a scratch clone + the
//           warp-wide full evaluator; acceptance rule of late_acceptance_base.rs:196-213;
//           accepted: apply in shared memory, refresh the touched edges / counts, push the score
//           to the late list, update_top_individual (agent_base.rs:220-224) straight to HBM
//   finish  write the chain back; the stored score is re-derived from a FULL evaluation in the
//           reference's summation order, so float drift is bounded to one launch
//
// Chains are independent between migrations (the reference's island model); migration and
// update_global_top run between launches (every min(migration_frequency, steps) steps -- the
// reference refreshes the global top after every step of every agent thread, asynchronously; this
// is the declared relaxation).
#pragma once


struct GjChainSmem {
    int32_t* t;         // [n_pad + 8], solution at +4 (sentinels around it)
    int32_t* cnt;       // [cnt_stride]
    int32_t* scratch;   // [n_pad] clone / permutation buffer
    uint32_t* bm;       // [words] full-evaluator bitmap
    uint32_t* tabu;     // [ctabu_words] the chain's tabu state
    double* edge;       // [n + 1] TSP
    double* late;       // [late_size][levels of the model]
};


__device__ __forceinline__ GjChainSmem gj_chain_carve(unsigned char* smem, int n_vars, int words,
                                                      int ctabu_words, int late_size, bool tsp) {
    const size_t n_pad = ((size_t)n_vars + 3) & ~(size_t)3;
    GjChainSmem s;
    size_t o = 0;
    s.t = (int32_t*)(smem + o) + 4; o += (n_pad + 8) * 4;
    s.cnt = (int32_t*)(smem + o); o += (size_t)32 * words * 4;
    s.scratch = (int32_t*)(smem + o); o += (n_pad < 32 ? 32 : n_pad) * 4;
    s.bm = (uint32_t*)(smem + o); o += (size_t)words * 4;
    s.tabu = (uint32_t*)(smem + o); o += (((size_t)ctabu_words + 3) & ~(size_t)3) * 4;
    o = (o + 15) & ~(size_t)15;
    s.edge = (double*)(smem + o);
    if (tsp) o += (((size_t)n_vars + 2) & ~(size_t)1) * 8;
    s.late = (double*)(smem + o);
    return s;
}

// counts of the chain's solution, one warp
template <int KIND>
__device__ __forceinline__ void gj_chain_counts(const GjProblemDev& P, const GjChainSmem& s, int lane) {
    const int cnt_stride = 32 * (P.bm_words + P.desc_words + P.asc_words);
    for (int i = lane; i < cnt_stride; i += 32) s.cnt[i] = 0;
    __syncwarp();
    for (int i = lane; i < P.n_vars; i += 32) {
        const int v = s.t[i];
        atomicAdd(&s.cnt[v - P.val_lo], 1);
        if constexpr (KIND == GJ_NQUEENS) {
            const int col = P.column_id[i];
            atomicAdd(&s.cnt[32 * P.bm_words + (col + v - P.desc_lo)], 1);
            atomicAdd(&s.cnt[32 * (P.bm_words + P.desc_words) + (col - v - P.asc_lo)], 1);
        }
    }
    __syncwarp();
}

// edge lengths of tour positions [lo, hi] (inclusive, clamped to [0, n]), one warp
__device__ __forceinline__ void gj_chain_edges(const GjProblemDev& P, const GjChainSmem& s, int lo, int hi,
                                               int lane) {
    const int n = P.n_vars;
    const size_t L = (size_t)P.n_locations;
    lo = max(lo, 0); hi = min(hi, n);
    for (int i = lo + lane; i <= hi; i += 32)
        s.edge[i] = __ldg(&P.D[(size_t)s.t[i - 1] * L + (size_t)s.t[i]]);     // sentinels: t[-1] = t[n] = 0
    __syncwarp();
}

// FULL evaluation of the staged chain by its warp -> unweighted terms (gj_eval.cuh evaluators,
// reference summation order when exact sums are on)
template <int KIND>
__device__ __forceinline__ void gj_chain_full_eval(const GjProblemDev& P, const int32_t* row, uint32_t* bm,
                                                   int lane, double& r0, double& r1) {
    GjSrcI32 src{row};
    if constexpr (KIND == GJ_NQUEENS) {
        r0 = gj_nqueens_eval_warp(P, src, bm, lane);
        r1 = 0.0;
    } else {
        gj_tsp_eval_warp(P, src, bm, lane, r0, r1);
        r1 = __shfl_sync(GJ_FULL_MASK, r1, 0);
        r0 = __shfl_sync(GJ_FULL_MASK, r0, 0);
    }
    __syncwarp();
}

template <int KIND>
__device__ __forceinline__ void gj_chain_combine(const GjProblemDev& P, double r0, double r1, GjScore& s) {
    s.v[0] = s.v[1] = s.v[2] = 0.0;
    if constexpr (KIND == GJ_NQUEENS) gj_combine_nqueens(P, r0, s.v);
    else gj_combine_tsp(P, true, r0, r1, s.v);
}


// MAXW = warps (chains) per CTA at most.  Two instantiations: 4 (several CTAs per SM, no register cap) and
// kChainWarpsWide = 28 (ONE CTA per SM, <= 72 registers): 4096 chains are 27.7 per SM, and at 96 registers
// only 20 fit -- the narrow variant then runs a second, 38 % full wave that takes as long as the first.
template <int KIND, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, MAXW > 4 ? 1 : 5)
k_la_chains(GjProblemDev P, GjGroups G, GjChainArgs A, size_t per_chain_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ GjMove sh_mv[MAXW];
    constexpr int LV = (KIND == GJ_NQUEENS) ? 1 : 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int island = blockIdx.x * (blockDim.x >> 5) + warp;
    if (island >= A.I) return;                       // whole warps only; no CTA-wide barrier below
    const int n = A.n_vars;
    const int words = P.bm_words + P.desc_words + P.asc_words;
    const GjChainSmem s = gj_chain_carve(smem_raw + (size_t)warp * per_chain_bytes, n, words,
                                         A.ctabu_words_per_island, A.late_size, KIND == GJ_TSP);
    int32_t* cur_row = A.cur + (size_t)island * A.stride;
    int32_t* best_row = A.best + (size_t)island * A.stride;

    // ---- stage ---------------------------------------------------------------------------------
    // update_global_top, adopt half (agent_base.rs:465-489).  The reference tests
    // `global.score < agent_top.score` at the end of EVERY iteration and, while it holds, resets
    // population[0] to the global top (LateAcceptance pushes the score it leaves behind each time).
    // A chain sees the global top published before its launch (`G`); the test of the launch's first
    // iteration is made here, the later ones inside the step loop ("holding" below).
    // gseen[island] == version  <=>  the chain's stored solution IS that version's gbest row.
    int adopted = 0, have_g = 0, cur_is_g = 0, g_ver = 0;
    if (lane == 0 && A.gver) {
        g_ver = *A.gver;
        have_g = g_ver != 0;
        cur_is_g = (have_g && A.gseen[island] == g_ver && A.dirty[island] != 1) ? 1 : 0;
        if (have_g && !cur_is_g) {
            const GjScore g = gj_load_score(A.gbest_score, LV);
            const GjScore mytop = gj_load_score(A.best_score + (size_t)island * GJ_MAX_LEVELS, LV);
            adopted = gj_score_le(mytop, g, LV) ? 0 : 1;                 // global < agent_top
        }
    }
    adopted = __shfl_sync(GJ_FULL_MASK, adopted, 0);
    have_g = __shfl_sync(GJ_FULL_MASK, have_g, 0);
    g_ver = __shfl_sync(GJ_FULL_MASK, g_ver, 0);
    cur_is_g = __shfl_sync(GJ_FULL_MASK, cur_is_g, 0) | adopted;
    const int32_t* src_row = adopted ? A.gbest : cur_row;
    for (int i = lane; i < n; i += 32) s.t[i] = src_row[i];
    if (lane == 0) { s.t[-1] = 0; s.t[n] = 0; }
    uint32_t* tabu_g = A.ctabu ? A.ctabu + (size_t)island * A.ctabu_words_per_island : nullptr;
    if (tabu_g) for (int w = lane; w < A.ctabu_words_per_island; w += 32) s.tabu[w] = tabu_g[w];
    double* late_g = A.late + (size_t)island * A.late_size * GJ_MAX_LEVELS;
    if (A.late) for (int i = lane; i < A.late_size * LV; i += 32) s.late[i] = late_g[(size_t)(i / LV) * GJ_MAX_LEVELS + (i % LV)];
    const bool is_la = A.agent == GJ_AGENT_LATE_ACCEPTANCE;
    int late_head = is_la ? A.late_head[island] : 0, late_len = is_la ? A.late_len[island] : 0;
    double temp[GJ_MAX_LEVELS] = {1.0, 1.0, 1.0};
    if (!is_la)
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) temp[l] = A.sa_temp[(size_t)island * GJ_MAX_LEVELS + l];
    __syncwarp();
    gj_chain_counts<KIND>(P, s, lane);
    if constexpr (KIND == GJ_TSP) gj_chain_edges(P, s, 0, n, lane);
    double raw0, raw1;
    gj_chain_full_eval<KIND>(P, s.t, s.bm, lane, raw0, raw1);
    raw0 = __shfl_sync(GJ_FULL_MASK, raw0, 0);
    GjScore cur = gj_load_score(A.cur_score + (size_t)island * GJ_MAX_LEVELS, LV);
    GjScore top = gj_load_score(A.best_score + (size_t)island * GJ_MAX_LEVELS, LV);
    GjScore gsc = cur;
    if (have_g) gsc = gj_load_score(A.gbest_score, LV);
    if (adopted) {
        if (is_la) {          // LateAcceptance remembers the score it leaves behind (agent_base.rs:467-471)
            late_head = (late_head == 0 ? A.late_size : late_head) - 1;
            if (lane == 0)
                for (int l = 0; l < LV; ++l) s.late[(size_t)late_head * LV + l] = cur.v[l];
            late_len = min(late_len + 1, A.late_size);
            __syncwarp();
        }
        cur = gsc;
    }
    // update_top_individual (agent_base.rs:220-224) runs once per iteration, AFTER the step: a
    // solution that arrived between steps (migrant, global top) is compared with the agent's top
    // only after the next step had its chance to replace it.
    bool top_check_pending = A.dirty[island] || adopted || (cur_is_g && have_g && !gj_score_le(top, gsc, LV));
    int accepted_total = 0;
    bool cur_from_step = false;                      // cur was produced by an accepted step of this launch
    bool top_from_step = false;                      // ... and so was the agent's top individual

    for (int it = 0; it < A.n_steps; ++it) {
        const uint64_t step = A.step0 + (uint64_t)it;
        // ---- generate (every lane computes the same move: no divergence, no broadcast) ----------
        GjMoverParams M = A.M;
        {
            // the move is parked in shared memory: 18 registers the rest of the step does not have (72-register
            // cap of the wide shape) would otherwise sit in local memory, which is L2 under its carve-out
            const GjMove gen = gj_generate_move(P, G, M, A.seed, (uint32_t)(A.island_base + island), step, 0u,
                                                tabu_g ? s.tabu : nullptr, A.ctabu_off);
            if (lane == 0) sh_mv[warp] = gen;
            __syncwarp();
        }
        const GjMove& m = sh_mv[warp];
        // ---- score -----------------------------------------------------------------------------
        int d_uniq = 0; double d_dist = 0.0;
        bool ok;
        if constexpr (KIND == GJ_NQUEENS) {
            ok = gj_nqueens_move_delta(P, G, m, A.noop != 0, s.t, s.cnt, d_uniq);
        } else {
            GjTspBase B{s.t, n, P.D, (size_t)P.n_locations, s.edge, true};
            ok = gj_tsp_move_delta(P, G, m, A.noop != 0, A.symmetric != 0, B, s.cnt, d_uniq, d_dist);
        }
        double n0 = raw0 - (double)d_uniq, n1 = raw1 + d_dist;
        if (!ok) {
            // scratch clone + full evaluator (a move the delta evaluator does not cover)
            for (int i = lane; i < n; i += 32) s.scratch[i] = s.t[i];
            __syncwarp();
            const GjMove ms = sh_mv[warp];
            gj_apply_move(P, ms, G, true, A.noop != 0, lane, 32,
                          [&](int id) { return s.t[id]; }, [&](int id, int v) { s.scratch[id] = v; });
            __syncwarp();
            gj_chain_full_eval<KIND>(P, s.scratch, s.bm, lane, n0, n1);
            n0 = __shfl_sync(GJ_FULL_MASK, n0, 0);
        }
        GjScore sc;
        gj_chain_combine<KIND>(P, n0, n1, sc);
        gj_score_round(sc, P);                          // agent_base.rs:311-314
        // ---- late acceptance (late_acceptance_base.rs:196-213) -----------------------------------
        bool accept;
        if (is_la) {
            GjScore late_native = cur;
            if (late_len > 0) late_native = gj_load_score(s.late + (size_t)gj_wrap_once(late_head + late_len - 1, A.late_size) * LV, LV);
            accept = gj_score_le(sc, late_native, LV) || gj_score_le(sc, cur, LV);
        } else {
            // SimulatedAnnealing (simulated_annealing_base.rs:198-233); every lane evaluates the same rule
            const double u = gj_accept_uniform(A.seed, (uint32_t)(A.island_base + island), step);
            double proba;
            accept = gj_sa_accept(sc, cur, LV, temp, A.sa, u, &proba);
            if (A.trace_aux && lane == 0) {
                double* o = A.trace_aux + (size_t)island * 5;
                o[0] = u; o[1] = proba; o[2] = temp[0]; o[3] = temp[1]; o[4] = temp[2];
            }
        }
        if (A.trace_moves) {
            if (lane == 0) {
                A.trace_moves[(size_t)it * A.I + island] = m;
                for (int l = 0; l < LV; ++l) A.trace_scores[((size_t)it * A.I + island) * LV + l] = sc.v[l];
                A.trace_accept[(size_t)it * A.I + island] = accept ? 1 : 0;
            }
        }
        // standing on the global top while it still beats the agent's own top (update_global_top)
        const bool holding = cur_is_g && have_g && !gj_score_le(top, gsc, LV);
        const bool bounced = accept && holding && !gj_score_le(sc, cur, LV);
        if (bounced) {
            // A WORSE neighbour accepted on the global top.  The reference moves to it
            // (late_acceptance_base.rs:207-211), update_top_individual may record it as agent_top, and
            // update_global_top of the same iteration puts population[0] back on the global top, pushing
            // the neighbour's score once more (agent_base.rs:465-471).  Net effect, reproduced without
            // moving: the solution stays, the late list grows by two, agent_top may take the neighbour.
            accepted_total += 1;
            if (is_la) {
                for (int rep = 0; rep < 2; ++rep) {
                    late_head = (late_head == 0 ? A.late_size : late_head) - 1;
                    if (lane == 0)
                        for (int l = 0; l < LV; ++l) s.late[(size_t)late_head * LV + l] = sc.v[l];
                    late_len = min(late_len + 1, A.late_size);
                }
            }
            if (gj_score_le(sc, top, LV)) {
                top = sc;
                top_from_step = true;
                if (!ok) {
                    for (int i = lane; i < n; i += 32) best_row[i] = s.scratch[i];
                } else {
                    for (int i = lane; i < n; i += 32) best_row[i] = s.t[i];
                    __syncwarp();
                            __syncwarp();
                    const GjMove ms = sh_mv[warp];
                    gj_apply_move(P, ms, G, true, A.noop != 0, lane, 32,
                                  [&](int id) { return s.t[id]; }, [&](int id, int v) { best_row[id] = v; });
                }
            }
            __syncwarp();
        } else if (accept) {
            cur_is_g = 0;
            top_check_pending = true;
            // ---- apply in shared memory ------------------------------------------------------------
            if (!ok) {
                for (int i = lane; i < n; i += 32) s.t[i] = s.scratch[i];
                __syncwarp();
                gj_chain_counts<KIND>(P, s, lane);
                if constexpr (KIND == GJ_TSP) gj_chain_edges(P, s, 0, n, lane);
            } else if (m.kind != GJ_MOVE_NULL) {
                    __syncwarp();
                const GjMove ms = sh_mv[warp];
                if (ms.kind <= 3) {
                    // small move: <= 16 (column, value) pairs, later pairs win
                    if (lane == 0) {
                        const int32_t* g = G.ids + G.offsets[ms.group];
                        // one (column, value) pair: counts, then the value; the column is remembered for the
                        // edge refresh below
                        auto put = [&](int i, int c, int vraw) {
                            const int v = gj_fix_column(P, c, vraw), o = s.t[c];
                            if (v != o) {
                                s.cnt[o - P.val_lo] -= 1; s.cnt[v - P.val_lo] += 1;
                                if constexpr (KIND == GJ_NQUEENS) {
                                    const int col = P.column_id[c];
                                    const int od = 32 * P.bm_words - P.desc_lo, oa = 32 * (P.bm_words + P.desc_words) - P.asc_lo;
                                    s.cnt[od + col + o] -= 1; s.cnt[od + col + v] += 1;
                                    s.cnt[oa + col - o] -= 1; s.cnt[oa + col - v] += 1;
                                }
                                s.t[c] = v;
                            }
                            s.scratch[1 + i] = c;
                        };
                        if (ms.kind == 1 && ms.k == 2) {
                            // a swap of two stops (the common move): both values are read before the first
                            // write, as the generic expansion does -- without its scratch arrays, which live
                            // in local memory (L2, with this kernel's shared-memory carve-out)
                            const int c0 = g[sh_mv[warp].a[0]], c1 = g[sh_mv[warp].a[1]];
                            const int v0 = s.t[c1], v1 = s.t[c0];
                            s.scratch[0] = 2;
                            put(0, c0, v0);
                            put(1, c1, v1);
                        } else {
                            int cols[GJ_MOVE_MAXPAIRS], vals[GJ_MOVE_MAXPAIRS];
                            const int np = gj_small_move_pairs(ms, g, true, A.noop != 0,
                                                               [&](int id) { return s.t[id]; }, cols, vals);
                            s.scratch[0] = np;
                            for (int i = 0; i < np; ++i) put(i, cols[i], vals[i]);
                        }
                    }
                    __syncwarp();
                    if constexpr (KIND == GJ_TSP) {
                        // the two edges around every changed stop
                        const int np = s.scratch[0];
                        if (lane < np) {
                            const int c = s.scratch[1 + lane];
                            const size_t L = (size_t)P.n_locations;
                            s.edge[c] = __ldg(&P.D[(size_t)s.t[c - 1] * L + (size_t)s.t[c]]);
                            s.edge[c + 1] = __ldg(&P.D[(size_t)s.t[c] * L + (size_t)s.t[c + 1]]);
                        }
                        __syncwarp();
                    }
                } else {
                    // segment move on consecutive columns (only those reach the delta path)
                    const int4 gi = G.info[ms.group];
                    int plo, phi;
                    gj_segment_bounds(ms, plo, phi);
                    const int a = gi.x + plo, len = phi - plo + 1;
                    for (int t0 = lane; t0 < len; t0 += 32) s.scratch[t0] = s.t[a + t0];
                    __syncwarp();
                    for (int t0 = lane; t0 < len; t0 += 32)
                        s.t[a + t0] = s.scratch[gj_segment_src_slot(ms, true, t0, len)];
                    __syncwarp();
                    if constexpr (KIND == GJ_TSP) gj_chain_edges(P, s, a, a + len, lane);
                    if constexpr (KIND == GJ_NQUEENS) gj_chain_counts<KIND>(P, s, lane);
                }
                __syncwarp();
            }
            raw0 = n0; raw1 = n1;
            cur = sc;
            cur_from_step = true;
            accepted_total += 1;
            if (is_la) {
                // push_front; pop_back when longer than late_acceptance_size
                late_head = (late_head == 0 ? A.late_size : late_head) - 1;
                if (lane == 0)
                    for (int l = 0; l < LV; ++l) s.late[(size_t)late_head * LV + l] = sc.v[l];
                late_len = min(late_len + 1, A.late_size);
            }
            __syncwarp();
        }
        // update_top_individual (agent_base.rs:220-224): population[0] <= agent_top -> replace
        if (top_check_pending && !bounced) {
            top_check_pending = false;
            if (gj_score_le(cur, top, LV)) {
                top = cur;
                top_from_step = cur_from_step;
                for (int i = lane; i < n; i += 32) best_row[i] = s.t[i];
            }
            __syncwarp();
        }
        // ---- tabu deque (mover.rs:75-96): every id the move selected enters, the oldest leave ------
        if (tabu_g && m.kind != GJ_MOVE_NULL && lane == 0) {
            int sel[GJ_MOVE_MAXK];
            const int cntsel = gj_move_selected(m, sel);
            const int glen = G.offsets[m.group + 1] - G.offsets[m.group];
            const int W = (glen + 31) >> 5;
            const int T = A.tabu_size[m.group];
            uint32_t* bits = s.tabu + A.ctabu_off[m.group];
            int32_t* ring = (int32_t*)(bits + W + 1);
            int head = ring[T], fill = ring[T + 1];
#pragma unroll
            for (int i = 0; i < GJ_MOVE_MAXK; ++i) {
                if (i < cntsel) {
                    const int pos = sel[i];
                    if (!((bits[pos >> 5] >> (pos & 31)) & 1u)) {
                        if (fill == T) {
                            const int old = ring[head];
                            bits[old >> 5] &= ~(1u << (old & 31));
                        } else {
                            fill += 1;
                        }
                        ring[head] = pos;
                        head = (head + 1 == T) ? 0 : head + 1;
                        bits[pos >> 5] |= 1u << (pos & 31);
                    }
                }
            }
            ring[T] = head; ring[T + 1] = fill;
        }
        __syncwarp();
    }

    // ---- finish: write the chain back ---------------------------------------------------------------
    for (int i = lane; i < n; i += 32) cur_row[i] = s.t[i];
    if (tabu_g) for (int w = lane; w < A.ctabu_words_per_island; w += 32) tabu_g[w] = s.tabu[w];
    if (A.late) for (int i = lane; i < A.late_size * LV; i += 32) late_g[(size_t)(i / LV) * GJ_MAX_LEVELS + (i % LV)] = s.late[i];
    __syncwarp();
    // stored scores are FULL evaluations of the stored vectors (reference summation order): float
    // drift of the delta sums never outlives a launch
    if (cur_from_step) {
        double r0, r1;
        gj_chain_full_eval<KIND>(P, s.t, s.bm, lane, r0, r1);
        gj_chain_combine<KIND>(P, r0, r1, cur);
        gj_score_round(cur, P);
    }
    if (top_from_step) {
        __syncwarp();
        double r0, r1;
        gj_chain_full_eval<KIND>(P, best_row, s.bm, lane, r0, r1);
        gj_chain_combine<KIND>(P, r0, r1, top);
        gj_score_round(top, P);
    }
    if (lane == 0) {
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) {
            A.cur_score[(size_t)island * GJ_MAX_LEVELS + l] = (l < LV) ? cur.v[l] : 0.0;
            A.best_score[(size_t)island * GJ_MAX_LEVELS + l] = (l < LV) ? top.v[l] : 0.0;
        }
        if (is_la) { A.late_head[island] = late_head; A.late_len[island] = late_len; }
        else for (int l = 0; l < GJ_MAX_LEVELS; ++l) A.sa_temp[(size_t)island * GJ_MAX_LEVELS + l] = temp[l];
        A.dirty[island] = 0;
        if (A.gseen) A.gseen[island] = cur_is_g ? g_ver : 0;
        atomicAdd(&A.counters[0], (unsigned long long)A.n_steps);
        if (island == 0) atomicAdd(&A.counters[1], (unsigned long long)A.n_steps);
        if (accepted_total) atomicAdd(&A.counters[2], (unsigned long long)accepted_total);
    }
}
