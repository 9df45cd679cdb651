// gj_islands.cu -- device-resident agents ("islands"): the reference's Agent::solve loop
// (greyjack/src/agents/base/agent_base.rs:124-188) for TabuSearch and LateAcceptance,
// with move generation, scoring, selection, ring migration and the shared global best all
// on the GPU.  One launch handles every island of the group:
//
//   k_gen_moves      Mover::do_move x neighbours_count            (mover.rs:98-128)
//   k_score_moves_*  request_score_incremental on base + move     (oop_score_requester.rs:443-463)
//   k_select         build_updated_population_incremental + update_top_individual
//                    (tabu_search_base.rs:157-188, late_acceptance_base.rs:188-241,
//                     agent_base.rs:220-224)
//   k_migrate_*      send_updates / receive_updates               (agent_base.rs:322-444)
//   k_global_*       update_global_top                            (agent_base.rs:446-490)
//
// The GeneticAlgorithm agent lives in gj_islands_ga.cu.
#include <algorithm>
#include <cmath>
#include <memory>
#include <cstdio>
#include <cstdlib>

#include "gj_islands_dev.cuh"

static constexpr int kWarps = 4;

// ---- generation -------------------------------------------------------------------------------
__global__ void k_gen_moves(GjProblemDev P, GjGroups G, GjMoverParams M, uint64_t seed,
                            uint64_t step, int I, int K, int island_base,
                            const uint32_t* __restrict__ tabu_bits, int tabu_words_per_island,
                            const int32_t* __restrict__ tabu_word_off, GjMove* __restrict__ moves) {
    const int64_t total = (int64_t)I * K;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total;
         j += (int64_t)gridDim.x * blockDim.x) {
        const int island = (int)(j / K), cand = (int)(j % K);
        const uint32_t* bits = tabu_bits ? tabu_bits + (size_t)island * tabu_words_per_island : nullptr;
        moves[j] = gj_generate_move(P, G, M, seed, (uint32_t)(island_base + island), step,
                                    (uint32_t)cand, bits, tabu_word_off);
    }
}

// ---- scoring of base + move -----------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(kWarps * 32)
k_score_moves_warp(GjProblemDev P, GjGroups G, const int32_t* __restrict__ cur, int stride,
                   const GjMove* __restrict__ moves, int K, int64_t total, int incremental,
                   int noop, int isc, double* __restrict__ scores) {
    extern __shared__ uint32_t smem_u32[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int words = P.bm_words + P.desc_words + P.asc_words;
    const int per_warp = words + P.n_vars;
    uint32_t* bm = smem_u32 + warp * per_warp;
    int32_t* cand = (int32_t*)(bm + words);
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;      // 1..kWarps warps per CTA
    if (j >= total) return;
    const int32_t* base = cur + (size_t)(j / K) * stride;
    for (int i = lane; i < P.n_vars; i += 32) cand[i] = base[i];
    __syncwarp();
    const GjMove m = moves[j];
    gj_apply_move(P, m, G, incremental != 0, noop != 0, lane, 32,
                  [&](int id) { return base[id]; }, [&](int id, int v) { cand[id] = v; });
    __syncwarp();
    GjSrcI32 src{cand};
    GjScore s;
    if constexpr (KIND == GJ_NQUEENS) {
        gj_combine_nqueens(P, gj_nqueens_eval_warp(P, src, bm, lane), s.v);
    } else {
        double dup, dist;
        gj_tsp_eval_warp(P, src, bm, lane, dup, dist);
        gj_combine_tsp(P, isc != 0, dup, dist, s.v);
    }
    if (lane == 0) {
        gj_score_round(s, P);           // agent_base.rs:311-314
        for (int l = 0; l < P.levels; ++l) scores[j * P.levels + l] = s.v[l];
    }
}

__global__ void __launch_bounds__(kVrpWarps * 32)
k_score_moves_vrp(GjProblemDev P, GjGroups G, const int32_t* __restrict__ cur, int stride,
                  const GjMove* __restrict__ moves, int K, int64_t total, int incremental,
                  int noop, int isc, double* __restrict__ scores) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = P.n_entities;
    GjVrpSmem s = gj_vrp_carve(smem_raw, n, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
    const int64_t j = blockIdx.x;
    const int32_t* base = cur + (size_t)(j / K) * stride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int2 pr = *reinterpret_cast<const int2*>(base + 2 * i);
        s.veh[i] = (uint16_t)pr.x;
        s.cust[i] = pr.y;
    }
    __syncthreads();
    const GjMove m = moves[j];
    gj_apply_move(P, m, G, incremental != 0, noop != 0, threadIdx.x, blockDim.x,
                  [&](int id) { return base[id]; },
                  [&](int id, int v) { if (id & 1) s.cust[id >> 1] = v; else s.veh[id >> 1] = (uint16_t)v; });
    __syncthreads();
    const int tw_mode = isc ? (P.kind == GJ_VRP_SERVICE ? GJ_TW_ISC_SERVICE : GJ_TW_ISC_FILE) : GJ_TW_PSC;
    double dup1000 = 0, cap = 0, dist = 0, late = 0;
    gj_vrp_eval_cta(P, s, tw_mode, dup1000, cap, dist, late);
    if (threadIdx.x == 0) {
        GjScore sc;
        gj_combine_vrp(P, isc != 0, dup1000, cap, dist, late, sc.v);
        gj_score_round(sc, P);
        for (int l = 0; l < 3; ++l) scores[j * 3 + l] = sc.v[l];
    }
}

// ---- delta scoring (GJ_SCORING_DELTA) ------------------------------------------------------------
// One THREAD per neighbour: the move is generated in registers from the counter RNG (never
// stored), evaluated against the island's cached state (gj_delta.cuh) and its rounded score is
// written out.  Neighbours whose move the delta evaluator does not cover are queued for
// k_score_fallback_warp (full evaluator, one warp each).
struct GjStepCtx {
    uint64_t seed, step;
    int I, K, island_base, noop, stride, symmetric;
    const uint32_t* tabu_bits; int tabu_words_per_island; const int32_t* tabu_word_off;
};

template <int KIND>
__global__ void __launch_bounds__(256)
k_score_delta(GjProblemDev P, GjGroups G, GjMoverParams M, GjStepCtx C,
              const int32_t* __restrict__ cur, GjDeltaState S, double* __restrict__ scores,
              int* __restrict__ worklist, int* __restrict__ work_count, GjMove* __restrict__ moves_out) {
    const int64_t total = (int64_t)C.I * C.K;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    const int island = (int)(j / C.K), cand = (int)(j - (int64_t)island * C.K);
    const uint32_t* bits = C.tabu_bits ? C.tabu_bits + (size_t)island * C.tabu_words_per_island : nullptr;
    const GjMove m = gj_generate_move(P, G, M, C.seed, (uint32_t)(C.island_base + island), C.step,
                                      (uint32_t)cand, bits, C.tabu_word_off);
    if (moves_out) moves_out[j] = m;
    const int32_t* row = cur + (size_t)island * C.stride;
    const int32_t* cnt = S.cnt + (size_t)island * S.cnt_stride;
    const double* raw = S.raw + (size_t)island * GJ_MAX_LEVELS;
    GjScore s;
    bool ok;
    if constexpr (KIND == GJ_NQUEENS) {
        int d_uniq;
        ok = gj_nqueens_move_delta(P, G, m, C.noop != 0, row, cnt, d_uniq);
        // (N - |rows|) + (N - |desc|) + (N - |asc|): integers, exact in f64
        gj_combine_nqueens(P, raw[0] - (double)d_uniq, s.v);
    } else {
        GjTspBase B{row, P.n_vars, P.D, (size_t)P.n_locations, nullptr, false};
        int d_uniq; double d_dist;
        ok = gj_tsp_move_delta(P, G, m, C.noop != 0, C.symmetric != 0, B, cnt, d_uniq, d_dist);
        gj_combine_tsp(P, true, raw[0] - (double)d_uniq, raw[1] + d_dist, s.v);
    }
    if (!ok) {
        worklist[atomicAdd(work_count, 1)] = (int)j;
        return;
    }
    gj_score_round(s, P);           // agent_base.rs:311-314
    for (int l = 0; l < P.levels; ++l) scores[j * P.levels + l] = s.v[l];
}

// Full evaluation of the queued neighbours: persistent warps walk the worklist.
template <int KIND>
__global__ void __launch_bounds__(kWarps * 32)
k_score_fallback_warp(GjProblemDev P, GjGroups G, GjMoverParams M, GjStepCtx C,
                      const int32_t* __restrict__ cur, const int* __restrict__ worklist,
                      const int* __restrict__ work_count, double* __restrict__ scores) {
    extern __shared__ uint32_t smem_u32[];
    __shared__ GjMove sh_mv[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int words = P.bm_words + P.desc_words + P.asc_words;
    const int per_warp = words + P.n_vars;
    uint32_t* bm = smem_u32 + warp * per_warp;
    int32_t* cand = (int32_t*)(bm + words);
    const int n_work = *work_count;
    const int cta_warps = blockDim.x >> 5;
    for (int w = blockIdx.x * cta_warps + warp; w < n_work; w += gridDim.x * cta_warps) {
        const int j = worklist[w];
        const int island = j / C.K, c = j - island * C.K;
        const int32_t* base = cur + (size_t)island * C.stride;
        if (lane == 0) {
            const uint32_t* bits = C.tabu_bits ? C.tabu_bits + (size_t)island * C.tabu_words_per_island : nullptr;
            sh_mv[warp] = gj_generate_move(P, G, M, C.seed, (uint32_t)(C.island_base + island), C.step,
                                           (uint32_t)c, bits, C.tabu_word_off);
        }
        for (int i = lane; i < P.n_vars; i += 32) cand[i] = base[i];
        __syncwarp();
        const GjMove m = sh_mv[warp];
        gj_apply_move(P, m, G, true, C.noop != 0, lane, 32,
                      [&](int id) { return base[id]; }, [&](int id, int v) { cand[id] = v; });
        __syncwarp();
        GjSrcI32 src{cand};
        GjScore s;
        if constexpr (KIND == GJ_NQUEENS) {
            gj_combine_nqueens(P, gj_nqueens_eval_warp(P, src, bm, lane), s.v);
        } else {
            double dup, dist;
            gj_tsp_eval_warp(P, src, bm, lane, dup, dist);
            gj_combine_tsp(P, true, dup, dist, s.v);
        }
        if (lane == 0) {
            gj_score_round(s, P);
            for (int l = 0; l < P.levels; ++l) scores[(size_t)j * P.levels + l] = s.v[l];
        }
        __syncwarp();
    }
}

// ---- VRP delta scoring (gj_vrp_delta.cuh) -------------------------------------------------------
// one thread per neighbour: generate the move, re-walk the routes it touches
__global__ void __launch_bounds__(128)
k_score_delta_vrp(GjProblemDev P, GjGroups G, GjMoverParams M, GjStepCtx C, const int32_t* __restrict__ cur,
                  GjDeltaState S, GjVrpState V, double* __restrict__ scores, int* __restrict__ worklist,
                  int* __restrict__ work_count, GjMove* __restrict__ moves_out) {
    const int64_t total = (int64_t)C.I * C.K;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    const int island = (int)(j / C.K), cand = (int)(j - (int64_t)island * C.K);
    const uint32_t* bits = C.tabu_bits ? C.tabu_bits + (size_t)island * C.tabu_words_per_island : nullptr;
    const GjMove m = gj_generate_move(P, G, M, C.seed, (uint32_t)(C.island_base + island), C.step,
                                      (uint32_t)cand, bits, C.tabu_word_off);
    if (moves_out) moves_out[j] = m;
    const int n = P.n_entities, K = P.n_vehicles;
    double dup1000, cap, dist, late;
    const bool ok = gj_vrp_move_delta(P, G, m, C.noop != 0, gj_vrp_tw_mode(P), cur + (size_t)island * C.stride,
                                      V.bucket + (size_t)island * n, V.bstop + (size_t)island * n,
                                      V.start + (size_t)island * (K + 1), V.rdist + (size_t)island * K,
                                      V.rload + (size_t)island * K, V.rlate + (size_t)island * K,
                                      V.tot + (size_t)island * 4, S.cnt + (size_t)island * S.cnt_stride,
                                      dup1000, cap, dist, late);
    if (!ok) {
        worklist[atomicAdd(work_count, 1)] = (int)j;
        return;
    }
    GjScore s;
    gj_combine_vrp(P, true, dup1000, cap, dist, late, s.v);
    gj_score_round(s, P);
    for (int l = 0; l < 3; ++l) scores[j * 3 + l] = s.v[l];
}

// the queued neighbours (segment moves): full evaluator, one CTA each, persistent over the worklist
__global__ void __launch_bounds__(kVrpWarps * 32)
k_score_fallback_vrp(GjProblemDev P, GjGroups G, GjMoverParams M, GjStepCtx C, const int32_t* __restrict__ cur,
                     const int* __restrict__ worklist, const int* __restrict__ work_count,
                     double* __restrict__ scores) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ GjMove sh_mv;
    const int n = P.n_entities;
    GjVrpSmem s = gj_vrp_carve(smem_raw, n, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
    const int n_work = *work_count;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int j = worklist[w];
        const int island = j / C.K, c = j - island * C.K;
        const int32_t* base = cur + (size_t)island * C.stride;
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t* bits = C.tabu_bits ? C.tabu_bits + (size_t)island * C.tabu_words_per_island : nullptr;
            sh_mv = gj_generate_move(P, G, M, C.seed, (uint32_t)(C.island_base + island), C.step, (uint32_t)c,
                                     bits, C.tabu_word_off);
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int2 pr = *reinterpret_cast<const int2*>(base + 2 * i);
            s.veh[i] = (uint16_t)pr.x;
            s.cust[i] = pr.y;
        }
        __syncthreads();
        const GjMove m = sh_mv;
        gj_apply_move(P, m, G, true, C.noop != 0, threadIdx.x, blockDim.x,
                      [&](int id) { return base[id]; },
                      [&](int id, int v) { if (id & 1) s.cust[id >> 1] = v; else s.veh[id >> 1] = (uint16_t)v; });
        __syncthreads();
        double dup1000 = 0, cap = 0, dist = 0, late = 0;
        gj_vrp_eval_cta(P, s, gj_vrp_tw_mode(P), dup1000, cap, dist, late);
        if (threadIdx.x == 0) {
            GjScore sc;
            gj_combine_vrp(P, true, dup1000, cap, dist, late, sc.v);
            gj_score_round(sc, P);
            for (int l = 0; l < 3; ++l) scores[(size_t)j * 3 + l] = sc.v[l];
        }
    }
}

// Rebuilds the cached state of every island whose current solution changed: value counts, and the
// FULL evaluation (reference summation order) of the solution -> raw terms; after an accepted
// neighbour (stale == 2) the stored score is replaced by that full evaluation, rounded, so a
// current score is always exactly what the reference's scorer returns for the stored vector.
// Then (update_top) update_top_individual, agent_base.rs:220-224.  One CTA per island.
template <int KIND>
__global__ void __launch_bounds__(1024)
k_refresh(GjProblemDev P, int stride, const int32_t* __restrict__ cur, double* cur_score,
          GjDeltaState S, int update_top, int32_t* best, double* best_score, int* dirty) {
    __shared__ int sh_i[32];
    __shared__ double sh_d[32];
    const int island = blockIdx.x;
    const int tid = threadIdx.x;
    const int32_t* row = cur + (size_t)island * stride;
    const int why = S.stale[island];
    if (why) {
        // value counts, then the FULL evaluation by the whole CTA: distinct count from the counts;
        // TSP edges gathered in parallel (tree: CTA-wide sum; exact: staged in HBM scratch and
        // folded by one thread strictly in the reference's order, tsp ISC :76-80)
        int32_t* cnt = S.cnt + (size_t)island * S.cnt_stride;
        for (int i = tid; i < S.cnt_stride; i += blockDim.x) cnt[i] = 0;
        __syncthreads();
        const int n = P.n_vars;
        for (int i = tid; i < n; i += blockDim.x) {
            const int v = row[i];
            atomicAdd(&cnt[v - P.val_lo], 1);
            if constexpr (KIND == GJ_NQUEENS) {
                const int col = P.column_id[i];
                atomicAdd(&cnt[32 * P.bm_words + (col + v - P.desc_lo)], 1);
                atomicAdd(&cnt[32 * (P.bm_words + P.desc_words) + (col - v - P.asc_lo)], 1);
            }
        }
        __syncthreads();
        int u = 0;
        for (int k = tid; k < S.cnt_stride; k += blockDim.x) u += (cnt[k] > 0) ? 1 : 0;
        const int uniq = gj_block_sum(u, sh_i);
        double r0, r1 = 0.0;
        if constexpr (KIND == GJ_NQUEENS) {
            r0 = (double)(3 * n - uniq);
        } else {
            r0 = (double)(n - uniq);
            const size_t L = (size_t)P.n_locations;
            double* edge = S.edge + (size_t)island * (size_t)(n + 1);
            double acc = 0.0;
            for (int i = tid; i <= n; i += blockDim.x) {
                const int a = (i == 0) ? 0 : row[i - 1];
                const int b = (i == n) ? 0 : row[i];
                const double d = __ldg(&P.D[(size_t)a * L + (size_t)b]);
                if (P.exact_sums) edge[i] = d; else acc += d;
            }
            if (!P.exact_sums) {
                r1 = gj_block_sum(acc, sh_d);
            } else {
                __syncthreads();
                if (tid == 0) {
                    double fold = 0.0;
#pragma unroll 8
                    for (int i = 1; i < n; ++i) fold = fold + edge[i];
                    double sample_distance = 0.0;
                    sample_distance += edge[0];
                    sample_distance += edge[n];
                    sample_distance += fold;
                    sh_d[0] = sample_distance;
                }
                __syncthreads();
                r1 = sh_d[0];
            }
        }
        if (tid == 0) {
            double* raw = S.raw + (size_t)island * GJ_MAX_LEVELS;
            raw[0] = r0; raw[1] = r1; raw[2] = 0.0;
            if (why == 2) {
                GjScore s;
                s.v[0] = s.v[1] = s.v[2] = 0.0;
                if constexpr (KIND == GJ_NQUEENS) gj_combine_nqueens(P, r0, s.v);
                else gj_combine_tsp(P, true, r0, r1, s.v);
                gj_score_round(s, P);
                for (int l = 0; l < P.levels; ++l) cur_score[(size_t)island * GJ_MAX_LEVELS + l] = s.v[l];
            }
            S.stale[island] = 0;
        }
        __syncthreads();
    }
    if (update_top) gj_update_top(island, P.levels, stride, P.n_vars, cur, cur_score, best, best_score, dirty);
}

// Rebuilds the VRP base state of every island whose solution changed: one FULL evaluation by the
// CTA (the same evaluator that scores plain candidates) whose by-products -- bucketed stops, route
// boundaries, per-route distance / demand / lateness, customer counts -- are kept in HBM.  After an
// accepted neighbour (stale == 2) the stored score is re-derived from that evaluation.  Then
// update_top_individual.
__global__ void __launch_bounds__(kVrpWarps * 32)
k_vrp_state(GjProblemDev P, int stride, const int32_t* __restrict__ cur, double* cur_score, GjDeltaState S,
            GjVrpState V, int update_top, int32_t* best, double* best_score, int* dirty) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int island = blockIdx.x;
    const int n = P.n_entities, K = P.n_vehicles;
    const int why = S.stale[island];
    if (why) {
        GjVrpSmem s = gj_vrp_carve(smem_raw, n, K, P.bm_words, kVrpWarps, !P.time_windowed);
        const int32_t* row = cur + (size_t)island * stride;
        int32_t* cnt = S.cnt + (size_t)island * S.cnt_stride;
        for (int i = threadIdx.x; i < S.cnt_stride; i += blockDim.x) cnt[i] = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int2 pr = *reinterpret_cast<const int2*>(row + 2 * i);
            s.veh[i] = (uint16_t)pr.x;
            s.cust[i] = pr.y;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&cnt[s.cust[i] - P.val_lo], 1);
        GjVrpOut out{V.bstop + (size_t)island * n, V.rload + (size_t)island * K, V.rlate + (size_t)island * K};
        double dup1000 = 0, cap = 0, dist = 0, late = 0;
        gj_vrp_eval_cta(P, s, gj_vrp_tw_mode(P), dup1000, cap, dist, late, &out);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) V.bucket[(size_t)island * n + i] = s.bucket[i];
        for (int v = threadIdx.x; v <= K; v += blockDim.x) V.start[(size_t)island * (K + 1) + v] = s.start[v];
        for (int v = threadIdx.x; v < K; v += blockDim.x) V.rdist[(size_t)island * K + v] = s.vdist[v];
        if (threadIdx.x == 0) {
            unsigned long long* tot = V.tot + (size_t)island * 4;
            tot[0] = (unsigned long long)llrint(dup1000 / 1000.0);
            tot[1] = s.acc[0];
            tot[2] = s.acc[1];
            if (why == 2) {
                GjScore sc;
                gj_combine_vrp(P, true, dup1000, cap, dist, late, sc.v);
                gj_score_round(sc, P);
                for (int l = 0; l < 3; ++l) cur_score[(size_t)island * GJ_MAX_LEVELS + l] = sc.v[l];
            }
            S.stale[island] = 0;
        }
        __syncthreads();
    }
    if (update_top) gj_update_top(island, P.levels, stride, P.n_vars, cur, cur_score, best, best_score, dirty);
}

// ---- selection ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_select(GjProblemDev P, GjGroups G, GjSelectArgs A) {
    extern __shared__ int32_t smem_row[];          // [n_vars] copy of the base for in-place apply
    __shared__ GjScore sh_score[32];
    __shared__ int sh_idx[32];
    __shared__ int sh_accept, sh_best;
    __shared__ int sh_scan[1024];
    const int island = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int K = A.K, levels = A.levels;
    const double* cs = A.cand_scores + (size_t)island * K * levels;
    const uint32_t* bits_island = A.tabu_bits ? A.tabu_bits + (size_t)island * A.tabu_words_per_island : nullptr;
    auto load_move = [&](int j) -> GjMove {
        if (A.moves) return A.moves[(size_t)island * K + j];
        return gj_generate_move(P, G, A.M, A.seed, (uint32_t)(A.island_base + island), A.step,
                                (uint32_t)j, bits_island, A.tabu_word_off);
    };

    // first minimum by Ord::cmp (tabu_search_base.rs:166-171: min_by keeps the first)
    GjScore mine; int mine_idx = -1;
    for (int l = 0; l < GJ_MAX_LEVELS; ++l) mine.v[l] = 0.0;
    for (int j = tid; j < K; j += blockDim.x) {
        GjScore s = gj_load_score(cs + (size_t)j * levels, levels);
        if (mine_idx < 0 || gj_score_cmp(s, mine, levels) < 0) { mine = s; mine_idx = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        GjScore other; int oidx = __shfl_xor_sync(GJ_FULL_MASK, mine_idx, o);
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) other.v[l] = __shfl_xor_sync(GJ_FULL_MASK, mine.v[l], o);
        if (oidx >= 0) {
            int c = (mine_idx < 0) ? 1 : gj_score_cmp(mine, other, levels);
            if (c > 0 || (c == 0 && oidx < mine_idx)) { mine = other; mine_idx = oidx; }
        }
    }
    if (lane == 0) { sh_score[warp] = mine; sh_idx[warp] = mine_idx; }
    __syncthreads();
    if (tid == 0) {
        GjScore b = sh_score[0]; int bi = sh_idx[0];
        for (int w = 1; w < nwarps; ++w) {
            if (sh_idx[w] < 0) continue;
            int c = (bi < 0) ? 1 : gj_score_cmp(b, sh_score[w], levels);
            if (c > 0 || (c == 0 && sh_idx[w] < bi)) { b = sh_score[w]; bi = sh_idx[w]; }
        }
        GjScore cur = gj_load_score(A.cur_score + (size_t)island * GJ_MAX_LEVELS, levels);
        bool accept;
        if (A.agent == GJ_AGENT_TABU_SEARCH) {
            accept = gj_score_le(b, cur, levels);                       // tabu_search_base.rs:174
        } else if (A.agent == GJ_AGENT_SIMULATED_ANNEALING) {
            accept = gj_sa_step_accept(A, island, b, cur);              // simulated_annealing_base.rs:198-233
        } else {
            // late_acceptance_base.rs:196-213
            double* late = A.late + (size_t)island * A.late_size * GJ_MAX_LEVELS;
            int head = A.late_head[island], len = A.late_len[island];
            GjScore late_native = cur;
            if (len > 0) {
                int back = (head + len - 1) % A.late_size;
                late_native = gj_load_score(late + (size_t)back * GJ_MAX_LEVELS, levels);
            }
            accept = gj_score_le(b, late_native, levels) || gj_score_le(b, cur, levels);
            if (accept) {
                // push_front; pop_back when longer than late_acceptance_size
                head = (head + A.late_size - 1) % A.late_size;
                for (int l = 0; l < GJ_MAX_LEVELS; ++l) late[(size_t)head * GJ_MAX_LEVELS + l] = b.v[l];
                len = min(len + 1, A.late_size);
                A.late_head[island] = head; A.late_len[island] = len;
            }
        }
        if (accept)
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) A.cur_score[(size_t)island * GJ_MAX_LEVELS + l] = b.v[l];
        sh_accept = accept ? 1 : 0;
        sh_best = bi;
        if (A.selected_out) { A.selected_out[island] = bi; A.accepted_out[island] = accept ? 1 : 0; }
        atomicAdd(&A.counters[0], (unsigned long long)K);
        if (island == 0) atomicAdd(&A.counters[1], 1ull);
        if (accept) atomicAdd(&A.counters[2], 1ull);
        if (island == 0 && A.work_count) *A.work_count = 0;
    }
    __syncthreads();
    int32_t* cur_row = A.cur + (size_t)island * A.stride;
    if (sh_accept) {
        // apply the winning deltas to the stored individual (tabu_search_base.rs:175-178)
        for (int i = tid; i < A.n_vars; i += blockDim.x) smem_row[i] = cur_row[i];
        __syncthreads();
        const GjMove m = load_move(sh_best);
        gj_apply_move(P, m, G, true, A.noop != 0, tid, blockDim.x,
                      [&](int id) { return smem_row[id]; }, [&](int id, int v) { cur_row[id] = v; });
        if (tid == 0) {
            A.dirty[island] = 1;
            if (A.stale) A.stale[island] = 2;
        }
        __syncthreads();
    }
    if (!A.defer_top)
        gj_update_top(island, levels, A.stride, A.n_vars, A.cur, A.cur_score, A.best, A.best_score, A.dirty);

    if (A.tabu_bits)
        gj_tabu_deque_advance(A.tabu_bits + (size_t)island * A.tabu_words_per_island,
                              A.tabu_ring_old + (size_t)island * A.tabu_ring_per_island,
                              A.tabu_ring_new + (size_t)island * A.tabu_ring_per_island, A.tabu_ring_off,
                              A.tabu_size, A.tabu_word_off, A.tabu_fill + (size_t)island * A.n_groups,
                              A.n_groups, G, K, load_move, sh_scan);
}

// Applies pending adoptions outside a step (before the host reads / exports current solutions).
__global__ void __launch_bounds__(128)
k_apply_adoption(GjSelectArgs A) {
    __shared__ int sh_take;
    const int island = blockIdx.x;
    if (threadIdx.x == 0) sh_take = gj_adopt_decide(A, island) ? 1 : 0;
    __syncthreads();
    if (sh_take)
        for (int i = threadIdx.x; i < A.n_vars; i += blockDim.x) A.cur[(size_t)island * A.stride + i] = A.gbest[i];
}


// ---- migration (ring i -> i+1, solver.rs:85-92) ------------------------------------------------------
// mailbox slot s: [stride int32][GJ_MAX_LEVELS f64]; slot[i+1] = island i's outgoing migrant,
// slot[0] = what island 0 receives (the wrap-around or another GPU's last island).

// `parity` 0 / 1 restricts the launch to the even / odd islands (-1 = all): agent_base.rs:161-183
// lets even agents send before they receive and odd agents receive before they send.
__global__ void k_migrate_pack(const int32_t* __restrict__ cur, const double* __restrict__ cur_score,
                               int stride, int n_vars, unsigned char* mailbox, int parity) {
    const int island = blockIdx.x;
    if (parity >= 0 && (island & 1) != parity) return;
    unsigned char* slot = mailbox + (size_t)(island + 1) * gj_slot_bytes(stride);
    int32_t* row = (int32_t*)slot;
    double* sc = (double*)(slot + (size_t)stride * 4);
    for (int i = threadIdx.x; i < n_vars; i += blockDim.x) row[i] = cur[(size_t)island * stride + i];
    if (threadIdx.x < GJ_MAX_LEVELS) sc[threadIdx.x] = cur_score[(size_t)island * GJ_MAX_LEVELS + threadIdx.x];
}

__global__ void k_migrate_wrap(int I, int stride, unsigned char* mailbox) {
    const size_t sb = gj_slot_bytes(stride);
    const uint32_t* src = (const uint32_t*)(mailbox + (size_t)I * sb);
    uint32_t* dst = (uint32_t*)mailbox;
    for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < sb / 4; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

__global__ void k_migrate_recv(int agent, int levels, int stride, int n_vars, int late_size,
                               const unsigned char* __restrict__ mailbox, int32_t* cur,
                               double* cur_score, int* dirty, double* late, int* late_head,
                               int* late_len, int* stale, int parity) {
    __shared__ int sh_take;
    const int island = blockIdx.x;
    if (parity >= 0 && (island & 1) != parity) return;
    const unsigned char* slot = mailbox + (size_t)island * gj_slot_bytes(stride);
    const int32_t* row = (const int32_t*)slot;
    const double* sc = (const double*)(slot + (size_t)stride * 4);
    if (threadIdx.x == 0) {
        GjScore mig = gj_load_score(sc, levels);
        GjScore cs = gj_load_score(cur_score + (size_t)island * GJ_MAX_LEVELS, levels);
        bool take;
        if (agent != GJ_AGENT_LATE_ACCEPTANCE) {
            take = gj_score_le(mig, cs, levels);                    // agent_base.rs:429-439
        } else {
            // agent_base.rs:416-428 (late_scores.back() of an empty deque would panic in the
            // reference; an empty deque falls back to the current score here)
            double* lt = late + (size_t)island * late_size * GJ_MAX_LEVELS;
            int head = late_head[island], len = late_len[island];
            GjScore back = cs;
            if (len > 0) back = gj_load_score(lt + (size_t)((head + len - 1) % late_size) * GJ_MAX_LEVELS, levels);
            take = gj_score_le(mig, back, levels) || gj_score_le(mig, cs, levels);
            if (take) {
                head = (head + late_size - 1) % late_size;
                for (int l = 0; l < GJ_MAX_LEVELS; ++l) lt[(size_t)head * GJ_MAX_LEVELS + l] = mig.v[l];
                late_head[island] = head; late_len[island] = min(len + 1, late_size);
            }
        }
        if (take) {
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) cur_score[(size_t)island * GJ_MAX_LEVELS + l] = mig.v[l];
            dirty[island] = 1;
            if (stale) stale[island] = 1;
        }
        sh_take = take ? 1 : 0;
    }
    __syncthreads();
    if (sh_take)
        for (int i = threadIdx.x; i < n_vars; i += blockDim.x) cur[(size_t)island * stride + i] = row[i];
}

// lean fused islands (TSP-20000): when the published top is its island's CURRENT solution, that island's
// edge lengths and unrounded score terms are published with it -- an adopter then copies 160 KB
// sequentially instead of re-gathering 20 000 matrix entries from a 3.2 GB matrix (37 of 190 us per step)
struct GjGtopLean {
    const int* top_is_cur; const int* stale; const int* dirty;
    const double* edge; const double* raw;
    double* gedge; double* graw; int* gedge_ver;
};

__global__ void __launch_bounds__(1024)
k_global_top(int I, int levels, int stride, int n_vars, const int32_t* __restrict__ best,
             const double* __restrict__ best_score, int32_t* gbest, double* gbest_score, int* gver, GjGtopLean X) {
    __shared__ GjScore sh_s[32];
    __shared__ int sh_i[32];
    __shared__ int sh_publish;
    gj_global_top_cta(I, levels, stride, n_vars, best, best_score, gbest, gbest_score, gver, sh_s, sh_i, &sh_publish);
    if (X.gedge && sh_publish) {                       // (sh_publish / sh_i[0]: valid after the barriers inside)
        const int w = sh_i[0];
        const bool ok = X.top_is_cur[w] != 0 && X.stale[w] == 0 && X.dirty[w] == 0;
        if (ok) {
            const double* src = X.edge + (size_t)w * (size_t)(n_vars + 1);
            for (int i = threadIdx.x; i <= n_vars; i += blockDim.x) X.gedge[i] = __ldcg(&src[i]);
            if (threadIdx.x < GJ_MAX_LEVELS) X.graw[threadIdx.x] = __ldcg(&X.raw[(size_t)w * GJ_MAX_LEVELS + threadIdx.x]);
        }
        __syncthreads();
        if (threadIdx.x == 0) *X.gedge_ver = ok ? *gver : 0;
    }
}

// Expands move descriptors into the (column, value) lists of the reference's incremental form.
__global__ void k_expand_moves(GjProblemDev P, GjGroups G, const int32_t* __restrict__ base,
                               const GjMove* __restrict__ moves, int K, int noop,
                               const uint64_t* __restrict__ offsets, uint64_t* ids, double* vals) {
    const int j = blockIdx.x;
    if (j >= K) return;
    const GjMove m = moves[j];
    const uint64_t o = offsets[j];
    if (m.kind == GJ_MOVE_NULL) return;
    const int32_t* g = G.ids + G.offsets[m.group];
    if (m.kind <= 3) {
        if (threadIdx.x == 0) {
            int cols[GJ_MOVE_MAXPAIRS], v[GJ_MOVE_MAXPAIRS];
            const int n = gj_small_move_pairs(m, g, true, noop != 0, [&](int id) { return base[id]; }, cols, v);
            for (int i = 0; i < n; ++i) { ids[o + i] = (uint64_t)cols[i]; vals[o + i] = (double)gj_fix_column(P, cols[i], v[i]); }
        }
        return;
    }
    int lo, hi;
    gj_segment_bounds(m, lo, hi);
    const int len = hi - lo + 1;
    for (int t = threadIdx.x; t < len; t += blockDim.x) {
        const int s = gj_segment_src_slot(m, true, t, len);
        ids[o + t] = (uint64_t)g[lo + t];
        vals[o + t] = (double)gj_fix_column(P, g[lo + t], base[g[lo + s]]);
    }
}

__global__ void k_i32_to_f64(const int32_t* __restrict__ in, double* __restrict__ out, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = (double)in[i];
}

// ===================================================================================================
// host side
// ===================================================================================================

gj_islands::~gj_islands() {
    cudaSetDevice(device);
    if (phase_clocks) {
        // development aid: phase durations (cycles) of the LAST fused step, island 0 and the mean
        std::vector<long long> h((size_t)I * 8);
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), phase_clocks, h.size() * 8, cudaMemcpyDeviceToHost);
        static const char* names[5] = {"P0 stage", "P1 score", "P2 select", "P3 apply", "P4 tabu"};
        for (int k = 0; k < 5; ++k) {
            double mean = 0;
            for (int i = 0; i < I; ++i) mean += (double)(h[(size_t)i * 8 + k + 1] - h[(size_t)i * 8 + k]);
            fprintf(stderr, "[gj phase] %-10s island0 %8lld cycles, mean %10.0f\n", names[k], h[k + 1] - h[k], mean / I);
        }
        if (h[6]) {
            double mean = 0;
            for (int i = 0; i < I; ++i) mean += (double)(h[(size_t)i * 8 + 6] - h[(size_t)i * 8]);
            fprintf(stderr, "[gj phase] P0 until the bulk copies landed: island0 %8lld cycles, mean %10.0f\n", h[6] - h[0], mean / I);
        }
    }
    if (ga_side) { cudaStreamSynchronize(ga_side); cudaStreamDestroy(ga_side); }
    for (auto& e : ga_ev) if (e) cudaEventDestroy(e);
    for (void* a : allocs) cudaFree(a);
    for (auto& e : prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
}

gj_status gj_prof_begin(gj_islands* g, cudaStream_t st) {
    if (!g->profiling) return GJ_OK;
    if (g->prof_used == g->prof_events.size()) {
        cudaEvent_t a, b;
        GJ_CUDA_TRY(cudaEventCreate(&a));
        GJ_CUDA_TRY(cudaEventCreate(&b));
        g->prof_events.emplace_back(a, b);
    }
    GJ_CUDA_TRY(cudaEventRecord(g->prof_events[g->prof_used].first, st));
    return GJ_OK;
}

gj_status gj_prof_end(gj_islands* g, cudaStream_t st) {
    if (!g->profiling) return GJ_OK;
    GJ_CUDA_TRY(cudaEventRecord(g->prof_events[g->prof_used].second, st));
    g->prof_used += 1;
    return GJ_OK;
}

extern "C" gj_status gj_islands_set_profiling(gj_islands* g, int32_t on) {
    if (!g) return gj_fail(GJ_ERR_INVALID, "null handle");
    g->profiling = on != 0;
    g->prof_used = 0;
    return GJ_OK;
}

extern "C" gj_status gj_islands_profile_read(gj_islands* g, double* total_ms, int64_t* launches) {
    if (!g) return gj_fail(GJ_ERR_INVALID, "null handle");
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    double tot = 0.0;
    for (size_t i = 0; i < g->prof_used; ++i) {
        GJ_CUDA_TRY(cudaEventSynchronize(g->prof_events[i].second));
        float ms = 0.f;
        GJ_CUDA_TRY(cudaEventElapsedTime(&ms, g->prof_events[i].first, g->prof_events[i].second));
        tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = (int64_t)g->prof_used;
    g->prof_used = 0;
    return GJ_OK;
}

template <class T>
static gj_status dev_alloc(gj_islands* g, size_t n, T** out, bool zero = true) {
    void* d = nullptr;
    size_t bytes = (n ? n : 1) * sizeof(T);
    GJ_CUDA_TRY(cudaMalloc(&d, bytes));
    g->allocs.push_back(d);
    if (zero) GJ_CUDA_TRY(cudaMemset(d, 0, bytes));
    *out = (T*)d;
    return GJ_OK;
}

template <class T>
static gj_status dev_upload(gj_islands* g, const std::vector<T>& h, const T** out) {
    T* d = nullptr;
    gj_status rc = dev_alloc(g, h.size(), &d, false);
    if (rc) return rc;
    if (!h.empty()) GJ_CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = d;
    return GJ_OK;
}

static uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Tabu deque size of a semantic group: max(ceil(rate * group_len), 1) (tabu_search_base.rs:115-121).
// The reference draws ids until it finds k that are not tabu (mover.rs:75-96) and would spin
// forever on a group with fewer than k free positions; here the deque is capped so that
// GJ_MOVE_MAXK positions always stay free (only bites when rate * group_len > group_len - 8).
int gj_tabu_deque_size(double rate, int group_len) {
    int T = std::max((int)std::ceil(rate * (double)group_len), 1);
    if (T > group_len - GJ_MOVE_MAXK) T = std::max(1, group_len - GJ_MOVE_MAXK);
    return T;
}

// Mover::new thresholds (mover.rs:36-62)
static gj_status build_thresholds(const gj_agent_params& prm, double* thr) {
    double probas[6];
    if (!prm.has_move_probas) {
        // round(1/6, 3) each, remainder to move 0
        const double inc = std::floor((1.0 / 6.0) * 1000.0) / 1000.0;
        double sum = 0.0;
        for (int i = 0; i < 6; ++i) { probas[i] = inc; sum += inc; }
        probas[0] += 1.0 - sum;
    } else {
        double sum = 0.0;
        for (int i = 0; i < 6; ++i) { probas[i] = prm.move_probas[i]; sum += probas[i]; }
        const double r1 = std::floor(sum) + std::floor((sum - std::floor(sum)) * 10.0) / 10.0;
        if (r1 != 1.0 && std::fabs(sum - 1.0) > 1e-9)
            return gj_fail(GJ_ERR_INVALID, "Optional move probas sum must be equal to 1.0");
    }
    double acc = 0.0;
    for (int i = 0; i < 6; ++i) { acc += probas[i]; thr[i] = acc; }
    if (thr[5] < 1.0) thr[5] = 1.0;     // the reference panics when u exceeds the last threshold
    return GJ_OK;
}

gj_status gj_islands_common_init(gj_islands* g, gj_problem* p, const gj_agent_params* prm) {
    g->p = p;
    g->device = p->device;
    g->prm = *prm;
    g->I = prm->n_islands;
    g->levels = p->dev.levels;
    g->n_vars = p->dev.n_vars;
    g->stride = (p->dev.n_vars + 3) & ~3;
    g->noop = prm->reference_noop_moves ? 1 : 0;
    gj_status rc;
    if ((rc = build_thresholds(*prm, g->mover.thresholds))) return rc;
    g->mover.tabu_entity_rate = prm->tabu_entity_rate;
    g->mover.mutation_rate_multiplier = prm->has_mutation_rate_multiplier ? prm->mutation_rate_multiplier : 0.0;
    // change count of a change / swap / swap_edges move ~ Binomial(n_vars, multiplier / group_len)
    // (mover.rs:138-140, unbounded); a device move descriptor holds GJ_MOVE_MAXK = 8 positions, so the
    // draw is truncated at 8 (declared in DESIGN.md section 2).  The reference's examples run at a mean of
    // <= 2 (multiplier <= 1, groups of n_vars / 2), where the truncation touches < 0.03 % of the moves;
    // 2.1 % at a mean of 4, 41 % at 8.  Beyond that the move-size distribution would mostly be the cap:
    // refused instead of silently distorted.
    for (auto& grp : p->groups) {
        if (grp.empty()) continue;
        const double mean = g->mover.mutation_rate_multiplier * (double)p->dev.n_vars / (double)grp.size();
        if (mean > 8.0 + 1e-9)
            return gj_fail(GJ_ERR_UNSUPPORTED, "mutation_rate_multiplier * n_vars / group_len > 8: most moves would change "
                                                "more than the 8 positions a device move descriptor holds");
    }

    // semantic groups
    std::vector<int32_t> offs(1, 0), ids;
    for (auto& grp : p->groups) {
        ids.insert(ids.end(), grp.begin(), grp.end());
        offs.push_back((int32_t)ids.size());
    }
    if (p->groups.empty() || p->groups.size() > 255) return gj_fail(GJ_ERR_INVALID, "1..255 semantic groups required");
    g->groups.n_groups = (int)p->groups.size();
    if ((rc = dev_upload(g, offs, &g->groups.offsets))) return rc;
    if ((rc = dev_upload(g, ids, &g->groups.ids))) return rc;
    {
        // per group {first, step, uniform bounds}: lets the delta evaluator treat a segment of an
        // affine group as a contiguous run of columns, and skip fix_deltas clamps
        std::vector<int4> info;
        const GjProblemDev& P = p->dev;
        std::vector<int32_t> lbi(P.n_vars), ubi(P.n_vars);
        GJ_CUDA_TRY(cudaMemcpy(lbi.data(), P.lbi, (size_t)P.n_vars * 4, cudaMemcpyDeviceToHost));
        GJ_CUDA_TRY(cudaMemcpy(ubi.data(), P.ubi, (size_t)P.n_vars * 4, cudaMemcpyDeviceToHost));
        for (auto& grp : p->groups) {
            int4 gi = make_int4(grp.empty() ? 0 : grp[0], 0, 1, 0);
            if (grp.size() >= 2) {
                gi.y = grp[1] - grp[0];
                for (size_t k = 2; k < grp.size() && gi.y != 0; ++k)
                    if (grp[k] - grp[k - 1] != gi.y) gi.y = 0;
            } else {
                gi.y = 1;
            }
            for (size_t k = 1; k < grp.size(); ++k)
                if (lbi[grp[k]] != lbi[grp[0]] || ubi[grp[k]] != ubi[grp[0]]) gi.z = 0;
            info.push_back(gi);
        }
        if ((rc = dev_upload(g, info, &g->groups.info))) return rc;
    }

    // tabu deques: size = max(ceil(rate * group_len), 1) (tabu_search_base.rs:115-121)
    if (prm->tabu_entity_rate != 0.0) {
        std::vector<int32_t> word_off, ring_off, tsize;
        int words = 0, ring = 0;
        for (auto& grp : p->groups) {
            word_off.push_back(words); ring_off.push_back(ring);
            const int T = gj_tabu_deque_size(prm->tabu_entity_rate, (int)grp.size());
            tsize.push_back(T);
            words += gj_tabu_region_words((int)grp.size());       // bits + free-prefix + free list (gj_moves.cuh)
            ring += T;
        }
        words = (words + 3) & ~3;                     // 16-byte multiple: staged with one TMA bulk copy
        g->tabu_words = words; g->tabu_ring_len = ring;
        if ((rc = dev_upload(g, word_off, &g->tabu_word_off))) return rc;
        if ((rc = dev_upload(g, ring_off, &g->tabu_ring_off))) return rc;
        if ((rc = dev_upload(g, tsize, &g->tabu_size))) return rc;
        if ((rc = dev_alloc(g, (size_t)g->I * words, &g->tabu_bits))) return rc;
        {
            // empty deques: every position free
            std::vector<uint32_t> table((size_t)g->I * words, 0u);
            for (int i = 0; i < g->I; ++i)
                for (size_t gi = 0; gi < p->groups.size(); ++gi) {
                    const int glen = (int)p->groups[gi].size(), W = (glen + 31) / 32;
                    uint32_t* prefix = table.data() + (size_t)i * words + word_off[gi] + W + 1;
                    for (int w = 0; w <= W; ++w) prefix[w] = (uint32_t)std::min(32 * w, glen);
                    uint32_t* free_list = prefix + W + 1;
                    for (int k = 0; k < glen; ++k) free_list[k] = (uint32_t)k;
                }
            GJ_CUDA_TRY(cudaMemcpy(g->tabu_bits, table.data(), table.size() * 4, cudaMemcpyHostToDevice));
        }
        if ((rc = dev_alloc(g, (size_t)g->I * ring, &g->tabu_ring[0]))) return rc;
        if ((rc = dev_alloc(g, (size_t)g->I * ring, &g->tabu_ring[1]))) return rc;
        if ((rc = dev_alloc(g, (size_t)g->I * g->groups.n_groups, &g->tabu_fill))) return rc;
    }
    if ((rc = dev_alloc(g, 4, &g->counters))) return rc;
    return GJ_OK;
}

// Start vectors: InitialSolutionVariants / GJInteger::get_initial_value (gj_integer.rs:98-112):
// the given initial value, else a uniform sample in [lb, ub].
void gj_islands_start_vector(const gj_problem* p, const double* given, uint64_t& rng, std::vector<int32_t>& row) {
    const int n = p->dev.n_vars;
    for (int i = 0; i < n; ++i) {
        double x = given ? given[i] : p->initial[i];
        if (p->frozen[i]) x = p->initial[i];
        if (!(x == x) || x < p->lb[i] - 0.5 || (given == nullptr && !(p->initial[i] == p->initial[i]))) {
            const int64_t lo = (int64_t)std::llround(p->lb[i]), hi = (int64_t)std::llround(p->ub[i]);
            x = (double)(lo + (int64_t)(splitmix64(rng) % (uint64_t)(hi - lo + 1)));
        }
        // decode exactly like the device would (clamp + rint ties-to-ceil)
        double lo = p->lb[i], hi = p->ub[i];
        if (!p->frozen[i]) {
            if (x < lo) x = lo;
            if (x > hi) x = hi;
            double f = std::floor(x), c = std::ceil(x);
            x = (std::fabs(x - f) < std::fabs(c - x)) ? f : c;
        }
        row[i] = (int32_t)x;
    }
}

static gj_status score_cur(gj_islands* g, cudaStream_t st) {
    // Agent::init_population (agent_base.rs:206-213): the start vector scored through the ISC
    // path as one delta list holding every variable; NOT rounded (rounding only happens in
    // step_*).  Scored here by the int32 plain kernel with ISC semantics (same arithmetic).
    return gj_launch_score_plain_i32(g->p, g->cur, g->stride, g->I, g->cand_scores, true, st);
}

__global__ void k_init_scores(int I, const double* __restrict__ scored, int levels, double* cur_score,
                              double* best_score, double* gbest_score) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < I) {
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) {
            const double v = (l < levels) ? scored[(size_t)i * levels + l] : 0.0;
            cur_score[(size_t)i * GJ_MAX_LEVELS + l] = v;
            best_score[(size_t)i * GJ_MAX_LEVELS + l] = v;
        }
    }
    if (i == 0)
        for (int l = 0; l < GJ_MAX_LEVELS; ++l)
            gbest_score[l] = (l < levels) ? 1.7976931348623157e308 : 0.0;   // get_stub_score()
}

static gj_status launch_refresh(gj_islands* g, cudaStream_t st, bool update_top);

static gj_status ls_create(gj_problem* p, const gj_agent_params* prm_in, const double* initial, gj_islands** out) {
    // GJ_SCORING_DELTA_F64 = DELTA minus the fixed-point TSP step
    gj_agent_params prm_delta = *prm_in;
    const bool no_fixed_point = prm_in->scoring_mode == GJ_SCORING_DELTA_F64;
    if (no_fixed_point) prm_delta.scoring_mode = GJ_SCORING_DELTA;
    const gj_agent_params* prm = &prm_delta;
    std::unique_ptr<gj_islands> g(new gj_islands());
    gj_status rc;
    if ((rc = gj_islands_common_init(g.get(), p, prm))) return rc;
    const bool ts = prm->agent == GJ_AGENT_TABU_SEARCH;
    g->K = ts ? (int)prm->neighbours_count : 1;
    if (g->K < 1) return gj_fail(GJ_ERR_INVALID, "neighbours_count must be >= 1");
    const bool la = prm->agent == GJ_AGENT_LATE_ACCEPTANCE;
    const bool sa = prm->agent == GJ_AGENT_SIMULATED_ANNEALING;
    if (la && prm->late_acceptance_size < 1) return gj_fail(GJ_ERR_INVALID, "late_acceptance_size must be >= 1");
    const int I = g->I, stride = g->stride;
    if ((rc = dev_alloc(g.get(), (size_t)I * stride, &g->cur))) return rc;
    if ((rc = dev_alloc(g.get(), (size_t)I * stride, &g->best))) return rc;
    if ((rc = dev_alloc(g.get(), (size_t)stride, &g->gbest))) return rc;
    if ((rc = dev_alloc(g.get(), (size_t)I * GJ_MAX_LEVELS, &g->cur_score))) return rc;
    if ((rc = dev_alloc(g.get(), (size_t)I * GJ_MAX_LEVELS, &g->best_score))) return rc;
    if ((rc = dev_alloc(g.get(), (size_t)GJ_MAX_LEVELS, &g->gbest_score))) return rc;
    if ((rc = dev_alloc(g.get(), (size_t)I, &g->dirty))) return rc;
    if ((rc = dev_alloc(g.get(), (size_t)I * g->K, &g->moves))) return rc;
    if ((rc = dev_alloc(g.get(), (size_t)I * std::max(g->K, 1) * GJ_MAX_LEVELS, &g->cand_scores))) return rc;
    if ((rc = dev_alloc(g.get(), (size_t)(I + 1) * ((size_t)stride * 4 + GJ_MAX_LEVELS * 8), &g->mailbox))) return rc;
    if (sa) {
        g->sa.has_cooling = prm->has_cooling_rate ? 1 : 0;
        g->sa.cooling_rate = prm->cooling_rate;
        g->sa.inv_rate = 1.0;                           // inverted_accomplish_rate starts at 1.0
        std::vector<double> t0((size_t)I * GJ_MAX_LEVELS, 1.0);
        for (int i = 0; i < I; ++i)
            for (int l = 0; l < p->dev.levels; ++l) {
                if (!(prm->initial_temperature[l] > 0.0)) return gj_fail(GJ_ERR_INVALID, "initial_temperature must be > 0 for every score level");
                t0[(size_t)i * GJ_MAX_LEVELS + l] = prm->initial_temperature[l];
            }
        if ((rc = dev_alloc(g.get(), t0.size(), &g->sa_temp, false))) return rc;
        GJ_CUDA_TRY(cudaMemcpy(g->sa_temp, t0.data(), t0.size() * 8, cudaMemcpyHostToDevice));
        if ((rc = dev_alloc(g.get(), (size_t)I * 5, &g->trace_aux))) return rc;
    }
    if (la) {
        g->late_size = (int)prm->late_acceptance_size;
        if ((rc = dev_alloc(g.get(), (size_t)I * g->late_size * GJ_MAX_LEVELS, &g->late))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I, &g->late_head))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I, &g->late_len))) return rc;
    }
    if ((rc = dev_alloc(g.get(), 1, &g->gver))) return rc;
    if ((rc = dev_alloc(g.get(), (size_t)I, &g->gseen))) return rc;
    if ((rc = dev_alloc(g.get(), (size_t)I, &g->selected))) return rc;
    if ((rc = dev_alloc(g.get(), (size_t)I, &g->accepted))) return rc;

    std::vector<int32_t> host((size_t)I * stride, 0), row(p->dev.n_vars);
    uint64_t rng = prm->seed ^ 0xA5A5A5A55A5A5A5Aull;
    for (int i = 0; i < I; ++i) {
        gj_islands_start_vector(p, initial ? initial + (size_t)i * p->dev.n_vars : nullptr, rng, row);
        std::copy(row.begin(), row.end(), host.begin() + (size_t)i * stride);
    }
    GJ_CUDA_TRY(cudaMemcpy(g->cur, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
    GJ_CUDA_TRY(cudaMemcpy(g->best, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
    cudaStream_t st = p->stream;
    if ((rc = score_cur(g.get(), st))) return rc;
    k_init_scores<<<(I + 127) / 128, 128, 0, st>>>(I, g->cand_scores, g->levels, g->cur_score, g->best_score, g->gbest_score);
    GJ_LAUNCH_CHECK();
    // delta scoring: cached per-island state (gj_delta.cuh).  The VRP models are scored by the
    // full evaluator in either mode for now.
    // VRP models: route-level delta evaluation (gj_vrp_delta.cuh), separate kernels per step.  The base
    // state is rebuilt (one full evaluation per island) after every accepted neighbour, which only
    // pays off when a step scores several neighbours: single-neighbour agents (LateAcceptance,
    // SimulatedAnnealing) keep the full evaluator.
    if ((prm->scoring_mode == GJ_SCORING_DELTA || prm->scoring_mode == GJ_SCORING_DELTA_UNFUSED) &&
        p->dev.kind >= GJ_VRP && g->K >= 8) {
        g->scoring_mode = GJ_SCORING_DELTA;
        const GjProblemDev& P = p->dev;
        const int n = P.n_entities, K = P.n_vehicles;
        g->ds.cnt_stride = 32 * P.bm_words;
        if ((rc = dev_alloc(g.get(), (size_t)I * g->ds.cnt_stride, &g->ds.cnt))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I * GJ_MAX_LEVELS, &g->ds.raw))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I, &g->ds.stale))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I * g->K, &g->worklist))) return rc;
        if ((rc = dev_alloc(g.get(), 1, &g->work_count))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I * n, &g->vs.bucket))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I * n, &g->vs.bstop))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I * (K + 1), &g->vs.start))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I * K, &g->vs.rdist))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I * K, &g->vs.rload))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I * K, &g->vs.rlate))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I * 4, &g->vs.tot))) return rc;
        std::vector<int> ones((size_t)I, 1);
        GJ_CUDA_TRY(cudaMemcpy(g->ds.stale, ones.data(), (size_t)I * sizeof(int), cudaMemcpyHostToDevice));
        g->delta_may_fallback = g->mover.thresholds[3] < 1.0;      // insertion / inverse possible
        if ((rc = launch_refresh(g.get(), st, false))) return rc;
    }
    // VRP models, single-neighbour agents (LateAcceptance / SimulatedAnnealing): chains that run many
    // steps per launch over a route index kept in HBM (gj_islands_vrp_chain.cuh).  Needs move_probas
    // without segment moves (insertion / inverse re-label O(segment) stops) and room for the index.
    if (prm->scoring_mode == GJ_SCORING_DELTA && p->dev.kind >= GJ_VRP && (la || sa) &&
        g->mover.thresholds[3] >= 1.0) {
        const GjProblemDev& P = p->dev;
        const size_t n = (size_t)P.n_entities, K = (size_t)P.n_vehicles;
        size_t free_b = 0, total_b = 0;
        GJ_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = ((size_t)I + 1) * (K * n + n) * 4;
        if (need <= free_b / 2 && K * n < ((size_t)1 << 31)) {
            g->scoring_mode = GJ_SCORING_DELTA;
            g->chain = true; g->vrp_chain = true;
            g->mover.tabu_layout = 1;
            GjVrpChainState& V = g->vcs;
            V.cnt_stride = 32 * P.bm_words;
            const size_t I1 = (size_t)I + 1;            // slot I: the published global top's index
            if ((rc = dev_alloc(g.get(), I1 * K * n, &V.rs, false))) return rc;
            if ((rc = dev_alloc(g.get(), I1 * K, &V.rlen))) return rc;
            if ((rc = dev_alloc(g.get(), I1 * K, &V.rdist))) return rc;
            if ((rc = dev_alloc(g.get(), I1 * K, &V.rload))) return rc;
            if ((rc = dev_alloc(g.get(), I1 * K, &V.rlate))) return rc;
            if ((rc = dev_alloc(g.get(), I1 * 4, &V.tot))) return rc;
            if ((rc = dev_alloc(g.get(), (size_t)I * n, &V.spare, false))) return rc;
            if ((rc = dev_alloc(g.get(), I1 * V.cnt_stride, &V.cnt))) return rc;
            if ((rc = dev_alloc(g.get(), n, &V.gstop))) return rc;
            if ((rc = dev_alloc(g.get(), n, &V.gdst))) return rc;
            if ((rc = dev_alloc(g.get(), 1, &V.gidx_ver))) return rc;
            if ((rc = dev_alloc(g.get(), K, &V.goff))) return rc;
            if ((rc = dev_alloc(g.get(), (size_t)I * GJ_VRPC_DIFF, &V.diff))) return rc;
            if ((rc = dev_alloc(g.get(), (size_t)I, &V.ndiff))) return rc;
            if ((rc = dev_alloc(g.get(), (size_t)I, &V.pend))) return rc;
            if ((rc = dev_alloc(g.get(), I1, &g->ds.stale))) return rc;
            V.stale = g->ds.stale;
            std::vector<int> ones((size_t)I, 1);
            GJ_CUDA_TRY(cudaMemcpy(g->ds.stale, ones.data(), (size_t)I * sizeof(int), cudaMemcpyHostToDevice));
            if (prm->tabu_entity_rate != 0.0) {
                std::vector<int32_t> coff;
                int cw = 0;
                for (auto& grp : p->groups) {
                    coff.push_back(cw);
                    const int T = gj_tabu_deque_size(prm->tabu_entity_rate, (int)grp.size());
                    cw += ((int)grp.size() + 31) / 32 + 1 + T + 2;      // bits | ring | head, fill
                }
                g->ctabu_words = cw;
                if ((rc = dev_upload(g.get(), coff, &g->ctabu_off))) return rc;
                if ((rc = dev_alloc(g.get(), (size_t)I * cw, &g->ctabu))) return rc;
            }
        }
    }
    if ((prm->scoring_mode == GJ_SCORING_DELTA || prm->scoring_mode == GJ_SCORING_DELTA_UNFUSED) && p->dev.kind <= GJ_TSP) {
        g->scoring_mode = GJ_SCORING_DELTA;
        const GjProblemDev& P = p->dev;
        g->ds.cnt_stride = 32 * (P.bm_words + P.desc_words + P.asc_words);
        if ((rc = dev_alloc(g.get(), (size_t)I * g->ds.cnt_stride, &g->ds.cnt))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I * GJ_MAX_LEVELS, &g->ds.raw))) return rc;
        if (P.kind == GJ_TSP && (rc = dev_alloc(g.get(), (size_t)I * (size_t)(P.n_vars + 1), &g->ds.edge, false))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I, &g->ds.stale))) return rc;
        if ((rc = dev_alloc(g.get(), (size_t)I * g->K, &g->worklist))) return rc;
        if ((rc = dev_alloc(g.get(), 1, &g->work_count))) return rc;
        std::vector<int> ones((size_t)I, 1);
        GJ_CUDA_TRY(cudaMemcpy(g->ds.stale, ones.data(), (size_t)I * sizeof(int), cudaMemcpyHostToDevice));
        // can a generated move need the full evaluator?  (thresholds are cumulative)
        const double* thr = g->mover.thresholds;
        const bool seg_moves = thr[3] < 1.0;                  // insertion / inverse possible
        const bool inverse = thr[4] < 1.0;
        bool affine = true;
        for (auto& grp : p->groups) {
            bool a = true, u = true;
            for (size_t k = 1; k < grp.size(); ++k) {
                if (grp[k] - grp[k - 1] != 1) a = false;
                if (p->lb[grp[k]] != p->lb[grp[0]] || p->ub[grp[k]] != p->ub[grp[0]]) u = false;
            }
            affine = affine && a && u;
        }
        if (P.kind == GJ_NQUEENS) g->delta_may_fallback = seg_moves;
        else g->delta_may_fallback = seg_moves && (!affine || (inverse && !p->symmetric_D));
        // fused single-kernel step when the island fits in shared memory
        if (prm->scoring_mode == GJ_SCORING_DELTA) {
            const int words = P.bm_words + P.desc_words + P.asc_words;
            // CTA size: an SM holds 1024 threads of this kernel (64 registers each); several
            // smaller CTAs (islands) per SM overlap one island's serial phases (stage, select,
            // apply, tabu) with another's scoring loop
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
            int per_sm = std::max(1, std::min(4, (I + sms - 1) / sms));   // 4 x ~50 KB smem per SM
            int threads = 1024;
            while (threads > 256 && 1024 / threads < per_sm) threads /= 2;
            threads = std::min(threads, std::max(32, ((g->K + 31) / 32) * 32));
            if (const char* e = getenv("GJ_FUSED_THREADS")) threads = std::max(32, std::min(1024, atoi(e) / 32 * 32));
            for (int clones = std::min(4, threads / 32); clones >= 0; --clones) {
                const size_t b = gj_fused_smem_bytes(P.n_vars, g->ds.cnt_stride, g->tabu_words, words, clones);
                if (b <= 200 * 1024 && (clones > 0 || !g->delta_may_fallback)) {
                    g->fused = true; g->fused_clones = clones; g->fused_smem = b; g->fused_threads = threads;
                    break;
                }
            }
            // long solutions (TSP-20000): lean layout -- only the solution in shared memory -- when
            // every move the mover can generate takes the delta fast path (two-stop swaps, 2-opt,
            // insertions on consecutive uniform columns; the reference's no-op scramble / swap_edges)
            if (!g->fused && P.kind == GJ_TSP && !g->delta_may_fallback) {
                const bool quirk_only = prm->reference_noop_moves != 0;
                const bool need_cnt = thr[0] > 0.0 || g->mover.mutation_rate_multiplier != 0.0 || !affine ||
                                      ((thr[2] - thr[1] > 0.0 || thr[3] - thr[2] > 0.0) && !quirk_only);
                const size_t b = gj_fused_smem_bytes_lean(P.n_vars, words);
                if (!need_cnt && b <= 110 * 1024) {
                    g->fused = true; g->fused_lean = true; g->fused_clones = 0; g->fused_smem = b;
                    g->fused_threads = std::max(threads, 512);      // at most two CTAs per SM fit
                    if ((rc = dev_alloc(g.get(), (size_t)I, &g->top_is_cur))) return rc;
                    if ((rc = dev_alloc(g.get(), (size_t)P.n_vars + 2, &g->gedge))) return rc;
                    if ((rc = dev_alloc(g.get(), (size_t)GJ_MAX_LEVELS, &g->graw))) return rc;
                    if ((rc = dev_alloc(g.get(), (size_t)1, &g->gedge_ver))) return rc;
                }
            }
        }
        // TSP + TabuSearch on a milli-unit matrix: the fused step in fixed point (gj_islands_tsfast.cuh)
        if (g->fused && !g->fused_lean && P.kind == GJ_TSP && ts && p->groups.size() == 1 && affine &&
            p->groups[0].size() >= 16 && p->symmetric_D && g->mover.mutation_rate_multiplier == 0.0 &&
            P.w[0] == 1.0 && p->precision[1] == 3 && !no_fixed_point && !getenv("GJ_NO_TSFAST")) {
            if ((rc = gj_problem_ensure_d32(p))) return rc;
            if (p->d32_state == 1) {
                g->ts_edge_stride = (P.n_vars + 2) & ~1;
                if ((rc = dev_alloc(g.get(), (size_t)I * g->ts_edge_stride, &g->ts_edge))) return rc;
                if ((rc = dev_alloc(g.get(), 1, &g->done_counter))) return rc;
                if ((rc = dev_alloc(g.get(), 4, &g->ts_pub, false))) return rc;
                if ((rc = dev_alloc(g.get(), (size_t)g->ts_edge_stride, &g->gedge))) return rc;
                if ((rc = dev_alloc(g.get(), (size_t)1, &g->gedge_ver))) return rc;
                {
                    const unsigned long long init[4] = {~0ull, ~0ull, 0ull, 0ull};
                    GJ_CUDA_TRY(cudaMemcpy(g->ts_pub, init, sizeof(init), cudaMemcpyHostToDevice));
                }
                g->ts_fast = true;
                g->fused_smem = gj_tsfast_smem(g.get());
                // CTA shape from the islands per SM: 256 threads x 4 resident CTAs (<= 64 registers) up to
                // 4 islands per SM, x 6 (<= 40 registers) beyond; wider CTAs when islands are few
                int sms = 148;
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
                int threads = I >= 3 * sms ? 256 : (I >= 2 * sms ? 512 : 1024);
                // (measured on C2: six resident CTAs at <= 40 registers spill in the scoring loop and lose to
                // four at 64 even when there are islands to fill them: 30 vs 38 G candidates/s at 1184 islands)
                int mb = threads == 256 ? 4 : (threads == 512 ? 2 : 1);
                if (const char* e = getenv("GJ_FUSED_THREADS")) threads = std::max(128, std::min(1024, atoi(e)));
                if (const char* e = getenv("GJ_FUSED_MB")) mb = atoi(e);
                g->fused_mb = mb;
                g->fused_threads = threads;
            }
        }
        // LateAcceptance (one neighbour per step): chains that run many steps per launch
        if (prm->scoring_mode == GJ_SCORING_DELTA && (la || sa)) {
            const int words = P.bm_words + P.desc_words + P.asc_words;
            std::vector<int32_t> coff;
            int cw = 0;
            for (auto& grp : p->groups) {
                coff.push_back(cw);
                const int T = gj_tabu_deque_size(prm->tabu_entity_rate, (int)grp.size());
                cw += ((int)grp.size() + 31) / 32 + 1 + T + 2;      // bits | ring | head, fill
            }
            const size_t per_chain = gj_chain_smem_bytes(P.n_vars, words, cw, g->late_size, P.kind == GJ_TSP);
            if (per_chain * kChainWarps <= 200 * 1024) {
                g->chain = true; g->fused = false;
                g->chain_bytes = per_chain;
                g->mover.tabu_layout = 1;
                if (prm->tabu_entity_rate != 0.0) {
                    g->ctabu_words = cw;
                    if ((rc = dev_upload(g.get(), coff, &g->ctabu_off))) return rc;
                    if ((rc = dev_alloc(g.get(), (size_t)I * cw, &g->ctabu))) return rc;
                }
            }
        }
        if (g->fused && getenv("GJ_PHASE_TIMING")) {
            if ((rc = dev_alloc(g.get(), (size_t)I * 8, &g->phase_clocks))) return rc;
        }
        if (g->fused) {
            // the first step's P0 builds the state (stale == 1)
        } else if ((rc = launch_refresh(g.get(), st, false))) return rc;
    }
    GJ_CUDA_TRY(cudaStreamSynchronize(st));
    g->steps_to_send = (int64_t)std::max<int64_t>(1, prm->migration_frequency);
    *out = g.release();
    return GJ_OK;
}

static gj_status launch_score_moves(gj_islands* g, cudaStream_t st) {
    const GjProblemDev& P = g->p->dev;
    const int64_t total = (int64_t)g->I * g->K;
    if (P.kind >= GJ_VRP) {
        size_t smem = gj_vrp_smem_bytes(P.n_entities, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
        if (smem > 48 * 1024) GJ_CUDA_TRY(cudaFuncSetAttribute(k_score_moves_vrp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_score_moves_vrp<<<(unsigned)total, kVrpWarps * 32, smem, st>>>(P, g->groups, g->cur, g->stride, g->moves, g->K, total, 1, g->noop, 1, g->cand_scores);
    } else {
        // one shared-memory clone per warp: fewer warps per CTA for large instances
        const size_t per_warp = (size_t)(P.bm_words + P.desc_words + P.asc_words + P.n_vars) * 4;
        const int warps = (int)std::min<size_t>(kWarps, (220 * 1024) / per_warp);
        if (warps < 1) return gj_fail(GJ_ERR_UNSUPPORTED, "instance too large for the shared-memory candidate clone");
        const size_t smem = per_warp * warps;
        unsigned grid = (unsigned)((total + warps - 1) / warps);
        if (P.kind == GJ_NQUEENS) {
            if (smem > 48 * 1024) GJ_CUDA_TRY(cudaFuncSetAttribute(k_score_moves_warp<GJ_NQUEENS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_score_moves_warp<GJ_NQUEENS><<<grid, warps * 32, smem, st>>>(P, g->groups, g->cur, g->stride, g->moves, g->K, total, 1, g->noop, 1, g->cand_scores);
        } else {
            if (smem > 48 * 1024) GJ_CUDA_TRY(cudaFuncSetAttribute(k_score_moves_warp<GJ_TSP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_score_moves_warp<GJ_TSP><<<grid, warps * 32, smem, st>>>(P, g->groups, g->cur, g->stride, g->moves, g->K, total, 1, g->noop, 1, g->cand_scores);
        }
    }
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

GjSelectArgs gj_make_select_args(gj_islands* g, bool trace, bool stored_moves) {
    GjSelectArgs A{};
    A.agent = g->prm.agent; A.K = g->K; A.stride = g->stride; A.levels = g->levels; A.n_vars = g->n_vars;
    A.late_size = g->late_size; A.noop = g->noop; A.n_groups = g->groups.n_groups;
    A.moves = stored_moves ? g->moves : nullptr;
    A.M = g->mover; A.seed = g->prm.seed; A.step = g->step; A.island_base = g->island_base;
    A.cand_scores = g->cand_scores;
    A.cur = g->cur; A.cur_score = g->cur_score; A.best = g->best; A.best_score = g->best_score;
    A.dirty = g->dirty; A.late = g->late; A.late_head = g->late_head; A.late_len = g->late_len;
    A.counters = g->counters;
    A.tabu_bits = g->tabu_bits; A.tabu_words_per_island = g->tabu_words; A.tabu_word_off = g->tabu_word_off;
    A.tabu_ring_old = g->tabu_ring[g->step & 1]; A.tabu_ring_new = g->tabu_ring[(g->step + 1) & 1];
    A.tabu_ring_per_island = g->tabu_ring_len; A.tabu_ring_off = g->tabu_ring_off;
    A.tabu_size = g->tabu_size; A.tabu_fill = g->tabu_fill;
    A.stale = g->ds.stale; A.defer_top = g->scoring_mode == GJ_SCORING_DELTA ? 1 : 0;
    A.work_count = g->work_count;
    A.sa_temp = g->sa_temp; A.sa = g->sa;
    A.compare_to_global = g->prm.agent == GJ_AGENT_TABU_SEARCH ? g->prm.compare_to_global : 1;
    A.gbest = g->gbest; A.gbest_score = g->gbest_score; A.gver = g->gver; A.gseen = g->gseen;
    A.chain_mode = g->chain ? 1 : 0;
    A.selected_out = trace ? g->selected : nullptr; A.accepted_out = trace ? g->accepted : nullptr;
    A.aux_out = trace ? g->trace_aux : nullptr;
    return A;
}

static GjStepCtx make_step_ctx(gj_islands* g) {
    GjStepCtx C{};
    C.seed = g->prm.seed; C.step = g->step; C.I = g->I; C.K = g->K; C.island_base = g->island_base;
    C.noop = g->noop; C.stride = g->stride; C.symmetric = g->p->symmetric_D ? 1 : 0;
    C.tabu_bits = g->tabu_bits; C.tabu_words_per_island = g->tabu_words; C.tabu_word_off = g->tabu_word_off;
    return C;
}

static size_t warp_eval_smem(const GjProblemDev& P, int warps, bool with_clone) {
    return (size_t)warps * (size_t)(P.bm_words + P.desc_words + P.asc_words + (with_clone ? P.n_vars : 0)) * 4;
}


// k_refresh for every island (early exit for islands whose state is current)
static gj_status launch_refresh(gj_islands* g, cudaStream_t st, bool update_top) {
    if (g->vrp_chain) return GJ_OK;        // the chain kernel rebuilds its own route index (stale flag)
    const GjProblemDev& P = g->p->dev;
    const size_t smem = 0;
    gj_status rc;
    if (P.kind >= GJ_VRP) {
        const size_t vsmem = gj_vrp_smem_bytes(P.n_entities, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
        if ((rc = opt_in_smem(k_vrp_state, vsmem))) return rc;
        k_vrp_state<<<g->I, kVrpWarps * 32, vsmem, st>>>(P, g->stride, g->cur, g->cur_score, g->ds, g->vs,
                                                         update_top ? 1 : 0, g->best, g->best_score, g->dirty);
        GJ_LAUNCH_CHECK();
        return GJ_OK;
    }
    if (P.kind == GJ_NQUEENS) {
        if ((rc = opt_in_smem(k_refresh<GJ_NQUEENS>, smem))) return rc;
        k_refresh<GJ_NQUEENS><<<g->I, g->n_vars > 4096 ? 1024 : 256, smem, st>>>(P, g->stride, g->cur, g->cur_score, g->ds, update_top ? 1 : 0,
                                                     g->best, g->best_score, g->dirty);
    } else {
        if ((rc = opt_in_smem(k_refresh<GJ_TSP>, smem))) return rc;
        k_refresh<GJ_TSP><<<g->I, g->n_vars > 4096 ? 1024 : 256, smem, st>>>(P, g->stride, g->cur, g->cur_score, g->ds, update_top ? 1 : 0,
                                                 g->best, g->best_score, g->dirty);
    }
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

// Delta scoring of the step's neighbourhood: generation + evaluation fused, then the full
// evaluator over whatever the delta evaluator queued.
static gj_status launch_score_delta(gj_islands* g, cudaStream_t st, bool trace) {
    const GjProblemDev& P = g->p->dev;
    const int64_t total = (int64_t)g->I * g->K;
    const GjStepCtx C = make_step_ctx(g);
    const unsigned grid = (unsigned)((total + 255) / 256);
    GjMove* moves_out = trace ? g->moves : nullptr;
    gj_status rc;
    if (P.kind >= GJ_VRP) {
        k_score_delta_vrp<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(P, g->groups, g->mover, C, g->cur, g->ds, g->vs,
                                                                           g->cand_scores, g->worklist, g->work_count, moves_out);
        GJ_LAUNCH_CHECK();
        if (g->delta_may_fallback) {
            const size_t vsmem = gj_vrp_smem_bytes(P.n_entities, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
            if ((rc = opt_in_smem(k_score_fallback_vrp, vsmem))) return rc;
            const unsigned fgrid = (unsigned)std::min<int64_t>(total, 148 * 4);
            k_score_fallback_vrp<<<fgrid, kVrpWarps * 32, vsmem, st>>>(P, g->groups, g->mover, C, g->cur, g->worklist,
                                                                      g->work_count, g->cand_scores);
            GJ_LAUNCH_CHECK();
        }
        return GJ_OK;
    }
    if (P.kind == GJ_NQUEENS)
        k_score_delta<GJ_NQUEENS><<<grid, 256, 0, st>>>(P, g->groups, g->mover, C, g->cur, g->ds, g->cand_scores,
                                                       g->worklist, g->work_count, moves_out);
    else
        k_score_delta<GJ_TSP><<<grid, 256, 0, st>>>(P, g->groups, g->mover, C, g->cur, g->ds, g->cand_scores,
                                                   g->worklist, g->work_count, moves_out);
    GJ_LAUNCH_CHECK();
    if (g->delta_may_fallback) {
        const int warps = (int)std::min<size_t>(kWarps, (220 * 1024) / warp_eval_smem(P, 1, true));
        if (warps < 1) return gj_fail(GJ_ERR_UNSUPPORTED, "instance too large for the shared-memory candidate clone");
        const size_t smem = warp_eval_smem(P, warps, true);
        const unsigned fgrid = (unsigned)std::min<int64_t>((total + warps - 1) / warps, 148 * 4);
        if (P.kind == GJ_NQUEENS) {
            if ((rc = opt_in_smem(k_score_fallback_warp<GJ_NQUEENS>, smem))) return rc;
            k_score_fallback_warp<GJ_NQUEENS><<<fgrid, warps * 32, smem, st>>>(P, g->groups, g->mover, C, g->cur,
                                                                               g->worklist, g->work_count, g->cand_scores);
        } else {
            if ((rc = opt_in_smem(k_score_fallback_warp<GJ_TSP>, smem))) return rc;
            k_score_fallback_warp<GJ_TSP><<<fgrid, warps * 32, smem, st>>>(P, g->groups, g->mover, C, g->cur,
                                                                           g->worklist, g->work_count, g->cand_scores);
        }
        GJ_LAUNCH_CHECK();
    }
    return GJ_OK;
}

static gj_status ls_one_step(gj_islands* g, cudaStream_t st, bool trace) {
    const GjProblemDev& P = g->p->dev;
    const int64_t total = (int64_t)g->I * g->K;
    const bool delta = g->scoring_mode == GJ_SCORING_DELTA;
    gj_status rc;
    if (g->fused) {
        // generation, delta scoring, selection, apply, exact re-score and tabu update in one kernel
        if ((rc = gj_prof_begin(g, st))) return rc;
        if ((rc = g->ts_fast ? gj_launch_tsfast_step(g, st, trace) : gj_launch_fused_step(g, st, trace))) return rc;
        if ((rc = gj_prof_end(g, st))) return rc;
        g->step += 1;
        return GJ_OK;
    }
    if (!delta) {
        k_gen_moves<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, st>>>(
            P, g->groups, g->mover, g->prm.seed, g->step, g->I, g->K, g->island_base, g->tabu_bits,
            g->tabu_words, g->tabu_word_off, g->moves);
        GJ_LAUNCH_CHECK();
        if ((rc = gj_prof_begin(g, st))) return rc;
        if ((rc = launch_score_moves(g, st))) return rc;
        if ((rc = gj_prof_end(g, st))) return rc;
    } else {
        if ((rc = gj_prof_begin(g, st))) return rc;
        if ((rc = launch_score_delta(g, st, trace))) return rc;
        if ((rc = gj_prof_end(g, st))) return rc;
    }
    size_t smem = (size_t)g->n_vars * 4;
    if ((rc = opt_in_smem(k_select, smem))) return rc;
    // row copies dominate for long solutions: a wider CTA then
    k_select<<<g->I, g->n_vars > 4096 ? 1024 : 256, smem, st>>>(P, g->groups, gj_make_select_args(g, trace, !delta));
    GJ_LAUNCH_CHECK();
    // delta mode: exact re-score of accepted neighbours + state rebuild, then update_top_individual
    if (delta && (rc = launch_refresh(g, st, true))) return rc;
    g->step += 1;
    return GJ_OK;
}

static gj_status ls_migrate_pack(gj_islands* g, cudaStream_t st, int parity) {
    k_migrate_pack<<<g->I, 128, 0, st>>>(g->cur, g->cur_score, g->stride, g->n_vars, g->mailbox, parity);
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

static gj_status ls_migrate_recv(gj_islands* g, cudaStream_t st, int parity) {
    k_migrate_recv<<<g->I, 128, 0, st>>>(g->prm.agent, g->levels, g->stride, g->n_vars, g->late_size,
                                        g->mailbox, g->cur, g->cur_score, g->dirty, g->late,
                                        g->late_head, g->late_len, g->ds.stale, parity);
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

// The send half of one migration (agent_base.rs:161-183): even islands send what they hold, odd islands
// take their migrant FIRST and send what they hold afterwards -- a migrant an odd island accepts travels
// two hops in one migration.  Leaves the last island's migrant in slot I (what leaves the group).
gj_status gj_ls_migrate_pack(gj_islands* g, cudaStream_t st) {
    gj_status rc;
    if ((rc = ls_migrate_pack(g, st, 0))) return rc;
    if (g->I > 1) {
        if ((rc = ls_migrate_recv(g, st, 1))) return rc;
        if ((rc = ls_migrate_pack(g, st, 1))) return rc;
    }
    return GJ_OK;
}

// The receive half: the even islands (slot 0 = what entered the group) take their migrants.
gj_status gj_ls_migrate_recv(gj_islands* g, cudaStream_t st) {
    return ls_migrate_recv(g, st, 0);
}

static gj_status apply_pending_adoption(gj_islands* g, cudaStream_t st);

gj_status gj_ls_global_top(gj_islands* g, cudaStream_t st) {
    // update_global_top (agent_base.rs:446-490).  Publish half: ONE CTA (arg-min over the agent tops
    // -> gbest row, score, version).  Adopt half: once per published version and island
    // (gj_adopt_decide) -- inside the next launch for fused islands and chains (staging phase),
    // right away by k_apply_adoption for the other paths.  O(I) work in total.
    k_global_top<<<1, (g->n_vars > 4096 || g->I > 2048) ? 1024 : 256, 0, st>>>(
        g->I, g->levels, g->stride, g->n_vars, g->best, g->best_score, g->gbest, g->gbest_score, g->gver,
        GjGtopLean{g->top_is_cur, g->ds.stale, g->dirty, g->ds.edge, g->ds.raw, g->fused_lean ? g->gedge : nullptr, g->graw,
                   g->gedge_ver});
    GJ_LAUNCH_CHECK();
    if (!g->fused && !g->chain) {
        k_apply_adoption<<<g->I, 128, 0, st>>>(gj_make_select_args(g, false, false));
        GJ_LAUNCH_CHECK();
    }
    // VRP chains: the route index of the global top is built lazily, by the first launch that follows
    // (gj_launch_vrp_chains; the kernel returns at once when the index is current) -- a publication here
    // and a better top merged from another GPU right after cost one re-index, not two.  On the side stream
    // (next to the in-GPU ring migration) it is built right away: that is what the overlap is for.
    if (g->vrp_chain && g->ga_side && st == g->ga_side) {
        gj_status rc = gj_launch_vrp_gindex(g, st);
        if (rc) return rc;
    }
    return GJ_OK;
}

// fused islands: pending adoptions must land before anyone looks at (or exports) current solutions
static gj_status apply_pending_adoption(gj_islands* g, cudaStream_t st) {
    if (!g->fused && !g->chain) return GJ_OK;          // the other paths adopt right after publishing
    k_apply_adoption<<<g->I, 128, 0, st>>>(gj_make_select_args(g, false, false));
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

// n consecutive LateAcceptance steps of every chain in one launch
static gj_status launch_chain_steps(gj_islands* g, int n, cudaStream_t st, bool trace) {
    const GjProblemDev& P = g->p->dev;
    GjChainArgs A{};
    A.I = g->I; A.stride = g->stride; A.n_vars = g->n_vars; A.late_size = g->late_size; A.noop = g->noop;
    A.n_groups = g->groups.n_groups; A.symmetric = g->p->symmetric_D ? 1 : 0; A.island_base = g->island_base;
    A.M = g->mover; A.seed = g->prm.seed; A.step0 = g->step; A.n_steps = n;
    A.cur = g->cur; A.cur_score = g->cur_score; A.best = g->best; A.best_score = g->best_score;
    A.dirty = g->dirty; A.late = g->late; A.late_head = g->late_head; A.late_len = g->late_len;
    A.counters = g->counters;
    A.ctabu = g->ctabu; A.ctabu_words_per_island = g->ctabu_words; A.ctabu_off = g->ctabu_off;
    A.tabu_size = g->tabu_size;
    A.agent = g->prm.agent; A.sa_temp = g->sa_temp; A.sa = g->sa;
    A.gbest = g->gbest; A.gbest_score = g->gbest_score; A.gver = g->gver; A.gseen = g->gseen;
    if (trace) A.trace_aux = g->trace_aux;
    // trace (n == 1): the step's move / score / decision land where the per-step path puts them
    if (trace) { A.trace_moves = g->moves; A.trace_scores = g->cand_scores; A.trace_accept = g->accepted; }
    gj_status rc = g->vrp_chain ? gj_launch_vrp_chains(g, A, st) : gj_launch_la_chains(g, A, st);
    if (rc) return rc;
    g->step += (uint64_t)n;
    return GJ_OK;
}

// side stream + events (shared with the GeneticAlgorithm generation, gj_islands_ga.cu), created on first use;
// GJ_LS_OVERLAP=0 switches the overlap off
static bool ls_side_ready(gj_islands* g) {
    if (g->ga_side_state == 0) {
        const char* e = getenv("GJ_LS_OVERLAP");
        g->ga_side_state = -1;
        if (!(e && e[0] == '0')) {
            bool ok = cudaStreamCreateWithFlags(&g->ga_side, cudaStreamNonBlocking) == cudaSuccess;
            for (int i = 0; ok && i < 4; ++i) ok = cudaEventCreateWithFlags(&g->ga_ev[i], cudaEventDisableTiming) == cudaSuccess;
            if (ok) g->ga_side_state = 1;
        }
    }
    return g->ga_side_state == 1;
}

static gj_status ls_step(gj_islands* g, int64_t n_steps, cudaStream_t st) {
    gj_status rc;
    if (g->chain) {
        // launches of up to steps_to_send steps; migration and the global top between launches
        int64_t left = n_steps;
        while (left > 0) {
            const int per_launch = g->prm.chain_steps_per_launch > 0 ? g->prm.chain_steps_per_launch : (g->vrp_chain ? 32 : 8);
            const int n = (int)std::min<int64_t>(std::min<int64_t>(left, per_launch),
                                                 std::max<int64_t>(1, g->steps_to_send));
            if ((rc = gj_prof_begin(g, st))) return rc;
            if ((rc = launch_chain_steps(g, n, st, false))) return rc;
            if ((rc = gj_prof_end(g, st))) return rc;
            left -= n;
            g->steps_to_send -= n;
            // update_global_top reads and writes the agent tops / the global top only, the ring migration the
            // current solutions only: on VRP chains, where publishing includes one CTA re-indexing the new
            // global top (80 us), the two run side by side on two streams
            const bool migrate_now = g->steps_to_send <= 0 && !g->external_ring;
            const bool side = migrate_now && g->vrp_chain && ls_side_ready(g);
            if (side) {
                GJ_CUDA_TRY(cudaEventRecord(g->ga_ev[0], st));
                GJ_CUDA_TRY(cudaStreamWaitEvent(g->ga_side, g->ga_ev[0], 0));
                if ((rc = gj_ls_global_top(g, g->ga_side))) return rc;
                GJ_CUDA_TRY(cudaEventRecord(g->ga_ev[1], g->ga_side));
            }
            if (g->steps_to_send <= 0) {
                if (!g->external_ring) {
                    if ((rc = gj_ls_migrate_pack(g, st))) return rc;
                    k_migrate_wrap<<<8, 256, 0, st>>>(g->I, g->stride, g->mailbox);
                    GJ_LAUNCH_CHECK();
                    if ((rc = gj_ls_migrate_recv(g, st))) return rc;
                }
                g->steps_to_send = std::max<int64_t>(1, g->prm.migration_frequency);
            }
            if (side) GJ_CUDA_TRY(cudaStreamWaitEvent(st, g->ga_ev[1], 0));
            else if ((rc = gj_ls_global_top(g, st))) return rc;
        }
        return GJ_OK;
    }
    for (int64_t s = 0; s < n_steps; ++s) {
        if ((rc = ls_one_step(g, st, false))) return rc;
        // agent_base.rs:161-183: every migration_frequency steps send + receive
        g->steps_to_send -= 1;
        if (g->steps_to_send <= 0) {
            if (!g->external_ring) {
                if ((rc = gj_ls_migrate_pack(g, st))) return rc;
                k_migrate_wrap<<<8, 256, 0, st>>>(g->I, g->stride, g->mailbox);
                GJ_LAUNCH_CHECK();
                if ((rc = gj_ls_migrate_recv(g, st))) return rc;
            }
            g->steps_to_send = std::max<int64_t>(1, g->prm.migration_frequency);
        }
        // agent_base.rs:185 (the fixed-point step publishes from its last island, inside the launch)
        if (!g->ts_fast && (rc = gj_ls_global_top(g, st))) return rc;
        // delta mode: islands that received a migrant / adopted the global best rebuild their state
        if (g->scoring_mode == GJ_SCORING_DELTA && !g->fused && (rc = launch_refresh(g, st, false))) return rc;
    }
    return GJ_OK;
}

// ---- C ABI -----------------------------------------------------------------------------------------------

extern "C" gj_status gj_islands_create(gj_problem* p, const gj_agent_params* params,
                                       const double* initial, gj_islands** out) {
    if (!p || !params || !out) return gj_fail(GJ_ERR_INVALID, "null argument");
    *out = nullptr;
    if (params->n_islands < 1) return gj_fail(GJ_ERR_INVALID, "n_islands must be >= 1");
    GJ_CUDA_TRY(cudaSetDevice(p->device));
    switch (params->agent) {
        case GJ_AGENT_TABU_SEARCH:
        case GJ_AGENT_LATE_ACCEPTANCE:
        case GJ_AGENT_SIMULATED_ANNEALING: return ls_create(p, params, initial, out);
        case GJ_AGENT_GENETIC_ALGORITHM: return gj_ga_create(p, params, initial, out);
        default: return gj_fail(GJ_ERR_INVALID, "unknown agent kind");
    }
}

extern "C" void gj_islands_destroy(gj_islands* g) { delete g; }

extern "C" gj_status gj_islands_step(gj_islands* g, int64_t n_steps, void* stream) {
    if (!g || n_steps < 0) return gj_fail(GJ_ERR_INVALID, "bad argument");
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    if (g->prm.agent == GJ_AGENT_GENETIC_ALGORITHM) return gj_ga_step(g, n_steps, (cudaStream_t)stream);
    return ls_step(g, n_steps, (cudaStream_t)stream);
}

extern "C" gj_status gj_islands_set_accomplish_rate(gj_islands* g, double accomplish_rate) {
    if (!g) return gj_fail(GJ_ERR_INVALID, "null handle");
    g->sa.inv_rate = 1.0 - accomplish_rate;             // agent_base.rs:544
    return GJ_OK;
}

extern "C" gj_status gj_islands_trace_aux(gj_islands* g, int32_t island, double* out) {
    if (!g || !out || island < 0 || island >= g->I) return gj_fail(GJ_ERR_INVALID, "bad argument");
    for (int i = 0; i < 5; ++i) out[i] = 0.0;
    if (!g->trace_aux) return GJ_OK;
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    GJ_CUDA_TRY(cudaDeviceSynchronize());
    GJ_CUDA_TRY(cudaMemcpy(out, g->trace_aux + (size_t)island * 5, 5 * sizeof(double), cudaMemcpyDeviceToHost));
    return GJ_OK;
}

extern "C" gj_status gj_islands_trace_tabu(gj_islands* g, int32_t island, int32_t group, int32_t* ids,
                                           int32_t capacity, int32_t* fill, int32_t* size) {
    if (!g || island < 0 || island >= g->I || group < 0 || group >= g->groups.n_groups)
        return gj_fail(GJ_ERR_INVALID, "bad island / group");
    if (fill) *fill = 0;
    if (size) *size = 0;
    if (g->prm.tabu_entity_rate == 0.0 || !g->tabu_size) return GJ_OK;
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    GJ_CUDA_TRY(cudaDeviceSynchronize());
    const int glen = (int)g->p->groups[group].size();
    const int T = gj_tabu_deque_size(g->prm.tabu_entity_rate, glen);
    if (size) *size = T;
    std::vector<int32_t> ring(T), out;
    int n_fill = 0;
    if (g->chain) {
        // chain layout per group: bits [W + 1] | ring [T] (circular) | head, fill
        if (!g->ctabu) return GJ_OK;
        std::vector<int32_t> off(g->groups.n_groups);
        GJ_CUDA_TRY(cudaMemcpy(off.data(), g->ctabu_off, off.size() * 4, cudaMemcpyDeviceToHost));
        const int W = (glen + 31) / 32;
        const uint32_t* base = g->ctabu + (size_t)island * g->ctabu_words + off[group] + W + 1;
        std::vector<int32_t> raw(T + 2);
        GJ_CUDA_TRY(cudaMemcpy(raw.data(), base, raw.size() * 4, cudaMemcpyDeviceToHost));
        const int head = raw[T];
        n_fill = raw[T + 1];
        for (int r = 0; r < n_fill; ++r) out.push_back(raw[((head - 1 - r) % T + T) % T]);
    } else {
        // TabuSearch / GeneticAlgorithm layout: rank-indexed (slot 0 = newest), double-buffered by step parity
        std::vector<int32_t> off(g->groups.n_groups);
        GJ_CUDA_TRY(cudaMemcpy(off.data(), g->tabu_ring_off, off.size() * 4, cudaMemcpyDeviceToHost));
        GJ_CUDA_TRY(cudaMemcpy(&n_fill, g->tabu_fill + (size_t)island * g->groups.n_groups + group, 4, cudaMemcpyDeviceToHost));
        GJ_CUDA_TRY(cudaMemcpy(ring.data(), g->tabu_ring[g->step & 1] + (size_t)island * g->tabu_ring_len + off[group],
                               (size_t)T * 4, cudaMemcpyDeviceToHost));
        out.assign(ring.begin(), ring.begin() + n_fill);
    }
    if (fill) *fill = n_fill;
    if (ids) for (int i = 0; i < n_fill && i < capacity; ++i) ids[i] = out[i];
    return GJ_OK;
}

// Which kernels a step of this group runs (decided at creation from the agent, the model, the
// instance size and move_probas); for logs and tests.
extern "C" const char* gj_islands_step_path(const gj_islands* g) {
    if (!g) return "";
    if (g->prm.agent == GJ_AGENT_GENETIC_ALGORITHM) return "ga";
    if (g->vrp_chain) return "vrp_chain";
    if (g->chain) return "chain";
    if (g->fused) return g->ts_fast ? "fused_fixed" : (g->fused_lean ? "fused_lean" : "fused");
    if (g->scoring_mode == GJ_SCORING_DELTA) return g->p->dev.kind >= GJ_VRP ? "vrp_delta" : "delta";
    return "full";
}

extern "C" gj_status gj_islands_set_external_ring(gj_islands* g, int32_t on, int32_t island_base) {
    if (!g) return gj_fail(GJ_ERR_INVALID, "null handle");
    g->external_ring = on != 0;
    g->island_base = island_base;
    g->ga_moves_step = -1;          // moves generated ahead were keyed by the old island ids
    return GJ_OK;
}

extern "C" gj_status gj_islands_stats(gj_islands* g, int64_t* candidates, int64_t* steps, int64_t* accepted) {
    if (!g) return gj_fail(GJ_ERR_INVALID, "null handle");
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    unsigned long long h[4];
    GJ_CUDA_TRY(cudaMemcpy(h, g->counters, sizeof(h), cudaMemcpyDeviceToHost));
    if (candidates) *candidates = (int64_t)h[0];
    if (steps) *steps = (int64_t)h[1];
    if (accepted) *accepted = (int64_t)h[2];
    return GJ_OK;
}

static gj_status fetch_individual(gj_islands* g, const int32_t* d_row, const double* d_score,
                                  double* vars, double* score) {
    std::vector<int32_t> row(g->n_vars);
    double sc[GJ_MAX_LEVELS];
    GJ_CUDA_TRY(cudaMemcpy(row.data(), d_row, (size_t)g->n_vars * 4, cudaMemcpyDeviceToHost));
    GJ_CUDA_TRY(cudaMemcpy(sc, d_score, sizeof(sc), cudaMemcpyDeviceToHost));
    if (vars) for (int i = 0; i < g->n_vars; ++i) vars[i] = (double)row[i];
    if (score) for (int l = 0; l < g->levels; ++l) score[l] = sc[l];
    return GJ_OK;
}

extern "C" gj_status gj_islands_best(gj_islands* g, int32_t island, double* vars, double* score) {
    if (!g || island >= g->I) return gj_fail(GJ_ERR_INVALID, "bad island");
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    GJ_CUDA_TRY(cudaDeviceSynchronize());
    if (island < 0) {
        // the group's global_top_individual; before the first step it is still the stub
        gj_status rc = (g->prm.agent == GJ_AGENT_GENETIC_ALGORITHM) ? gj_ga_global_top(g, g->p->stream)
                                                                   : gj_ls_global_top(g, g->p->stream);
        if (rc) return rc;
        if (g->scoring_mode == GJ_SCORING_DELTA && !g->fused && (rc = launch_refresh(g, g->p->stream, false))) return rc;
        GJ_CUDA_TRY(cudaStreamSynchronize(g->p->stream));
        return fetch_individual(g, g->gbest, g->gbest_score, vars, score);
    }
    return fetch_individual(g, g->best + (size_t)island * g->stride,
                            g->best_score + (size_t)island * GJ_MAX_LEVELS, vars, score);
}

extern "C" gj_status gj_islands_current(gj_islands* g, int32_t island, double* vars, double* score) {
    if (!g || island < 0 || island >= g->I) return gj_fail(GJ_ERR_INVALID, "bad island");
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    GJ_CUDA_TRY(cudaDeviceSynchronize());
    if (g->prm.agent == GJ_AGENT_GENETIC_ALGORITHM) return gj_ga_current(g, island, vars, score);
    {
        gj_status rc = apply_pending_adoption(g, g->p->stream);
        if (rc) return rc;
        GJ_CUDA_TRY(cudaStreamSynchronize(g->p->stream));
    }
    return fetch_individual(g, g->cur + (size_t)island * g->stride,
                            g->cur_score + (size_t)island * GJ_MAX_LEVELS, vars, score);
}

extern "C" int64_t gj_islands_migrant_bytes(const gj_islands* g) {
    if (!g) return 0;
    return (int64_t)g->migrants * (int64_t)((size_t)g->stride * 4 + GJ_MAX_LEVELS * 8);
}

// the group's outgoing migrants, packed in place: *d_slot = device address of what leaves the group
// (gj_islands_migrant_bytes long).  Used by gj_islands_export_migrants and by the peer ring (gj_ring.cu),
// which stores it straight into the next rank's inbox.
gj_status gj_islands_pack_outgoing(gj_islands* g, cudaStream_t st, const unsigned char** d_slot) {
    gj_status rc;
    if (g->prm.agent == GJ_AGENT_GENETIC_ALGORITHM) return gj_ga_pack_outgoing(g, st, d_slot);
    if ((rc = apply_pending_adoption(g, st))) return rc;
    if ((rc = gj_ls_migrate_pack(g, st))) return rc;
    const size_t sb = (size_t)g->stride * 4 + GJ_MAX_LEVELS * 8;
    *d_slot = g->mailbox + (size_t)g->I * sb;
    return GJ_OK;
}

extern "C" gj_status gj_islands_export_migrants(gj_islands* g, void* d_buffer, void* stream) {
    if (!g || !d_buffer) return gj_fail(GJ_ERR_INVALID, "bad argument");
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    cudaStream_t st = (cudaStream_t)stream;
    gj_status rc;
    if (g->prm.agent == GJ_AGENT_GENETIC_ALGORITHM) return gj_ga_export(g, d_buffer, st);
    if ((rc = apply_pending_adoption(g, st))) return rc;
    if ((rc = gj_ls_migrate_pack(g, st))) return rc;
    const size_t sb = (size_t)g->stride * 4 + GJ_MAX_LEVELS * 8;
    GJ_CUDA_TRY(cudaMemcpyAsync(d_buffer, g->mailbox + (size_t)g->I * sb, sb, cudaMemcpyDeviceToDevice, st));
    return GJ_OK;
}

extern "C" gj_status gj_islands_import_migrants(gj_islands* g, const void* d_buffer, void* stream) {
    if (!g || !d_buffer) return gj_fail(GJ_ERR_INVALID, "bad argument");
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (g->prm.agent == GJ_AGENT_GENETIC_ALGORITHM) return gj_ga_import(g, d_buffer, st);
    const size_t sb = (size_t)g->stride * 4 + GJ_MAX_LEVELS * 8;
    GJ_CUDA_TRY(cudaMemcpyAsync(g->mailbox, d_buffer, sb, cudaMemcpyDeviceToDevice, st));
    gj_status rc;
    if ((rc = gj_ls_migrate_recv(g, st))) return rc;
    if (g->scoring_mode == GJ_SCORING_DELTA && !g->fused) return launch_refresh(g, st, false);
    return GJ_OK;
}

extern "C" gj_status gj_islands_trace_step(gj_islands* g, int32_t island, uint64_t* offsets,
                                           uint64_t* var_ids, double* values, int64_t delta_capacity,
                                           int32_t* move_kinds, int32_t* move_desc, double* scores,
                                           int64_t* selected, int32_t* accepted) {
    if (!g || island < 0 || island >= g->I) return gj_fail(GJ_ERR_INVALID, "bad island");
    if (g->prm.agent == GJ_AGENT_GENETIC_ALGORITHM) return gj_fail(GJ_ERR_UNSUPPORTED, "trace_step is for TS / LA islands");
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    cudaStream_t st = g->p->stream;
    const GjProblemDev& P = g->p->dev;
    const int K = g->K;
    // base of the island BEFORE the step
    {
        gj_status rc0 = apply_pending_adoption(g, st);
        if (rc0) return rc0;
        GJ_CUDA_TRY(cudaStreamSynchronize(st));
    }
    std::vector<int32_t> base(g->n_vars);
    GJ_CUDA_TRY(cudaMemcpy(base.data(), g->cur + (size_t)island * g->stride, (size_t)g->n_vars * 4, cudaMemcpyDeviceToHost));
    // scratch buffers are grow-only members of nothing: scoped holders free them on every return path
    struct DevBuf {
        void* p = nullptr;
        ~DevBuf() { if (p) cudaFree(p); }
    } b_base, b_offs, b_ids, b_vals;
    GJ_CUDA_TRY(cudaMalloc(&b_base.p, (size_t)g->n_vars * 4));
    int32_t* d_base = (int32_t*)b_base.p;
    GJ_CUDA_TRY(cudaMemcpy(d_base, base.data(), (size_t)g->n_vars * 4, cudaMemcpyHostToDevice));
    gj_status rc = g->chain ? launch_chain_steps(g, 1, st, true) : ls_one_step(g, st, true);
    if (rc) return rc;
    GJ_CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<GjMove> mv(K);
    GJ_CUDA_TRY(cudaMemcpy(mv.data(), g->moves + (size_t)island * K, (size_t)K * sizeof(GjMove), cudaMemcpyDeviceToHost));
    std::vector<uint64_t> offs(K + 1, 0);
    for (int j = 0; j < K; ++j) {
        const GjMove& m = mv[j];
        uint64_t n = 0;
        if (m.kind == GJ_MOVE_NULL) n = 0;
        else if (m.kind == 2) n = 2ull * m.k;
        else if (m.kind <= 3) n = m.k;
        else n = (uint64_t)(std::abs(m.a[0] - m.a[1]) + 1);
        offs[j + 1] = offs[j] + n;
        if (move_kinds) move_kinds[j] = m.kind;
        if (move_desc) {
            int32_t* d = move_desc + (size_t)j * 20;
            d[0] = m.kind; d[1] = m.group; d[2] = m.k; d[3] = 0;
            for (int i = 0; i < GJ_MOVE_MAXK; ++i) { d[4 + i] = m.a[i]; d[12 + i] = m.v[i]; }
        }
    }
    if ((int64_t)offs[K] > delta_capacity) return gj_fail(GJ_ERR_INVALID, "delta_capacity too small");
    GJ_CUDA_TRY(cudaMalloc(&b_offs.p, (size_t)(K + 1) * 8));
    GJ_CUDA_TRY(cudaMalloc(&b_ids.p, (size_t)(offs[K] + 1) * 8));
    GJ_CUDA_TRY(cudaMalloc(&b_vals.p, (size_t)(offs[K] + 1) * 8));
    uint64_t* d_offs = (uint64_t*)b_offs.p; uint64_t* d_ids = (uint64_t*)b_ids.p; double* d_vals = (double*)b_vals.p;
    GJ_CUDA_TRY(cudaMemcpy(d_offs, offs.data(), (size_t)(K + 1) * 8, cudaMemcpyHostToDevice));
    k_expand_moves<<<K, 128, 0, st>>>(P, g->groups, d_base, g->moves + (size_t)island * K, K, g->noop, d_offs, d_ids, d_vals);
    GJ_LAUNCH_CHECK();
    GJ_CUDA_TRY(cudaStreamSynchronize(st));
    if (offsets) std::copy(offs.begin(), offs.end(), offsets);
    if (var_ids && offs[K]) GJ_CUDA_TRY(cudaMemcpy(var_ids, d_ids, (size_t)offs[K] * 8, cudaMemcpyDeviceToHost));
    if (values && offs[K]) GJ_CUDA_TRY(cudaMemcpy(values, d_vals, (size_t)offs[K] * 8, cudaMemcpyDeviceToHost));
    if (scores) GJ_CUDA_TRY(cudaMemcpy(scores, g->cand_scores + (size_t)island * K * g->levels, (size_t)K * g->levels * 8, cudaMemcpyDeviceToHost));
    long long sel = 0; int acc = 0;
    GJ_CUDA_TRY(cudaMemcpy(&sel, g->selected + island, 8, cudaMemcpyDeviceToHost));
    GJ_CUDA_TRY(cudaMemcpy(&acc, g->accepted + island, 4, cudaMemcpyDeviceToHost));
    if (selected) *selected = sel;
    if (accepted) *accepted = acc;
    // the trace bypasses migration / global-top bookkeeping of gj_islands_step on purpose
    return GJ_OK;
}
