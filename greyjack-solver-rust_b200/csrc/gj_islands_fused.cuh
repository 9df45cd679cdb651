// gj_islands_fused.cuh -- one TabuSearch / LateAcceptance step of an island as ONE kernel
// (included by gj_islands.cu).  One CTA owns one island for the whole step:
//
//   P0  stage the island in shared memory: current solution, value counts (rebuilt with shared
//       atomics, one per thread), tabu bitmap; re-score the solution if it was replaced by a
//       migrant / the global best since the last step
//   P1  every thread generates + delta-scores K / blockDim neighbours straight from registers
//       (Mover::do_move -> gj_delta.cuh), keeping its first minimum
//   P1b neighbours the delta evaluator does not cover: full evaluator, one warp each
//   P2  CTA arg-min (first minimum, tabu_search_base.rs:166-171), acceptance rule
//       (tabu_search_base.rs:174 | late_acceptance_base.rs:196-213)
//   P3  accepted: apply the winning move, write the solution back, full re-score in the
//       reference's summation order, update_top_individual (agent_base.rs:220-224)
//   P4  tabu deque update (mover.rs:75-96)
//
// Nothing but the winning solution ever leaves the SM: neighbours, their moves and their scores
// live in registers.  Migration and the global best stay separate (tiny) kernels because they
// couple islands.
#pragma once

struct GjFusedArgs {
    GjSelectArgs A;             // same state as the unfused select
    GjDeltaState S;
    int symmetric;
    int n_clone;                // shared-memory solution clones available to the full evaluator
    int lean;                   // 1: long solutions -- only the solution itself is staged in shared
                                // memory; edge lengths and the tabu table stay in HBM / L2 and there are
                                // no value counts (the host guarantees every move takes the fast path)
    double* scores_out;         // trace only: [I][K][levels]
    GjMove* moves_out;          // trace only
    int* worklist;              // [I][K]
    long long* phase_clocks;    // development aid (GJ_PHASE_TIMING=1): [I][8] clock64 at phase ends
    // lean layout: edge lengths / unrounded terms published with the global top (k_global_top), valid for
    // version *gedge_ver; top_is_cur[island]: the agent's top row is its current row
    int* top_is_cur; const double* gedge; const double* graw; const int* gedge_ver;
};

struct GjFusedSmem {
    int32_t* t;                 // [n_pad] current solution
    int32_t* cnt;               // [cnt_stride]
    uint32_t* bits;             // [tabu_words_pad] tabu membership snapshot
    uint32_t* bm;               // [n_clone][words] bitmaps of the full evaluator
    int32_t* clone;             // [n_clone][n_pad]
    double* edge;               // [n + 1] TSP: edge[i] = D[t[i-1]][t[i]], depot at both ends
};


__device__ __forceinline__ GjFusedSmem gj_fused_carve(unsigned char* smem, int n_vars, int cnt_stride,
                                                      int tabu_words, int words, int n_clone, bool lean) {
    const size_t n_pad = ((size_t)n_vars + 3) & ~(size_t)3;
    GjFusedSmem s;
    size_t o = 0;
    if (lean) {                 // solution + one distinct-count bitmap; the caller points edge / bits at HBM
        s.t = (int32_t*)(smem + o) + 4; o += (n_pad + 8) * 4;
        s.bm = (uint32_t*)(smem + o);
        s.cnt = nullptr; s.bits = nullptr; s.clone = nullptr; s.edge = nullptr;
        return s;
    }
    // the solution sits 16 bytes into its slot: t[-1] and t[n] are sentinels (TSP: the depot)
    s.t = (int32_t*)(smem + o) + 4; o += (n_pad + 8) * 4;
    s.cnt = (int32_t*)(smem + o); o += (size_t)cnt_stride * 4;
    s.bits = (uint32_t*)(smem + o); o += (((size_t)tabu_words + 3) & ~(size_t)3) * 4;
    s.bm = (uint32_t*)(smem + o); o += (size_t)n_clone * (size_t)words * 4;
    s.clone = (int32_t*)(smem + o); o += (size_t)n_clone * n_pad * 4;
    o = (o + 15) & ~(size_t)15;
    s.edge = (double*)(smem + o);
    return s;
}

// Value counts of the staged solution (shared atomics), all threads.
template <int KIND>
__device__ __forceinline__ void gj_fused_counts(const GjProblemDev& P, const GjFusedSmem& s, int cnt_stride) {
    if (!s.cnt) return;                              // lean layout: no move needs them
    for (int i = threadIdx.x; i < cnt_stride; i += blockDim.x) s.cnt[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < P.n_vars; i += blockDim.x) {
        const int v = s.t[i];
        atomicAdd(&s.cnt[v - P.val_lo], 1);
        if constexpr (KIND == GJ_NQUEENS) {
            const int col = P.column_id[i];
            atomicAdd(&s.cnt[32 * P.bm_words + (col + v - P.desc_lo)], 1);
            atomicAdd(&s.cnt[32 * (P.bm_words + P.desc_words) + (col - v - P.asc_lo)], 1);
        }
    }
    __syncthreads();
}

// TSP: lengths of the staged tour's n + 1 edges (depot -> s0, s0 -> s1, .., s_last -> depot),
// one gather per thread.  They serve the full evaluation below and, in P1, every "removed edge"
// of a neighbour (gj_delta.cuh) -- half of the matrix gathers of a delta evaluation.
__device__ __forceinline__ void gj_fused_edges(const GjProblemDev& P, const GjFusedSmem& s) {
    const int n = P.n_vars;
    const size_t L = (size_t)P.n_locations;
    for (int i = threadIdx.x; i <= n; i += blockDim.x) {
        const int a = (i == 0) ? 0 : s.t[i - 1];
        const int b = (i == n) ? 0 : s.t[i];
        s.edge[i] = __ldg(&P.D[(size_t)a * L + (size_t)b]);
    }
    __syncthreads();
}

// Lean layout (edges persistent in HBM): after an accepted fast-path move only the edges it touched
// are refreshed -- a two-stop swap gathers its <= 4 edges, a 2-opt reverses the interior edge run
// in place (symmetric matrix: the same lengths in reverse order) and gathers the two boundary edges,
// an insertion re-gathers its segment -- instead of all n + 1 random matrix reads.
__device__ __forceinline__ void gj_fused_edges_after_move(const GjProblemDev& P, const GjFusedSmem& s,
                                                          const GjGroups& G, const GjMove& m, bool noop_quirk) {
    if (m.kind == GJ_MOVE_NULL) return;
    if (noop_quirk && (m.kind == 3 || (m.kind == 2 && m.k == 2))) return;
    const int n = P.n_vars;
    const size_t L = (size_t)P.n_locations;
    const int4 gi = G.info[m.group];
    const int c0 = gi.x + m.a[0] * gi.y, c1 = gi.x + m.a[1] * gi.y;
    const int p = min(c0, c1), q = max(c0, c1);
    auto regather = [&](int i) {                    // s.t carries the sentinels t[-1] = t[n] = 0
        if (i >= 0 && i <= n) s.edge[i] = __ldg(&P.D[(size_t)s.t[i - 1] * L + (size_t)s.t[i]]);
    };
    if (m.kind == 1) {
        if (threadIdx.x == 0) regather(p);
        if (threadIdx.x == 1) regather(p + 1);
        if (threadIdx.x == 2) regather(q);
        if (threadIdx.x == 3) regather(q + 1);
    } else if (m.kind == 5) {
        const int len = q - p;                      // interior edges p+1 .. q
        for (int j = threadIdx.x; j < len / 2; j += blockDim.x) {
            const double x = s.edge[p + 1 + j], y = s.edge[q - j];
            s.edge[p + 1 + j] = y; s.edge[q - j] = x;
        }
        if (threadIdx.x == 0) regather(p);
        if (threadIdx.x == 1) regather(q + 1);
    } else {
        for (int i = p + (int)threadIdx.x; i <= q + 1; i += blockDim.x) regather(i);
    }
    __syncthreads();
}

// FULL evaluation of the staged solution by the whole CTA -> unweighted terms raw[0..1]
// (counts and, for TSP, edges must be current).  TSP: dup count from the counts; tour length
// either as per-thread partial sums (tree) or, with exact sums, folded by one thread strictly in
// the reference's order (tsp ISC :76-80): ((0 + D[0][s0]) + D[s_last][0]) + fold_{i>=1} D[s_{i-1}][s_i].
#define GJ_FOLD_CHUNK 1024
template <int KIND>
__device__ __forceinline__ void gj_fused_full_eval(const GjProblemDev& P, const GjFusedSmem& s,
                                                   int cnt_stride, int* iscratch,
                                                   double* dscratch, double* raw /*shared [2]*/,
                                                   double* fold_stage = nullptr /*shared [GJ_FOLD_CHUNK], lean layout*/) {
    int u = 0;
    if (s.cnt) {
        for (int k = threadIdx.x; k < cnt_stride; k += blockDim.x) u += (s.cnt[k] > 0) ? 1 : 0;
    } else {
        // lean layout: distinct values through a bitmap (as the warp evaluators do)
        const int words = cnt_stride >> 5;
        for (int w = threadIdx.x; w < words; w += blockDim.x) s.bm[w] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < P.n_vars; i += blockDim.x) {
            const unsigned b = (unsigned)(s.t[i] - P.val_lo);
            atomicOr(&s.bm[b >> 5], 1u << (b & 31));
        }
        __syncthreads();
        for (int w = threadIdx.x; w < words; w += blockDim.x) u += __popc(s.bm[w]);
    }
    const int uniq = gj_block_sum(u, iscratch);
    if constexpr (KIND == GJ_NQUEENS) {
        // (N - |rows|) + (N - |desc|) + (N - |asc|): integers, exact whatever the grouping
        if (threadIdx.x == 0) { raw[0] = (double)(3 * P.n_vars - uniq); raw[1] = 0.0; }
        __syncthreads();
        return;
    } else {
        const int n = P.n_vars;
        if (!P.exact_sums) {
            double acc = 0.0;
            for (int i = threadIdx.x; i <= n; i += blockDim.x) acc += s.edge[i];
            const double dist = gj_block_sum(acc, dscratch);
            if (threadIdx.x == 0) { raw[0] = (double)(n - uniq); raw[1] = dist; }
            __syncthreads();
            return;
        }
        if (fold_stage) {
            // lean layout: the edge lengths live in HBM.  One thread folding them straight from there pays
            // a cache round trip per unrolled batch (640 us for 20 000 edges); instead the CTA stages
            // GJ_FOLD_CHUNK of them at a time in shared memory and thread 0 folds the chunk from there --
            // the same sequential order, bound by the DADD chain alone.
            double fold = 0.0;
            for (int base = 1; base < n; base += GJ_FOLD_CHUNK) {
                const int m = min(GJ_FOLD_CHUNK, n - base);
                __syncthreads();
                for (int i = threadIdx.x; i < m; i += blockDim.x) fold_stage[i] = s.edge[base + i];
                __syncthreads();
                if (threadIdx.x == 0) {
#pragma unroll 8
                    for (int i = 0; i < m; ++i) fold = fold + fold_stage[i];
                }
            }
            if (threadIdx.x == 0) {
                double sample_distance = 0.0;
                sample_distance += s.edge[0];
                sample_distance += s.edge[n];
                sample_distance += fold;
                raw[0] = (double)(n - uniq); raw[1] = sample_distance;
            }
            __syncthreads();
            return;
        }
        if (threadIdx.x == 0) {
            double fold = 0.0;
#pragma unroll 8
            for (int i = 1; i < n; ++i) fold = fold + s.edge[i];
            double sample_distance = 0.0;
            sample_distance += s.edge[0];                 // D[0][s0]
            sample_distance += s.edge[n];                 // D[s_last][0]
            sample_distance += fold;
            raw[0] = (double)(n - uniq); raw[1] = sample_distance;
        }
        __syncthreads();
    }
}

template <int KIND>
__device__ __forceinline__ void gj_fused_combine(const GjProblemDev& P, const double* raw, int d_uniq,
                                                 double d_dist, GjScore& s) {
    s.v[0] = 0.0; s.v[1] = 0.0; s.v[2] = 0.0;
    if constexpr (KIND == GJ_NQUEENS) gj_combine_nqueens(P, raw[0] - (double)d_uniq, s.v);
    else gj_combine_tsp(P, true, raw[0] - (double)d_uniq, raw[1] + d_dist, s.v);
}

// Ordering key of a score level under ScoreTrait::round (math_utils.rs:10-13):
// round(v) = floor(v) + floor((v - floor(v)) * 10^p) / 10^p is strictly increasing in the integer
// pair (floor(v), floor(frac * 10^p)), so comparing floor(v) * 10^p + floor(frac * 10^p) (exact in
// f64 for |v| < 2^53 / 10^p) orders neighbours exactly like their rounded scores -- without the
// two divisions per neighbour.  Only the winner is actually rounded.
__device__ __forceinline__ double gj_round_key(double v, double mult) {
    if (mult == 0.0) return v;
    const double fl = floor(v);
    return fl * mult + floor((v - fl) * mult);
}

template <int LV>
struct GjBest {
    double key[LV];
    double val[LV];                 // unrounded score of the same neighbour
    int idx;
};

// -1 / 0 / +1: lexicographic order of two key vectors
template <int LV>
__device__ __forceinline__ int gj_key_cmp(const double* a, const double* b) {
    int c = 0;
#pragma unroll
    for (int l = LV - 1; l >= 0; --l) {
        if (a[l] < b[l]) c = -1;
        else if (a[l] > b[l]) c = 1;
    }
    return c;
}

// keeps the first minimum: replace on strictly smaller key, or equal key and smaller index
template <int LV>
__device__ __forceinline__ void gj_best_merge(GjBest<LV>& mine, const GjBest<LV>& o) {
    const int c = gj_key_cmp<LV>(o.key, mine.key);
    const bool take = (o.idx >= 0) && (mine.idx < 0 || c < 0 || (c == 0 && o.idx < mine.idx));
    if (take) {
#pragma unroll
        for (int l = 0; l < LV; ++l) { mine.key[l] = o.key[l]; mine.val[l] = o.val[l]; }
        mine.idx = o.idx;
    }
}

template <int LV>
__device__ __forceinline__ GjBest<LV> gj_best_shfl_xor(const GjBest<LV>& b, int o) {
    GjBest<LV> r;
    r.idx = __shfl_xor_sync(GJ_FULL_MASK, b.idx, o);
#pragma unroll
    for (int l = 0; l < LV; ++l) {
        r.key[l] = __shfl_xor_sync(GJ_FULL_MASK, b.key[l], o);
        r.val[l] = __shfl_xor_sync(GJ_FULL_MASK, b.val[l], o);
    }
    return r;
}

template <int KIND, int NT>
__global__ void __launch_bounds__(NT, (NT <= 512) ? 1024 / NT : 1)
k_ls_step_fused(GjProblemDev P, GjGroups G, GjFusedArgs F) {
    constexpr int LV = (KIND == GJ_NQUEENS) ? 1 : 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ GjBest<LV> sh_bestw[32];
    __shared__ int sh_sel0[NT], sh_sel1[NT], sh_selinfo[NT];   // ids the last chunk's moves selected
    __shared__ int sh_iscratch[32];
    __shared__ double sh_dscratch[32];
    __shared__ double sh_raw[2];
    __shared__ double sh_fold[GJ_FOLD_CHUNK];          // staging of the exact fold (lean layout: edges in HBM)
    __shared__ int sh_accept, sh_best, sh_nwork;
    __shared__ int sh_scan[NT];
    __shared__ GjMove sh_mv[4];

    const GjSelectArgs& A = F.A;
    const int island = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int K = A.K, levels = A.levels, n = P.n_vars;
    const int words = P.bm_words + P.desc_words + P.asc_words;
    const int cnt_stride = 32 * words;
    GjFusedSmem s = gj_fused_carve(smem_raw, n, cnt_stride, A.tabu_words_per_island, words, F.n_clone, F.lean != 0);
    if (F.lean) {
        s.edge = F.S.edge + (size_t)island * (size_t)(n + 1);
        s.bits = A.tabu_bits ? A.tabu_bits + (size_t)island * A.tabu_words_per_island : nullptr;
    }
    int32_t* cur_row = A.cur + (size_t)island * A.stride;
    double* raw_g = F.S.raw + (size_t)island * GJ_MAX_LEVELS;

    auto stamp = [&](int k) {
        if (F.phase_clocks && tid == 0) F.phase_clocks[(size_t)island * 8 + k] = clock64();
    };
    stamp(0);
    // ---- P0: stage ---------------------------------------------------------------------------
    // The island's solution row and tabu table travel global -> shared as two TMA bulk copies
    // issued by one thread (rows and tables are 16-byte aligned and padded, see ls_create).
    __shared__ __align__(8) uint64_t sh_mbar;
    __shared__ int sh_adopt;
    if (tid == 0) {
        gj_mbar_init(&sh_mbar, 1);
        sh_nwork = 0;
        // update_global_top, adopt half: the global top published after the previous step replaces
        // this island's solution when it beats the island's own top
        sh_adopt = gj_adopt_decide(A, island) ? 1 : 0;
    }
    __syncthreads();
    const bool adopted = sh_adopt != 0;
    if (tid == 0) {
        const uint32_t row_bytes = (uint32_t)(((n + 3) & ~3) * 4);
        const uint32_t tabu_bytes = (A.tabu_bits && !F.lean) ? (uint32_t)(A.tabu_words_per_island * 4) : 0u;
        gj_mbar_expect_tx(&sh_mbar, row_bytes + tabu_bytes);
        gj_tma_load_1d(s.t, adopted ? A.gbest : cur_row, row_bytes, &sh_mbar);
        if (tabu_bytes)
            gj_tma_load_1d(s.bits, A.tabu_bits + (size_t)island * A.tabu_words_per_island, tabu_bytes, &sh_mbar);
    }
    if (tid == 0) {                                 // one poller (spinning warps burn issue slots); the rest sleep
        gj_mbar_wait(&sh_mbar, 0);
        s.t[-1] = 0; s.t[n] = 0;                    // depot before the first and after the last stop
    }
    __syncthreads();
    if (adopted)
        for (int i = tid; i < n; i += blockDim.x) cur_row[i] = s.t[i];
    if constexpr (KIND == GJ_TSP) {
        // lean layout: the adopted row came with its edge lengths and score terms (see k_global_top) --
        // a sequential copy instead of n + 1 gathers from a matrix that does not fit in L2
        if (adopted && F.lean && F.gedge && *F.gedge_ver == *A.gver) {
            for (int i = tid; i <= n; i += blockDim.x) s.edge[i] = F.gedge[i];
            if (tid == 0) { raw_g[0] = F.graw[0]; raw_g[1] = F.graw[1]; F.S.stale[island] = 0; }
            __syncthreads();
        }
    }
    gj_fused_counts<KIND>(P, s, cnt_stride);
    const int state_stale = F.S.stale[island];
    // edge lengths: shared memory does not outlive the launch, so they are re-gathered every step;
    // the lean layout keeps them in HBM, where they stay valid until the solution changes
    if constexpr (KIND == GJ_TSP) if (!F.lean || state_stale) gj_fused_edges(P, s);
    if (state_stale) {                             // replaced by a migrant / the global best
        gj_fused_full_eval<KIND>(P, s, cnt_stride, sh_iscratch, sh_dscratch, sh_raw, F.lean ? sh_fold : nullptr);
        if (tid == 0) { raw_g[0] = sh_raw[0]; raw_g[1] = sh_raw[1]; F.S.stale[island] = 0; }
    } else if (tid == 0) {
        sh_raw[0] = raw_g[0]; sh_raw[1] = raw_g[1];
    }
    __syncthreads();
    const double raw0 = sh_raw[0], raw1 = sh_raw[1];
    const double raw[2] = {raw0, raw1};
    const uint32_t* bits = A.tabu_bits ? s.bits : nullptr;
    auto make_move = [&](int c) -> GjMove {
        return gj_generate_move(P, G, A.M, A.seed, (uint32_t)(A.island_base + island), A.step,
                                (uint32_t)c, bits, A.tabu_word_off);
    };

    stamp(1);
    // ---- P1: generate + delta-score ------------------------------------------------------------
    GjBest<LV> mine;
    mine.idx = -1;
#pragma unroll
    for (int l = 0; l < LV; ++l) { mine.key[l] = 0.0; mine.val[l] = 0.0; }
    int* worklist = F.worklist + (size_t)island * K;
    const int n_chunks = (K + blockDim.x - 1) / blockDim.x;
    sh_selinfo[tid] = 0;
    // `in_order`: neighbours offered in increasing index order (the main loop).  ScoreTrait::round is
    // monotone per level, so a neighbour that is >= the thread's best on EVERY level cannot round to a
    // lexicographically smaller score, and on a tie the earlier index stays: the key arithmetic is only
    // needed for the others.  (A lexicographic test on the unrounded values would not do: two different
    // upper levels may round to the same value and leave the decision to a lower one.)
    auto offer = [&](const GjScore& sc, int c, bool in_order) {
        if (F.scores_out) {
            GjScore r = sc;
            gj_score_round(r, P);                  // agent_base.rs:311-314
            for (int l = 0; l < levels; ++l) F.scores_out[((size_t)island * K + c) * levels + l] = r.v[l];
        }
        if (in_order && mine.idx >= 0) {
            bool not_below = true;
#pragma unroll
            for (int l = 0; l < LV; ++l) not_below = not_below && (sc.v[l] >= mine.val[l]);
            if (not_below) return;
        }
        GjBest<LV> o;
        o.idx = c;
#pragma unroll
        for (int l = 0; l < LV; ++l) {
            o.val[l] = sc.v[l];
            o.key[l] = gj_round_key(sc.v[l], P.round_mult[l]);
        }
        gj_best_merge<LV>(mine, o);
    };
    for (int c = tid; c < K; c += blockDim.x) {
        const GjMove m = make_move(c);
        if (F.moves_out) F.moves_out[(size_t)island * K + c] = m;
        if (c >= (n_chunks - 1) * (int)blockDim.x && m.kind != GJ_MOVE_NULL) {
            // remembered for the tabu update (P4): at most two ids, else regenerated there
            int sel[GJ_MOVE_MAXK];
            const int cnt = gj_move_selected(m, sel);
            sh_sel0[tid] = sel[0]; sh_sel1[tid] = sel[1];
            sh_selinfo[tid] = (cnt + 1) | ((int)m.group << 8);
        }
        int d_uniq = 0; double d_dist = 0.0;
        bool ok;
        if constexpr (KIND == GJ_NQUEENS) {
            ok = gj_nqueens_move_delta(P, G, m, A.noop != 0, s.t, s.cnt, d_uniq);
        } else {
            GjTspBase B{s.t, n, P.D, (size_t)P.n_locations, s.edge, true};
            ok = gj_tsp_move_delta(P, G, m, A.noop != 0, F.symmetric != 0, B, s.cnt, d_uniq, d_dist);
        }
        if (!ok) { worklist[atomicAdd(&sh_nwork, 1)] = c; continue; }
        GjScore sc;
        gj_fused_combine<KIND>(P, raw, d_uniq, d_dist, sc);
        offer(sc, c, true);
    }
    __syncthreads();

    // ---- P1b: full evaluator for the queued neighbours (warp per neighbour, smem clone) ---------
    const int n_work = sh_nwork;
    if (n_work > 0 && warp < F.n_clone) {
        const size_t n_pad = ((size_t)n + 3) & ~(size_t)3;
        uint32_t* bm = s.bm + (size_t)warp * words;
        int32_t* cand = s.clone + (size_t)warp * n_pad;
        for (int w = warp; w < n_work; w += F.n_clone) {
            const int c = worklist[w];
            if (lane == 0) sh_mv[warp] = make_move(c);
            for (int i = lane; i < n; i += 32) cand[i] = s.t[i];
            __syncwarp();
            const GjMove m = sh_mv[warp];
            gj_apply_move(P, m, G, true, A.noop != 0, lane, 32,
                          [&](int id) { return s.t[id]; }, [&](int id, int v) { cand[id] = v; });
            __syncwarp();
            GjSrcI32 src{cand};
            GjScore sc;
            sc.v[0] = sc.v[1] = sc.v[2] = 0.0;
            if constexpr (KIND == GJ_NQUEENS) {
                gj_combine_nqueens(P, gj_nqueens_eval_warp(P, src, bm, lane), sc.v);
            } else {
                double dup, dist;
                gj_tsp_eval_warp(P, src, bm, lane, dup, dist);
                gj_combine_tsp(P, true, dup, dist, sc.v);
            }
            if (lane == 0) offer(sc, c, false);
            __syncwarp();
        }
    }

    stamp(2);
    // ---- P2: first minimum + acceptance ------------------------------------------------------------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gj_best_merge<LV>(mine, gj_best_shfl_xor<LV>(mine, o));
    if (lane == 0) sh_bestw[warp] = mine;
    __syncthreads();
    if (warp == 0) {
        GjBest<LV> b = sh_bestw[lane < nwarps ? lane : 0];
        if (lane >= nwarps) b.idx = -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gj_best_merge<LV>(b, gj_best_shfl_xor<LV>(b, o));
        if (lane == 0) sh_bestw[0] = b;
    }
    __syncthreads();
    if (tid == 0) {
        const int bi = sh_bestw[0].idx;
        GjScore b;
        b.v[0] = b.v[1] = b.v[2] = 0.0;
#pragma unroll
        for (int l = 0; l < LV; ++l) b.v[l] = sh_bestw[0].val[l];
        gj_score_round(b, P);                       // agent_base.rs:311-314 (the winner only; see gj_round_key)
        GjScore cur = gj_load_score(A.cur_score + (size_t)island * GJ_MAX_LEVELS, levels);
        bool accept;
        if (A.agent == GJ_AGENT_TABU_SEARCH) {
            accept = gj_score_le(b, cur, levels);                       // tabu_search_base.rs:174
        } else if (A.agent == GJ_AGENT_SIMULATED_ANNEALING) {
            accept = gj_sa_step_accept(A, island, b, cur);              // simulated_annealing_base.rs:198-233
        } else {
            double* late = A.late + (size_t)island * A.late_size * GJ_MAX_LEVELS;   // late_acceptance_base.rs:196-213
            int head = A.late_head[island], len = A.late_len[island];
            GjScore late_native = cur;
            if (len > 0) late_native = gj_load_score(late + (size_t)((head + len - 1) % A.late_size) * GJ_MAX_LEVELS, levels);
            accept = gj_score_le(b, late_native, levels) || gj_score_le(b, cur, levels);
            if (accept) {
                head = (head + A.late_size - 1) % A.late_size;
                for (int l = 0; l < GJ_MAX_LEVELS; ++l) late[(size_t)head * GJ_MAX_LEVELS + l] = b.v[l];
                A.late_head[island] = head; A.late_len[island] = min(len + 1, A.late_size);
            }
        }
        if (accept)
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) A.cur_score[(size_t)island * GJ_MAX_LEVELS + l] = b.v[l];
        sh_accept = accept ? 1 : 0;
        sh_best = bi;
        if (accept) sh_mv[0] = make_move(bi);
        if (A.selected_out) { A.selected_out[island] = bi; A.accepted_out[island] = accept ? 1 : 0; }
        atomicAdd(&A.counters[0], (unsigned long long)K);
        if (island == 0) atomicAdd(&A.counters[1], 1ull);
        if (accept) atomicAdd(&A.counters[2], 1ull);
    }
    __syncthreads();

    stamp(3);
    // ---- P3: apply, write back, exact re-score, update_top_individual ----------------------------------
    if (sh_accept) {
        const GjMove m = sh_mv[0];
        // the global row still holds the base: read it, write the staged copy
        gj_apply_move(P, m, G, true, A.noop != 0, tid, blockDim.x,
                      [&](int id) { return cur_row[id]; }, [&](int id, int v) { s.t[id] = v; });
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) cur_row[i] = s.t[i];
        gj_fused_counts<KIND>(P, s, cnt_stride);
        if constexpr (KIND == GJ_TSP) {
            if (F.lean) gj_fused_edges_after_move(P, s, G, m, A.noop != 0);
            else gj_fused_edges(P, s);
        }
        gj_fused_full_eval<KIND>(P, s, cnt_stride, sh_iscratch, sh_dscratch, sh_raw, F.lean ? sh_fold : nullptr);
        if (tid == 0) {
            raw_g[0] = sh_raw[0]; raw_g[1] = sh_raw[1];
            GjScore sc;
            const double r[2] = {sh_raw[0], sh_raw[1]};
            gj_fused_combine<KIND>(P, r, 0, 0.0, sc);
            gj_score_round(sc, P);
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) A.cur_score[(size_t)island * GJ_MAX_LEVELS + l] = (l < levels) ? sc.v[l] : 0.0;
            A.dirty[island] = 1;
        }
        __syncthreads();
    }
    gj_update_top(island, levels, A.stride, n, A.cur, A.cur_score, A.best, A.best_score, A.dirty);
    if (F.top_is_cur && tid == 0) {
        // update_top_individual copies on `<=`: after it, equal scores <=> the top row is the current row
        bool same = true;
        for (int l = 0; l < levels; ++l)
            same = same && A.best_score[(size_t)island * GJ_MAX_LEVELS + l] == A.cur_score[(size_t)island * GJ_MAX_LEVELS + l];
        F.top_is_cur[island] = same ? 1 : 0;
    }

    stamp(4);
    // ---- P4: tabu deque update (see k_select) --------------------------------------------------------------
    if (A.tabu_bits) {
        uint32_t* bits_rw = A.tabu_bits + (size_t)island * A.tabu_words_per_island;
        const int32_t* ring_old_island = A.tabu_ring_old + (size_t)island * A.tabu_ring_per_island;
        int32_t* ring_new_island = A.tabu_ring_new + (size_t)island * A.tabu_ring_per_island;
        for (int g = 0; g < A.n_groups; ++g) {
            const int T = A.tabu_size[g];
            const int32_t* ring_old = ring_old_island + A.tabu_ring_off[g];
            int32_t* ring_new = ring_new_island + A.tabu_ring_off[g];
            const int glen = G.offsets[g + 1] - G.offsets[g];
            const int fill_old = A.tabu_fill[island * A.n_groups + g];
            int collected = 0;
            for (int chunk = n_chunks - 1; chunk >= 0 && collected < T; --chunk) {
                const int j = chunk * blockDim.x + tid;
                int sel[GJ_MOVE_MAXK]; int cnt = 0;
                if (j < K) {
                    const int info = sh_selinfo[tid];
                    if (chunk == n_chunks - 1 && (info & 0xff) <= 3) {
                        // remembered from P1 (cnt + 1 in the low byte; 0 = null move)
                        if ((info & 0xff) != 0 && (info >> 8) == g) {
                            cnt = (info & 0xff) - 1;
                            sel[0] = sh_sel0[tid]; sel[1] = sh_sel1[tid];
                        }
                    } else {
                        const GjMove m = make_move(j);
                        if (m.kind != GJ_MOVE_NULL && m.group == g) cnt = gj_move_selected(m, sel);
                    }
                }
                int total;
                const int incl = gj_block_scan_incl(cnt, sh_scan, &total);
                const int after = total - incl;
#pragma unroll
                for (int i = 0; i < GJ_MOVE_MAXK; ++i) {
                    if (i < cnt) {
                        const int rank = collected + after + (cnt - 1 - i);
                        if (rank < T) ring_new[rank] = sel[i];
                    }
                }
                __syncthreads();
                collected += total;
            }
            for (int r = collected + tid; r < T; r += blockDim.x) {
                const int rho = r - collected;
                if (rho < fill_old) ring_new[r] = ring_old[rho];
            }
            const int fill = min(T, fill_old + collected);
            if (tid == 0) A.tabu_fill[island * A.n_groups + g] = fill;
            gj_tabu_table_rebuild(bits_rw + A.tabu_word_off[g], glen, ring_new, fill, sh_scan);
        }
    }
    stamp(5);
}
