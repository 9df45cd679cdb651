// gj_islands_ga.cu -- the GeneticAlgorithm agent on the device
// (greyjack/src/agents/metaheuristic_bases/genetic_algorithm_base.rs, Agent::step_plain
// agent_base.rs:273-298).  One generation of every island per launch sequence:
//
//   k_ga_offspring  select_p_best x2, cross, Mover::do_move(plain) x2, fix_variables (:141-187)
//   plain scorer    request_score_plain on the offspring (PSC semantics) + round
//   k_ga_replace    build_updated_population: candidate i vs random p-worst native (:198-213)
//   k_ga_rank       population.sort() (agent_base.rs:149-151) by counting -> rank table
//   k_ga_top        update_top_individual (agent_base.rs:220-224)
//   k_ga_migrate_*  send_updates / receive_updates for Population agents (:337-341, 405-412)
//
// VRP models (one CTA scores one candidate) run the same generation without ever materialising the
// offspring -- an offspring is one parent plus one plain-form move (see k_ga_plan):
//
//   k_ga_gen_moves           the moves of the NEXT generation, on the side stream (RNG + tabu deque only)
//   k_ga_plan                select_p_best x2 + cross -> parent slot of every offspring
//   k_ga_score_planned_vrp   parent row -> shared memory, move applied in place, PSC score, round
//   k_ga_decide              build_updated_population: scores / sources of the new population
//   k_ga_copy_planned || k_ga_rank (side stream)   rows of the survivors  ||  population.sort()
//   k_ga_finish              order from rank, update_top_individual, one-island global top
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <memory>

#include "gj_islands_dev.cuh"

struct GjGaArgs {
    int I, pop, half, n_cand, stride, n_vars, levels, noop;
    double crossover_probability, p_best_rate;
    uint64_t seed, step;
    int island_base;
};

__device__ __forceinline__ int gj_ga_p_rank(GjPhilox& rng, double p_best_rate, int pop, bool worst,
                                            double* trace /* nullable: {p, last_top, id} */) {
    // select_p_best / select_p_worst (genetic_algorithm_base.rs:83-103):
    // p ~ U(1e-6, p_best_rate); last_top = ceil(p * pop); id ~ U[0, last_top) | U[pop-last_top, pop)
    const double p = 0.000001 + gj_rng_f64(rng) * (p_best_rate - 0.000001);
    int last_top = (int)ceil(p * (double)pop);
    last_top = max(1, min(last_top, pop));
    const int id = (int)gj_rng_below(rng, (uint32_t)last_top);
    if (trace) { trace[0] = p; trace[1] = (double)last_top; trace[2] = (double)id; }
    return worst ? (pop - last_top + id) : id;
}

// One CTA per offspring.
__global__ void __launch_bounds__(128)
k_ga_offspring(GjProblemDev P, GjGroups G, GjMoverParams M, GjGaArgs A,
               const int32_t* __restrict__ pop_rows, const int* __restrict__ order,
               int32_t* __restrict__ cand_rows, GjMove* __restrict__ moves,
               const uint32_t* __restrict__ tabu_bits, int tabu_words_per_island,
               const int32_t* __restrict__ tabu_word_off, double* __restrict__ trace_sel) {
    __shared__ int sh_parent;
    __shared__ GjMove sh_move;
    const int island = blockIdx.x / A.n_cand;
    const int c = blockIdx.x % A.n_cand;
    const int q = c >> 1, child = c & 1;
    if (threadIdx.x == 0) {
        GjPhilox rng;
        gj_rng_init(rng, A.seed, (uint32_t)(A.island_base + island), (uint32_t)A.step,
                    (uint32_t)(A.step >> 32), 0x40000000u + (uint32_t)q);
        // trace record of pair q (written by its first child): {p1, last_top1, id1, p2, last_top2, id2,
        // u_cross, w_raw (-1 = no crossover)}
        double* tr = (trace_sel && child == 0) ? trace_sel + ((size_t)island * A.half + q) * 8 : nullptr;
        int r1 = gj_ga_p_rank(rng, A.p_best_rate, A.pop, false, tr);
        int r2 = gj_ga_p_rank(rng, A.p_best_rate, A.pop, false, tr ? tr + 3 : nullptr);
        // cross (:105-134): ONE weight for every gene (vec![sample(); n] evaluates the RNG once);
        // integer variables get rint(w) in {0, 1}, so the children are the parents, possibly
        // swapped.  w ~ U[0,1]; rint ties (0.5) go to ceil.
        const double u_cross = gj_rng_f64(rng);
        double w_raw = -1.0;
        if (u_cross <= A.crossover_probability) {
            w_raw = gj_rng_f64(rng);
            const double w = gj_rint(w_raw);
            if (w == 0.0) { int t = r1; r1 = r2; r2 = t; }
        }
        if (tr) { tr[6] = u_cross; tr[7] = w_raw; }
        sh_parent = order[(size_t)island * A.pop + (child == 0 ? r1 : r2)];
        // Mover::do_move with the agent's tabu deque (genetic_algorithm_base.rs:148-155, mover.rs:75-96):
        // every offspring of the generation sees the deque as the generation found it; it advances once
        // per generation (k_ga_tabu_update), like a TabuSearch step's neighbourhood
        const uint32_t* bits = tabu_bits ? tabu_bits + (size_t)island * tabu_words_per_island : nullptr;
        sh_move = gj_generate_move(P, G, M, A.seed, (uint32_t)(A.island_base + island), A.step,
                                   (uint32_t)c, bits, tabu_word_off);
        moves[(size_t)island * A.n_cand + c] = sh_move;
    }
    __syncthreads();
    const int32_t* parent = pop_rows + ((size_t)island * A.pop + sh_parent) * A.stride;
    int32_t* out = cand_rows + ((size_t)island * A.n_cand + c) * A.stride;
    for (int i = threadIdx.x; i < A.stride; i += blockDim.x) out[i] = (i < A.n_vars) ? parent[i] : 0;
    __syncthreads();
    // Mover::do_move(.., incremental = false) + fix_variables(changed columns)
    const GjMove m = sh_move;
    gj_apply_move(P, m, G, false, A.noop != 0, threadIdx.x, blockDim.x,
                  [&](int id) { return parent[id]; }, [&](int id, int v) { out[id] = v; });
}

// The GA mover's tabu deque after a generation: the ids its n_cand moves selected, in candidate order.
__global__ void __launch_bounds__(256)
k_ga_tabu_update(GjGroups G, int n_cand, int n_groups, const GjMove* __restrict__ moves, uint32_t* tabu_bits,
                 int tabu_words_per_island, const int32_t* tabu_word_off, const int32_t* ring_old,
                 int32_t* ring_new, int ring_per_island, const int32_t* ring_off, const int32_t* tabu_size,
                 int* tabu_fill) {
    __shared__ int sh_scan[256];
    const int island = blockIdx.x;
    const GjMove* mv = moves + (size_t)island * n_cand;
    gj_tabu_deque_advance(tabu_bits + (size_t)island * tabu_words_per_island,
                          ring_old + (size_t)island * ring_per_island, ring_new + (size_t)island * ring_per_island,
                          ring_off, tabu_size, tabu_word_off, tabu_fill + (size_t)island * n_groups, n_groups, G,
                          n_cand, [&](int j) { return mv[j]; }, sh_scan);
}

__global__ void k_round_scores(GjProblemDev P, double* scores, int64_t n) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        GjScore s;
        for (int l = 0; l < P.levels; ++l) s.v[l] = scores[j * P.levels + l];
        gj_score_round(s, P);
        for (int l = 0; l < P.levels; ++l) scores[j * P.levels + l] = s.v[l];
    }
}

static constexpr int kGaPairInts = 2 + 2 * GJ_MOVE_MAXPAIRS;

// Mover::do_move(plain) of a small move, literally: the reference's own sequence of assignments / swaps on
// the candidate (mover.rs:145-317), executed by ONE thread on a row it can read and write (`rd` / `wr`,
// e.g. the shared-memory copy).  `touch(col)` is called for every column the move may have changed, in
// emission order (repeats possible).  Same result as gj_small_move_pairs(incremental = false) without its
// scratch arrays -- that expansion kept one thread busy for ~10 000 cycles per offspring.
template <class Ids, class Rd, class Wr, class Touch>
__device__ __forceinline__ void gj_ga_small_move_inplace(const GjMove& m, Ids g, Rd rd, Wr wr, Touch touch) {
    const int k = m.k;
    auto swp = [&](int ca, int cb) { const int x = rd(ca), y = rd(cb); wr(ca, y); wr(cb, x); };
    switch (m.kind) {
        case 0:     // change_move
            for (int i = 0; i < k; ++i) wr(g[m.a[i]], m.v[i]);
            for (int i = 0; i < k; ++i) touch(g[m.a[i]]);
            break;
        case 1:     // swap_move: candidate.swap(c[i-1], c[i])
            for (int i = 1; i < k; ++i) swp(g[m.a[i - 1]], g[m.a[i]]);
            for (int i = 0; i < k; ++i) touch(g[m.a[i]]);
            break;
        case 2:     // swap_edges_move: edges.rotate_left(1), then swaps of the left ends and of the right ends
            for (int i = 1; i < k; ++i) {
                const int pa = m.a[i], pb = m.a[(i + 1 == k) ? 0 : i + 1];     // rotated edges i-1 and i
                swp(g[pa], g[pb]);
                swp(g[pa + 1], g[pb + 1]);
            }
            for (int i = 0; i < k; ++i) { touch(g[m.a[i]]); touch(g[m.a[i] + 1]); }
            break;
        case 3: {   // scramble_move: candidate.swap(native[i], scrambled[i])
            const int start = m.a[0];
            for (int i = 0; i < k; ++i) swp(g[start + i], g[start + m.v[i]]);
            for (int i = 0; i < k; ++i) touch(g[start + i]);
            break;
        }
        default: break;
    }
}


// Applies a planned move to a row copy: `pr` = the plan's pair list (shared memory), rd reads the parent.
template <class Rd, class Wr>
__device__ __forceinline__ void gj_ga_apply_planned(const GjProblemDev& P, const GjGroups& G, const GjMove& m,
                                                    const int32_t* pr, bool noop, Rd rd, Wr wr) {
    const int np = pr[0];
    if (np >= 0) {
        if (threadIdx.x == 0)
            for (int i = 0; i < np; ++i) wr(pr[2 + 2 * i], pr[3 + 2 * i]);
        return;
    }
    (void)noop;
    if (m.kind == GJ_MOVE_NULL) return;
    const int32_t* g = G.ids + G.offsets[m.group];          // the segment half of gj_apply_move
    int lo, hi;
    gj_segment_bounds(m, lo, hi);
    const int len = hi - lo + 1;
    for (int t = threadIdx.x; t < len; t += blockDim.x) {
        const int sl = gj_segment_src_slot(m, false, t, len);
        const int col = g[lo + t];
        wr(col, gj_fix_column(P, col, rd(g[lo + sl])));
    }
}

// ---- VRP models: offspring are never materialised -----------------------------------------------------
// cross() hands integer genes over whole (rint(w) in {0, 1}), so an offspring is ONE parent
// plus one plain-form move.  k_ga_plan draws parents and moves (one thread per offspring), the scorer
// below rebuilds the offspring in shared memory straight from the parent's row -- parents are the p-best
// slice of the population, i.e. L2-resident -- and only the offspring that survive build_updated_population
// are ever written (k_ga_decide + k_ga_copy_planned).  A generation moves 2 x 131 MB less through HBM than
// copy -> score -> copy.
__global__ void __launch_bounds__(128)
k_ga_plan(GjProblemDev P, GjGroups G, GjMoverParams M, GjGaArgs A, const int* __restrict__ order,
          int* __restrict__ parent_slot, GjMove* __restrict__ moves, const uint32_t* __restrict__ tabu_bits,
          int tabu_words_per_island, const int32_t* __restrict__ tabu_word_off, double* __restrict__ trace_sel) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)A.I * A.n_cand) return;
    const int island = (int)(t / A.n_cand);
    const int c = (int)(t % A.n_cand);
    const int q = c >> 1, child = c & 1;
    GjPhilox rng;
    gj_rng_init(rng, A.seed, (uint32_t)(A.island_base + island), (uint32_t)A.step,
                (uint32_t)(A.step >> 32), 0x40000000u + (uint32_t)q);
    double* tr = (trace_sel && child == 0) ? trace_sel + ((size_t)island * A.half + q) * 8 : nullptr;
    int r1 = gj_ga_p_rank(rng, A.p_best_rate, A.pop, false, tr);
    int r2 = gj_ga_p_rank(rng, A.p_best_rate, A.pop, false, tr ? tr + 3 : nullptr);
    const double u_cross = gj_rng_f64(rng);                       // same draws as k_ga_offspring
    double w_raw = -1.0;
    if (u_cross <= A.crossover_probability) {
        w_raw = gj_rng_f64(rng);
        const double w = gj_rint(w_raw);
        if (w == 0.0) { int tmp = r1; r1 = r2; r2 = tmp; }
    }
    if (tr) { tr[6] = u_cross; tr[7] = w_raw; }
    const int slot = order[(size_t)island * A.pop + (child == 0 ? r1 : r2)];
    parent_slot[t] = slot;
    if (!moves) return;                                // generated ahead of time (k_ga_gen_moves)
    const uint32_t* bits = tabu_bits ? tabu_bits + (size_t)island * tabu_words_per_island : nullptr;
    moves[t] = gj_generate_move(P, G, M, A.seed, (uint32_t)(A.island_base + island), A.step, (uint32_t)c, bits,
                                tabu_word_off);
}

// The moves of generation `step` depend on the RNG counters and on the mover's tabu deque only -- not on
// the population -- so they are generated on the side stream as soon as the deque of the generation
// before has advanced, off the critical path plan -> score -> replace -> sort.
__global__ void __launch_bounds__(128)
k_ga_gen_moves(GjProblemDev P, GjGroups G, GjMoverParams M, uint64_t seed, uint64_t step, int island_base, int I,
               int n_cand, GjMove* __restrict__ moves, const uint32_t* __restrict__ tabu_bits,
               int tabu_words_per_island, const int32_t* __restrict__ tabu_word_off) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)I * n_cand) return;
    const int island = (int)(t / n_cand);
    const int c = (int)(t % n_cand);
    const uint32_t* bits = tabu_bits ? tabu_bits + (size_t)island * tabu_words_per_island : nullptr;
    moves[t] = gj_generate_move(P, G, M, seed, (uint32_t)(island_base + island), step, (uint32_t)c, bits, tabu_word_off);
}

// request_score_plain on an offspring = parent row + move, one CTA per offspring, PSC semantics,
// rounded (agent_base.rs:284-287).  `cand_out` (trace only): the offspring row.
#ifndef GJ_GA_SCORE_MINBLOCKS
#define GJ_GA_SCORE_MINBLOCKS 5
#endif
__global__ void __launch_bounds__(kVrpWarps * 32, GJ_GA_SCORE_MINBLOCKS)
k_ga_score_planned_vrp(GjProblemDev P, GjGroups G, GjGaArgs A, const int32_t* __restrict__ pop_rows,
                       const int* __restrict__ parent_slot, const GjMove* __restrict__ moves,
                       int32_t* __restrict__ pairs, double* __restrict__ scores, int32_t* __restrict__ cand_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ GjMove sh_move;
    __shared__ int32_t sh_pairs[kGaPairInts];
    const int n = P.n_entities;
    GjVrpSmem s = gj_vrp_carve(smem_raw, n, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
    const int64_t j = blockIdx.x;
    const int island = (int)(j / A.n_cand);
#ifdef GJ_VRP_PHASE_CLOCKS
    GJ_PHASE_DECL;
#endif
    const int32_t* parent = pop_rows + ((size_t)island * A.pop + parent_slot[j]) * A.stride;
    if (threadIdx.x < (int)(sizeof(GjMove) / 4))
        reinterpret_cast<int32_t*>(&sh_move)[threadIdx.x] = reinterpret_cast<const int32_t*>(moves + j)[threadIdx.x];
    {
        constexpr int U = 8;
        for (int i0 = threadIdx.x; i0 < n; i0 += U * blockDim.x) {
            int2 pr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * blockDim.x;
                pr[u] = make_int2(0, 0);
                if (i < n) pr[u] = __ldg(reinterpret_cast<const int2*>(parent + 2 * i));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i < n) { s.veh[i] = (uint16_t)pr[u].x; s.cust[i] = pr[u].y; }
            }
        }
    }
    gj_vrp_eval_zero(P, s);
    __syncthreads();
    GJ_PHASE_MARK(0);
    // Mover::do_move(.., incremental = false) + fix_variables(changed columns).  A small move: thread 0 runs
    // the reference's own assignments / swaps on the shared-memory copy and leaves the changed (column,
    // value) pairs in HBM for k_ga_copy_planned (pairs[0] = count, -1 = segment move).  A segment move is
    // shifted along the parent's row by all threads.
    {
        auto rd_s = [&](int id) { return (id & 1) ? s.cust[id >> 1] : (int)s.veh[id >> 1]; };
        auto wr_s = [&](int id, int v) { if (id & 1) s.cust[id >> 1] = v; else s.veh[id >> 1] = (uint16_t)v; };
        const int kind = sh_move.kind;
        if (kind != GJ_MOVE_NULL && kind > 3) {
            if (threadIdx.x == 0) pairs[(size_t)j * kGaPairInts] = -1;
            sh_pairs[0] = -1;           // every thread writes the same value: no barrier needed before the call
            gj_ga_apply_planned(P, G, sh_move, sh_pairs, A.noop != 0, [&](int id) { return __ldg(parent + id); }, wr_s);
        } else if (threadIdx.x == 0) {
            int32_t* pg = pairs + (size_t)j * kGaPairInts;
            int np = 0;
            if (kind != GJ_MOVE_NULL) {
                const int4 info = G.info[sh_move.group];
                // uniform group: every column has the same bounds, values moved inside it (and the change
                // move's in-bounds draws) never clamp -- fix_variables is the identity
                const bool fix = info.z == 0;
                auto touch = [&](int col) {
                    int v = rd_s(col);                          // all swaps are done when touch() runs
                    if (fix) { v = gj_fix_column(P, col, v); wr_s(col, v); }
                    pg[2 + 2 * np] = col; pg[3 + 2 * np] = v;
                    ++np;
                };
                if (info.y != 0) gj_ga_small_move_inplace(sh_move, GjAffineIds{info.x, info.y}, rd_s, wr_s, touch);
                else gj_ga_small_move_inplace(sh_move, G.ids + G.offsets[sh_move.group], rd_s, wr_s, touch);
            }
            pg[0] = np;
        }
    }
    __syncthreads();
    GJ_PHASE_MARK(1);
    if (cand_out) {
        int32_t* out = cand_out + (size_t)j * A.stride;
        for (int i = threadIdx.x; i < n; i += blockDim.x) { out[2 * i] = s.veh[i]; out[2 * i + 1] = s.cust[i]; }
        __syncthreads();
    }
    double dup1000 = 0, cap = 0, dist = 0, late = 0;
    gj_vrp_eval_cta(P, s, GJ_TW_PSC, dup1000, cap, dist, late, nullptr, true);
    if (threadIdx.x == 0) {
        GjScore sc = {};
        gj_combine_vrp(P, false, dup1000, cap, dist, late, sc.v);
        gj_score_round(sc, P);
        for (int l = 0; l < 3; ++l) scores[j * 3 + l] = sc.v[l];
    }
    GJ_PHASE_MARK(13);
}

#ifdef GJ_VRP_PHASE_CLOCKS
extern "C" __attribute__((visibility("default"))) int gj_debug_vrp_phases(unsigned long long* out16, int reset) {
    if (out16) cudaMemcpyFromSymbol(out16, gj_vrp_phase_cycles, sizeof(gj_vrp_phase_cycles));
    if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(gj_vrp_phase_cycles, z, sizeof(z)); }
    return 0;
}
#endif

// build_updated_population for planned offspring (:198-213), in two launches so that the sort of the
// new scores (k_ga_rank, side stream) runs while the rows are being copied:
//   k_ga_decide   slot i takes offspring i or a random p-worst native -> scores of the new population, source
//   k_ga_copy_planned   the rows: the native's, or the parent's with the planned move applied
__global__ void __launch_bounds__(256)
k_ga_decide(GjGaArgs A, const double* __restrict__ pop_scores, const int* __restrict__ order,
            const int* __restrict__ parent_slot, const double* __restrict__ cand_scores,
            double* __restrict__ pop_scores_next, int* __restrict__ take_src, int* __restrict__ ga_src,
            double* __restrict__ trace_rep) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)A.I * A.pop) return;
    const int island = (int)(t / A.pop), i = (int)(t % A.pop);
    GjPhilox rng;
    gj_rng_init(rng, A.seed, (uint32_t)(A.island_base + island), (uint32_t)A.step,
                (uint32_t)(A.step >> 32), 0x80000000u + (uint32_t)i);
    const int rank = gj_ga_p_rank(rng, A.p_best_rate, A.pop, true, trace_rep ? trace_rep + (size_t)t * 3 : nullptr);
    const int native = order[(size_t)island * A.pop + rank];
    GjScore c, w;
    for (int l = 0; l < GJ_MAX_LEVELS; ++l) {
        c.v[l] = (l < A.levels) ? cand_scores[((size_t)island * A.n_cand + i) * A.levels + l] : 0.0;
        w.v[l] = pop_scores[((size_t)island * A.pop + native) * GJ_MAX_LEVELS + l];
    }
    const bool take = gj_score_le(c, w, A.levels);      // :207
    for (int l = 0; l < GJ_MAX_LEVELS; ++l) pop_scores_next[(size_t)t * GJ_MAX_LEVELS + l] = take ? c.v[l] : w.v[l];
    take_src[t] = take ? (parent_slot[(size_t)island * A.n_cand + i] | (int)0x80000000) : native;
    if (ga_src) ga_src[t] = take ? i : -(rank + 1);
}

__global__ void __launch_bounds__(128)
k_ga_copy_planned(GjProblemDev P, GjGroups G, GjGaArgs A, const int32_t* __restrict__ pop_rows,
                  const int* __restrict__ take_src, const GjMove* __restrict__ moves,
                  const int32_t* __restrict__ pairs, int32_t* __restrict__ pop_next) {
    __shared__ GjMove sh_move;
    __shared__ int32_t sh_pairs[kGaPairInts];
    const int island = blockIdx.x / A.pop, i = blockIdx.x % A.pop;
    const int ts = take_src[blockIdx.x];
    const bool from_cand = ts < 0;
    const int src_slot = ts & 0x7fffffff;
    if (from_cand) {
        const size_t c = (size_t)island * A.n_cand + i;
        if (threadIdx.x < (int)(sizeof(GjMove) / 4))
            reinterpret_cast<int32_t*>(&sh_move)[threadIdx.x] = reinterpret_cast<const int32_t*>(moves + c)[threadIdx.x];
        else if (threadIdx.x >= 32 && threadIdx.x < 32 + kGaPairInts)
            sh_pairs[threadIdx.x - 32] = pairs[c * kGaPairInts + threadIdx.x - 32];
    }
    const int32_t* src = pop_rows + ((size_t)island * A.pop + src_slot) * A.stride;
    int32_t* dst = pop_next + ((size_t)island * A.pop + i) * A.stride;
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (int k = threadIdx.x; k < A.stride / 4; k += blockDim.x) d4[k] = s4[k];
    if (from_cand) {
        __syncthreads();
        gj_ga_apply_planned(P, G, sh_move, sh_pairs, A.noop != 0, [&](int id) { return __ldg(src + id); },
                            [&](int id, int v) { dst[id] = v; });
    }
}

// One CTA per (island, slot).
__global__ void __launch_bounds__(128)
k_ga_replace(GjGaArgs A, const int32_t* __restrict__ pop_rows, const double* __restrict__ pop_scores,
             const int* __restrict__ order, const int32_t* __restrict__ cand_rows,
             const double* __restrict__ cand_scores, int32_t* __restrict__ pop_next,
             double* __restrict__ pop_scores_next, int* __restrict__ ga_src, double* __restrict__ trace_rep) {
    __shared__ int sh_from_cand, sh_native;
    const int island = blockIdx.x / A.pop, i = blockIdx.x % A.pop;
    if (threadIdx.x == 0) {
        GjPhilox rng;
        gj_rng_init(rng, A.seed, (uint32_t)(A.island_base + island), (uint32_t)A.step,
                    (uint32_t)(A.step >> 32), 0x80000000u + (uint32_t)i);
        const int rank = gj_ga_p_rank(rng, A.p_best_rate, A.pop, true,
                                      trace_rep ? trace_rep + ((size_t)island * A.pop + i) * 3 : nullptr);
        const int native = order[(size_t)island * A.pop + rank];
        GjScore c, w;
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) {
            c.v[l] = (l < A.levels) ? cand_scores[((size_t)island * A.n_cand + i) * A.levels + l] : 0.0;
            w.v[l] = pop_scores[((size_t)island * A.pop + native) * GJ_MAX_LEVELS + l];
        }
        const bool take = gj_score_le(c, w, A.levels);      // :207
        sh_from_cand = take ? 1 : 0;
        sh_native = native;
        for (int l = 0; l < GJ_MAX_LEVELS; ++l)
            pop_scores_next[((size_t)island * A.pop + i) * GJ_MAX_LEVELS + l] = take ? c.v[l] : w.v[l];
        if (ga_src) ga_src[(size_t)island * A.pop + i] = take ? i : -(rank + 1);
    }
    __syncthreads();
    const int32_t* src = sh_from_cand ? cand_rows + ((size_t)island * A.n_cand + i) * A.stride
                                      : pop_rows + ((size_t)island * A.pop + sh_native) * A.stride;
    int32_t* dst = pop_next + ((size_t)island * A.pop + i) * A.stride;
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (int k = threadIdx.x; k < A.stride / 4; k += blockDim.x) d4[k] = s4[k];
}

// population.sort() by counting: rank(i) = #{ j : (score_j, j) < (score_i, i) } under Ord::cmp --
// a strict total order, so the ranks are a permutation and order[rank(i)] = i is the stable sort.
// O(pop^2) comparisons, but embarrassingly parallel: the (i-tile, j-slice) grid fills every SM,
// where a single-CTA sorting network leaves 147 of them idle.
static constexpr int kRankTile = 256;

// f64 -> int64 whose signed order is f64::total_cmp's order (the transform inside gj_total_cmp)
__device__ __forceinline__ long long gj_total_order_key(double x) {
    long long l = __double_as_longlong(x);
    return l ^ (long long)(((unsigned long long)(l >> 63)) >> 1);
}

__global__ void __launch_bounds__(kRankTile)
k_ga_rank(int pop, int levels, int j_slices, const double* __restrict__ pop_scores, int* __restrict__ rank) {
    __shared__ long long sh[kRankTile * GJ_MAX_LEVELS];
    const int island = blockIdx.z;
    const double* sc = pop_scores + (size_t)island * pop * GJ_MAX_LEVELS;
    const int i = blockIdx.x * kRankTile + threadIdx.x;
    // unused levels hold 0.0 in pop_scores: they compare equal and drop out
    long long m0 = 0, m1 = 0, m2 = 0;
    if (i < pop) {
        m0 = gj_total_order_key(sc[(size_t)i * GJ_MAX_LEVELS + 0]);
        m1 = gj_total_order_key(sc[(size_t)i * GJ_MAX_LEVELS + 1]);
        m2 = gj_total_order_key(sc[(size_t)i * GJ_MAX_LEVELS + 2]);
    }
    const int per = (pop + j_slices - 1) / j_slices;
    const int j0 = blockIdx.y * per, j1 = min(pop, j0 + per);
    int below = 0;
    for (int base = j0; base < j1; base += kRankTile) {
        const int m = min(kRankTile, j1 - base);
        __syncthreads();
        for (int k = threadIdx.x; k < m * GJ_MAX_LEVELS; k += kRankTile)
            sh[k] = gj_total_order_key(sc[(size_t)base * GJ_MAX_LEVELS + k]);
        __syncthreads();
        if (i < pop) {
#pragma unroll 4
            for (int k = 0; k < m; ++k) {
                const long long o0 = sh[k * GJ_MAX_LEVELS], o1 = sh[k * GJ_MAX_LEVELS + 1], o2 = sh[k * GJ_MAX_LEVELS + 2];
                const bool lt = (o0 < m0) || (o0 == m0 && ((o1 < m1) || (o1 == m1 && ((o2 < m2) || (o2 == m2 && base + k < i)))));
                below += lt ? 1 : 0;
            }
        }
    }
    if (i < pop && below) atomicAdd(&rank[(size_t)island * pop + i], below);
}

__global__ void k_ga_order_from_rank(int pop, int I, int* __restrict__ rank, int* __restrict__ order) {
    const int64_t n = (int64_t)pop * I;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int island = (int)(t / pop), i = (int)(t % pop);
        order[(size_t)island * pop + rank[t]] = i;
        rank[t] = 0;                                   // ready for the next generation
    }
}

// update_top_individual: population[0] <= agent_top -> replace (agent_base.rs:220-224)
__global__ void k_ga_top(int pop, int stride, int n_vars, int levels, int n_cand,
                         const int32_t* __restrict__ pop_rows, const double* __restrict__ pop_scores,
                         const int* __restrict__ order, int32_t* best, double* best_score,
                         unsigned long long* counters) {
    __shared__ int sh_take;
    const int island = blockIdx.x;
    const int r0 = order[(size_t)island * pop];
    if (threadIdx.x == 0) {
        GjScore c = {}, t = {};
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) {
            c.v[l] = pop_scores[((size_t)island * pop + r0) * GJ_MAX_LEVELS + l];
            t.v[l] = best_score[(size_t)island * GJ_MAX_LEVELS + l];
        }
        const bool take = gj_score_le(c, t, levels);
        if (take) for (int l = 0; l < GJ_MAX_LEVELS; ++l) best_score[(size_t)island * GJ_MAX_LEVELS + l] = c.v[l];
        sh_take = take ? 1 : 0;
        if (counters) {
            atomicAdd(&counters[0], (unsigned long long)n_cand);
            if (island == 0) atomicAdd(&counters[1], 1ull);
        }
    }
    __syncthreads();
    if (sh_take) {
        const int32_t* src = pop_rows + ((size_t)island * pop + r0) * stride;
        for (int i = threadIdx.x; i < n_vars; i += blockDim.x) best[(size_t)island * stride + i] = src[i];
    }
}

// population.sort() + update_top_individual (+ update_global_top's publish half when the group is one
// island) in one launch: order[rank[t]] = t; the CTA that meets rank 0 of an island compares that
// individual with the agent's top (agent_base.rs:220-224) and, if it wins, copies the row.
__global__ void __launch_bounds__(256)
k_ga_finish(int pop, int I, int stride, int n_vars, int levels, int n_cand, int* __restrict__ rank,
            int* __restrict__ order, const int32_t* __restrict__ pop_rows, const double* __restrict__ pop_scores,
            int32_t* best, double* best_score, int32_t* gbest, double* gbest_score, unsigned long long* counters) {
    __shared__ int sh_first, sh_take, sh_gtake;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool mine = false;                                 // this thread holds an island's rank-0 individual
    if (t < (int64_t)pop * I) {
        const int r = rank[t];
        order[(t / pop) * pop + r] = (int)(t % pop);
        rank[t] = 0;                                   // ready for the next generation
        mine = r == 0;
    }
    // a CTA meets one rank-0 individual per island it spans (usually none or one)
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sh_first = -1;
        __syncthreads();
        if (mine) atomicMax(&sh_first, (int)t);
        __syncthreads();
        const int first = sh_first;
        if (first < 0) break;
        if ((int)t == first) mine = false;
        const int island = first / pop, r0 = first % pop;
        if (threadIdx.x == 0) {
            GjScore c = {}, tp = {}, g = {};
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) {
                c.v[l] = pop_scores[(size_t)first * GJ_MAX_LEVELS + l];
                tp.v[l] = best_score[(size_t)island * GJ_MAX_LEVELS + l];
                g.v[l] = gbest_score[l];
            }
            const bool take = gj_score_le(c, tp, levels);
            if (take) for (int l = 0; l < GJ_MAX_LEVELS; ++l) best_score[(size_t)island * GJ_MAX_LEVELS + l] = c.v[l];
            sh_take = take ? 1 : 0;
            // a group of one island: its top is the only candidate for the group's global top
            // (update_global_top's publish half, agent_base.rs:451-461)
            const GjScore nt = take ? c : tp;
            const bool gt = (I == 1) && !gj_score_le(g, nt, levels);
            if (gt) for (int l = 0; l < GJ_MAX_LEVELS; ++l) gbest_score[l] = nt.v[l];
            sh_gtake = gt ? 1 : 0;
            if (counters) {
                atomicAdd(&counters[0], (unsigned long long)n_cand);
                if (island == 0) atomicAdd(&counters[1], 1ull);
            }
        }
        __syncthreads();
        const int32_t* src = pop_rows + ((size_t)island * pop + r0) * stride;
        if (sh_take)
            for (int i = threadIdx.x; i < n_vars; i += blockDim.x) best[(size_t)island * stride + i] = src[i];
        if (sh_gtake) {
            const int32_t* gsrc = sh_take ? src : best + (size_t)island * stride;
            for (int i = threadIdx.x; i < n_vars; i += blockDim.x) gbest[i] = gsrc[i];
        }
    }
}

// migrants = the island's first ceil(migration_rate * pop) individuals by rank
// mailbox slot s (s = 0..I): [migrants][stride int32 + 3 f64]
__global__ void k_ga_migrate_pack(int pop, int stride, int migrants, const int32_t* __restrict__ pop_rows,
                                  const double* __restrict__ pop_scores, const int* __restrict__ order,
                                  unsigned char* mailbox) {
    const int island = blockIdx.x / migrants, mgr = blockIdx.x % migrants;
    const size_t ind_bytes = (size_t)stride * 4 + GJ_MAX_LEVELS * 8;
    unsigned char* slot = mailbox + ((size_t)(island + 1) * migrants + mgr) * ind_bytes;
    const int r = order[(size_t)island * pop + mgr];
    const int32_t* src = pop_rows + ((size_t)island * pop + r) * stride;
    int32_t* row = (int32_t*)slot;
    double* sc = (double*)(slot + (size_t)stride * 4);
    for (int i = threadIdx.x; i < stride; i += blockDim.x) row[i] = src[i];
    if (threadIdx.x < GJ_MAX_LEVELS) sc[threadIdx.x] = pop_scores[((size_t)island * pop + r) * GJ_MAX_LEVELS + threadIdx.x];
}

__global__ void k_copy_bytes(const unsigned char* src, unsigned char* dst, size_t bytes) {
    const uint32_t* s = (const uint32_t*)src; uint32_t* d = (uint32_t*)dst;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < bytes / 4; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}

// receive_updates for Population agents: migrant i vs population[pop - count + i] (by rank);
// replace when migrant <= native (agent_base.rs:405-412, 435-439)
__global__ void k_ga_migrate_recv(int pop, int stride, int migrants, int levels,
                                  const unsigned char* __restrict__ mailbox, int32_t* pop_rows,
                                  double* pop_scores, const int* __restrict__ order) {
    __shared__ int sh_take;
    const int island = blockIdx.x / migrants, mgr = blockIdx.x % migrants;
    const size_t ind_bytes = (size_t)stride * 4 + GJ_MAX_LEVELS * 8;
    const unsigned char* slot = mailbox + ((size_t)island * migrants + mgr) * ind_bytes;
    const int32_t* row = (const int32_t*)slot;
    const double* sc = (const double*)(slot + (size_t)stride * 4);
    const int r = order[(size_t)island * pop + (pop - migrants + mgr)];
    if (threadIdx.x == 0) {
        GjScore m = {}, n = {};
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) { m.v[l] = sc[l]; n.v[l] = pop_scores[((size_t)island * pop + r) * GJ_MAX_LEVELS + l]; }
        const bool take = gj_score_le(m, n, levels);
        if (take) for (int l = 0; l < GJ_MAX_LEVELS; ++l) pop_scores[((size_t)island * pop + r) * GJ_MAX_LEVELS + l] = m.v[l];
        sh_take = take ? 1 : 0;
    }
    __syncthreads();
    if (sh_take) {
        int32_t* dst = pop_rows + ((size_t)island * pop + r) * stride;
        for (int i = threadIdx.x; i < stride; i += blockDim.x) dst[i] = row[i];
    }
}

__global__ void k_ga_init_scores(int I, int pop, int levels, const double* __restrict__ scored,
                                 double* pop_scores, double* best_score, double* gbest_score) {
    const int64_t n = (int64_t)I * pop;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        for (int l = 0; l < GJ_MAX_LEVELS; ++l)
            pop_scores[j * GJ_MAX_LEVELS + l] = (l < levels) ? scored[j * levels + l] : 0.0;
    if (blockIdx.x == 0 && threadIdx.x < GJ_MAX_LEVELS) {
        gbest_score[threadIdx.x] = (threadIdx.x < levels) ? 1.7976931348623157e308 : 0.0;
        for (int i = 0; i < I; ++i)
            best_score[(size_t)i * GJ_MAX_LEVELS + threadIdx.x] = (threadIdx.x < levels) ? 1.7976931348623157e308 : 0.0;
    }
}

// ---- host -------------------------------------------------------------------------------------------

template <class T>
static gj_status ga_alloc(gj_islands* g, size_t n, T** out) {
    void* d = nullptr;
    size_t bytes = (n ? n : 1) * sizeof(T);
    GJ_CUDA_TRY(cudaMalloc(&d, bytes));
    g->allocs.push_back(d);
    GJ_CUDA_TRY(cudaMemset(d, 0, bytes));
    *out = (T*)d;
    return GJ_OK;
}

static gj_status ga_sort_and_top(gj_islands* g, cudaStream_t st, bool count) {
    const int i_tiles = (g->pop + kRankTile - 1) / kRankTile;
    int j_slices = std::max(1, std::min((g->pop + kRankTile - 1) / kRankTile, (148 * 8) / std::max(1, i_tiles * g->I)));
    k_ga_rank<<<dim3(i_tiles, j_slices, g->I), kRankTile, 0, st>>>(g->pop, g->levels, j_slices, g->pop_scores, g->ga_rank);
    GJ_LAUNCH_CHECK();
    k_ga_order_from_rank<<<(unsigned)std::min<int64_t>(((int64_t)g->pop * g->I + 255) / 256, 1184), 256, 0, st>>>(
        g->pop, g->I, g->ga_rank, g->order);
    GJ_LAUNCH_CHECK();
    k_ga_top<<<g->I, 128, 0, st>>>(g->pop, g->stride, g->n_vars, g->levels, g->n_cand, g->pop_rows,
                                  g->pop_scores, g->order, g->best, g->best_score, count ? g->counters : nullptr);
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

gj_status gj_ga_create(gj_problem* p, const gj_agent_params* prm, const double* initial, gj_islands** out) {
    std::unique_ptr<gj_islands> g(new gj_islands());
    gj_status rc;
    if ((rc = gj_islands_common_init(g.get(), p, prm))) return rc;
    if (prm->population_size < 2) return gj_fail(GJ_ERR_INVALID, "population_size must be >= 2");
    if (prm->population_size > 65536) return gj_fail(GJ_ERR_UNSUPPORTED, "population_size > 65536");
    if (!(prm->p_best_rate > 0.000001)) return gj_fail(GJ_ERR_INVALID, "p_best_rate must exceed 1e-6");
    g->pop = (int)prm->population_size;
    g->half = (int)std::ceil(0.5 * (double)g->pop);          // genetic_algorithm_base.rs:51
    g->n_cand = 2 * g->half;
    g->K = g->n_cand;
    g->migrants = std::max<int64_t>(1, (int64_t)std::ceil(prm->migration_rate * (double)g->pop));   // agent_base.rs:339
    g->migrants = std::min<int64_t>(g->migrants, g->pop);
    const int I = g->I, stride = g->stride, pop = g->pop;
    if ((rc = ga_alloc(g.get(), (size_t)I * pop * stride, &g->pop_rows))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * pop * stride, &g->pop_next))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * pop * GJ_MAX_LEVELS, &g->pop_scores))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * pop * GJ_MAX_LEVELS, &g->pop_scores_next))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * g->n_cand * stride, &g->cand_rows))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * std::max(g->n_cand, pop) * GJ_MAX_LEVELS, &g->cand_scores))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * g->n_cand, &g->moves))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * pop, &g->order))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * pop, &g->ga_rank))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * pop, &g->ga_src))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * g->n_cand, &g->ga_parent))) return rc;
    if (p->dev.kind >= GJ_VRP && (rc = ga_alloc(g.get(), (size_t)I * g->n_cand, &g->ga_moves_next))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * pop, &g->ga_take))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * g->n_cand * kGaPairInts, &g->ga_pairs))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * stride, &g->best))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)I * GJ_MAX_LEVELS, &g->best_score))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)stride, &g->gbest))) return rc;
    if ((rc = ga_alloc(g.get(), (size_t)GJ_MAX_LEVELS, &g->gbest_score))) return rc;
    const size_t ind_bytes = (size_t)stride * 4 + GJ_MAX_LEVELS * 8;
    if ((rc = ga_alloc(g.get(), (size_t)(I + 1) * g->migrants * ind_bytes, &g->mailbox))) return rc;

    // Agent::init_population (agent_base.rs:193-203): population_size samples scored in one
    // plain batch.  `initial` (if given) seeds individual 0 of every island; the rest are
    // sample_variables() draws (initial value where defined, else uniform).
    {
        std::vector<int32_t> host((size_t)I * pop * stride, 0), row(p->dev.n_vars);
        uint64_t rng = prm->seed ^ 0x5DEECE66Dull;
        for (int i = 0; i < I; ++i)
            for (int k = 0; k < pop; ++k) {
                const double* given = (initial && k == 0) ? initial + (size_t)i * p->dev.n_vars : nullptr;
                gj_islands_start_vector(p, given, rng, row);
                std::copy(row.begin(), row.end(), host.begin() + ((size_t)i * pop + k) * stride);
            }
        GJ_CUDA_TRY(cudaMemcpy(g->pop_rows, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
    }
    cudaStream_t st = p->stream;
    if ((rc = gj_launch_score_plain_i32(p, g->pop_rows, stride, (int64_t)I * pop, g->cand_scores, false, st))) return rc;
    k_ga_init_scores<<<148, 256, 0, st>>>(I, pop, g->levels, g->cand_scores, g->pop_scores, g->best_score, g->gbest_score);
    GJ_LAUNCH_CHECK();
    if ((rc = ga_sort_and_top(g.get(), st, false))) return rc;
    GJ_CUDA_TRY(cudaStreamSynchronize(st));
    g->steps_to_send = std::max<int64_t>(1, prm->migration_frequency);
    *out = g.release();
    return GJ_OK;
}

static GjGaArgs ga_args(gj_islands* g) {
    GjGaArgs A{};
    A.I = g->I; A.pop = g->pop; A.half = g->half; A.n_cand = g->n_cand; A.stride = g->stride;
    A.n_vars = g->n_vars; A.levels = g->levels; A.noop = g->noop;
    A.crossover_probability = g->prm.crossover_probability; A.p_best_rate = g->prm.p_best_rate;
    A.seed = g->prm.seed; A.step = g->step; A.island_base = g->island_base;
    return A;
}

static gj_status ga_migrate_pack(gj_islands* g, cudaStream_t st) {
    k_ga_migrate_pack<<<g->I * (int)g->migrants, 128, 0, st>>>(g->pop, g->stride, (int)g->migrants, g->pop_rows,
                                                             g->pop_scores, g->order, g->mailbox);
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

static gj_status ga_migrate_recv(gj_islands* g, cudaStream_t st) {
    k_ga_migrate_recv<<<g->I * (int)g->migrants, 128, 0, st>>>(g->pop, g->stride, (int)g->migrants, g->levels,
                                                             g->mailbox, g->pop_rows, g->pop_scores, g->order);
    GJ_LAUNCH_CHECK();
    // the next sample_candidates_plain starts with population.sort() (:157)
    return ga_sort_and_top(g, st, false);
}

// side stream + events of the overlapped generation, created on first use (GJ_GA_OVERLAP=0 switches it off)
static bool ga_side_ready(gj_islands* g) {
    if (g->ga_side_state == 0) {
        const char* e = getenv("GJ_GA_OVERLAP");
        g->ga_side_state = -1;
        if (!(e && e[0] == '0')) {
            bool ok = cudaStreamCreateWithFlags(&g->ga_side, cudaStreamNonBlocking) == cudaSuccess;
            for (int i = 0; ok && i < 4; ++i) ok = cudaEventCreateWithFlags(&g->ga_ev[i], cudaEventDisableTiming) == cudaSuccess;
            if (ok) g->ga_side_state = 1;
        }
    }
    return g->ga_side_state == 1;
}

gj_status gj_ga_step(gj_islands* g, int64_t n_steps, cudaStream_t st) {
    const GjProblemDev& P = g->p->dev;
    gj_status rc;
    // VRP models (one CTA scores one candidate): offspring are planned, scored from their parent's row and
    // written only when they survive; the other models copy, then score whole rows
    const bool planned = P.kind >= GJ_VRP;
    for (int64_t s = 0; s < n_steps; ++s) {
        GjGaArgs A = ga_args(g);
        const int64_t S = (int64_t)g->I * g->n_cand;
        const bool trace = g->ga_trace_sel != nullptr;
        if (planned) {
            // moves generated ahead on the side stream (see k_ga_gen_moves)?
            const bool ahead = g->ga_moves_next && g->ga_moves_step == (int64_t)g->step && g->ga_side_state == 1;
            if (ahead) {
                std::swap(g->moves, g->ga_moves_next);
                GJ_CUDA_TRY(cudaStreamWaitEvent(st, g->ga_ev[3], 0));
            }
            k_ga_plan<<<(unsigned)((S + 127) / 128), 128, 0, st>>>(P, g->groups, g->mover, A, g->order, g->ga_parent,
                                                                  ahead ? nullptr : g->moves, g->tabu_bits, g->tabu_words,
                                                                  g->tabu_word_off, g->ga_trace_sel);
            GJ_LAUNCH_CHECK();
        } else {
            k_ga_offspring<<<g->I * g->n_cand, 128, 0, st>>>(P, g->groups, g->mover, A, g->pop_rows, g->order, g->cand_rows, g->moves,
                                                             g->tabu_bits, g->tabu_words, g->tabu_word_off, g->ga_trace_sel);
            GJ_LAUNCH_CHECK();
        }
        // the mover's tabu deque and, later, the sort of the new population run on a side stream next to
        // the scorer and the row copies (both only have to be done before the next generation is planned)
        const bool overlap = planned && ga_side_ready(g);
        if (overlap) {
            GJ_CUDA_TRY(cudaEventRecord(g->ga_ev[0], st));
            GJ_CUDA_TRY(cudaStreamWaitEvent(g->ga_side, g->ga_ev[0], 0));
        }
        if (g->tabu_bits) {
            k_ga_tabu_update<<<g->I, 256, 0, overlap ? g->ga_side : st>>>(
                g->groups, g->n_cand, g->groups.n_groups, g->moves, g->tabu_bits, g->tabu_words, g->tabu_word_off,
                g->tabu_ring[g->step & 1], g->tabu_ring[(g->step + 1) & 1], g->tabu_ring_len, g->tabu_ring_off, g->tabu_size,
                g->tabu_fill);
            GJ_LAUNCH_CHECK();
        }
        if (overlap && g->ga_moves_next) {
            // the next generation's moves, behind this generation's deque update on the side stream
            k_ga_gen_moves<<<(unsigned)((S + 127) / 128), 128, 0, g->ga_side>>>(
                P, g->groups, g->mover, g->prm.seed, g->step + 1, g->island_base, g->I, g->n_cand, g->ga_moves_next, g->tabu_bits,
                g->tabu_words, g->tabu_word_off);
            GJ_LAUNCH_CHECK();
            GJ_CUDA_TRY(cudaEventRecord(g->ga_ev[3], g->ga_side));
            g->ga_moves_step = (int64_t)g->step + 1;
        }
        if ((rc = gj_prof_begin(g, st))) return rc;
        if (planned) {
            const size_t smem = gj_vrp_smem_bytes(P.n_entities, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
            if (smem > 48 * 1024)
                GJ_CUDA_TRY(cudaFuncSetAttribute(k_ga_score_planned_vrp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            // (no shared-memory carve-out preference: the L1 the default leaves serves the customer-fact
            // gathers; asking for the maximum carve-out cost 100 us per generation)
            k_ga_score_planned_vrp<<<(unsigned)S, kVrpWarps * 32, smem, st>>>(P, g->groups, A, g->pop_rows, g->ga_parent, g->moves,
                                                                             g->ga_pairs, g->cand_scores, trace ? g->cand_rows : nullptr);
            GJ_LAUNCH_CHECK();
        } else {
            if ((rc = gj_launch_score_plain_i32(g->p, g->cand_rows, g->stride, S, g->cand_scores, false, st))) return rc;
        }
        if ((rc = gj_prof_end(g, st))) return rc;
        if (planned) {
            const int64_t NP = (int64_t)g->I * g->pop;
            k_ga_decide<<<(unsigned)((NP + 255) / 256), 256, 0, st>>>(A, g->pop_scores, g->order, g->ga_parent, g->cand_scores,
                                                                     g->pop_scores_next, g->ga_take, g->ga_src, g->ga_trace_rep);
            GJ_LAUNCH_CHECK();
            cudaStream_t rs = st;
            if (overlap) {
                GJ_CUDA_TRY(cudaEventRecord(g->ga_ev[1], st));
                GJ_CUDA_TRY(cudaStreamWaitEvent(g->ga_side, g->ga_ev[1], 0));
                rs = g->ga_side;
            }
            const int i_tiles = (g->pop + kRankTile - 1) / kRankTile;
            const int j_slices = std::max(1, std::min((g->pop + kRankTile - 1) / kRankTile, (148 * 8) / std::max(1, i_tiles * g->I)));
            k_ga_rank<<<dim3(i_tiles, j_slices, g->I), kRankTile, 0, rs>>>(g->pop, g->levels, j_slices, g->pop_scores_next, g->ga_rank);
            GJ_LAUNCH_CHECK();
            if (overlap) GJ_CUDA_TRY(cudaEventRecord(g->ga_ev[2], g->ga_side));
            k_ga_copy_planned<<<g->I * g->pop, 128, 0, st>>>(P, g->groups, A, g->pop_rows, g->ga_take, g->moves, g->ga_pairs, g->pop_next);
            GJ_LAUNCH_CHECK();
            if (overlap) GJ_CUDA_TRY(cudaStreamWaitEvent(st, g->ga_ev[2], 0));
            std::swap(g->pop_rows, g->pop_next);
            std::swap(g->pop_scores, g->pop_scores_next);
            k_ga_finish<<<(unsigned)((NP + 255) / 256), 256, 0, st>>>(g->pop, g->I, g->stride, g->n_vars, g->levels, g->n_cand, g->ga_rank,
                                                                     g->order, g->pop_rows, g->pop_scores, g->best, g->best_score,
                                                                     g->gbest, g->gbest_score, g->counters);
            GJ_LAUNCH_CHECK();
        } else {
            k_round_scores<<<(unsigned)std::min<int64_t>((S + 255) / 256, 1184), 256, 0, st>>>(P, g->cand_scores, S);   // agent_base.rs:284-287
            GJ_LAUNCH_CHECK();
            k_ga_replace<<<g->I * g->pop, 128, 0, st>>>(A, g->pop_rows, g->pop_scores, g->order, g->cand_rows, g->cand_scores,
                                                       g->pop_next, g->pop_scores_next, g->ga_src, g->ga_trace_rep);
            GJ_LAUNCH_CHECK();
            std::swap(g->pop_rows, g->pop_next);
            std::swap(g->pop_scores, g->pop_scores_next);
            if ((rc = ga_sort_and_top(g, st, true))) return rc;
        }
        g->step += 1;
        g->steps_to_send -= 1;
        if (g->steps_to_send <= 0) {
            if (!g->external_ring) {
                if ((rc = ga_migrate_pack(g, st))) return rc;
                const size_t slot = (size_t)g->migrants * ((size_t)g->stride * 4 + GJ_MAX_LEVELS * 8);
                k_copy_bytes<<<64, 256, 0, st>>>(g->mailbox + (size_t)g->I * slot, g->mailbox, slot);
                GJ_LAUNCH_CHECK();
                if ((rc = ga_migrate_recv(g, st))) return rc;
            }
            g->steps_to_send = std::max<int64_t>(1, g->prm.migration_frequency);
        }
        if (!(planned && g->I == 1) && (rc = gj_ga_global_top(g, st))) return rc;
    }
    return GJ_OK;
}

// update_global_top for GA: only the "publish" half (agent_base.rs:451-461); Population agents
// never adopt the global best (:487 `_ => ()`).
__global__ void k_ga_global_reduce(int I, int levels, int stride, int n_vars, const int32_t* __restrict__ best,
                                   const double* __restrict__ best_score, int32_t* gbest, double* gbest_score);

gj_status gj_ga_global_top(gj_islands* g, cudaStream_t st) {
    k_ga_global_reduce<<<1, 256, 0, st>>>(g->I, g->levels, g->stride, g->n_vars, g->best, g->best_score, g->gbest, g->gbest_score);
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

__global__ void k_ga_global_reduce(int I, int levels, int stride, int n_vars, const int32_t* __restrict__ best,
                                   const double* __restrict__ best_score, int32_t* gbest, double* gbest_score) {
    __shared__ int sh_win;
    if (threadIdx.x == 0) {
        GjScore g = {};
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) g.v[l] = gbest_score[l];
        int win = -1;
        for (int i = 0; i < I; ++i) {
            GjScore s = {};
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) s.v[l] = best_score[(size_t)i * GJ_MAX_LEVELS + l];
            if (!gj_score_le(g, s, levels)) { g = s; win = i; }
        }
        if (win >= 0) for (int l = 0; l < GJ_MAX_LEVELS; ++l) gbest_score[l] = g.v[l];
        sh_win = win;
    }
    __syncthreads();
    if (sh_win >= 0)
        for (int i = threadIdx.x; i < n_vars; i += blockDim.x) gbest[i] = best[(size_t)sh_win * stride + i];
}

gj_status gj_ga_current(gj_islands* g, int32_t island, double* vars, double* score) {
    int r0 = 0;
    GJ_CUDA_TRY(cudaMemcpy(&r0, g->order + (size_t)island * g->pop, 4, cudaMemcpyDeviceToHost));
    std::vector<int32_t> row(g->n_vars);
    double sc[GJ_MAX_LEVELS];
    GJ_CUDA_TRY(cudaMemcpy(row.data(), g->pop_rows + ((size_t)island * g->pop + r0) * g->stride, (size_t)g->n_vars * 4, cudaMemcpyDeviceToHost));
    GJ_CUDA_TRY(cudaMemcpy(sc, g->pop_scores + ((size_t)island * g->pop + r0) * GJ_MAX_LEVELS, sizeof(sc), cudaMemcpyDeviceToHost));
    if (vars) for (int i = 0; i < g->n_vars; ++i) vars[i] = (double)row[i];
    if (score) for (int l = 0; l < g->levels; ++l) score[l] = sc[l];
    return GJ_OK;
}

// packs the migrants of every island; *d_slot = the LAST island's outgoing slot (what leaves the group)
gj_status gj_ga_pack_outgoing(gj_islands* g, cudaStream_t st, const unsigned char** d_slot) {
    gj_status rc;
    if ((rc = ga_migrate_pack(g, st))) return rc;
    const size_t slot = (size_t)g->migrants * ((size_t)g->stride * 4 + GJ_MAX_LEVELS * 8);
    *d_slot = g->mailbox + (size_t)g->I * slot;
    return GJ_OK;
}

gj_status gj_ga_export(gj_islands* g, void* d_buffer, cudaStream_t st) {
    gj_status rc;
    if ((rc = ga_migrate_pack(g, st))) return rc;
    const size_t slot = (size_t)g->migrants * ((size_t)g->stride * 4 + GJ_MAX_LEVELS * 8);
    GJ_CUDA_TRY(cudaMemcpyAsync(d_buffer, g->mailbox + (size_t)g->I * slot, slot, cudaMemcpyDeviceToDevice, st));
    return GJ_OK;
}

gj_status gj_ga_import(gj_islands* g, const void* d_buffer, cudaStream_t st) {
    const size_t slot = (size_t)g->migrants * ((size_t)g->stride * 4 + GJ_MAX_LEVELS * 8);
    GJ_CUDA_TRY(cudaMemcpyAsync(g->mailbox, d_buffer, slot, cudaMemcpyDeviceToDevice, st));
    return ga_migrate_recv(g, st);
}

// ---- test / inspection hooks ---------------------------------------------------------------------------
extern "C" gj_status gj_islands_ga_population(gj_islands* g, int32_t island, double* rows, double* scores,
                                              int32_t* order) {
    if (!g || g->prm.agent != GJ_AGENT_GENETIC_ALGORITHM || island < 0 || island >= g->I)
        return gj_fail(GJ_ERR_INVALID, "bad GA island");
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    GJ_CUDA_TRY(cudaDeviceSynchronize());
    const int pop = g->pop, n = g->n_vars;
    if (rows) {
        std::vector<int32_t> h((size_t)pop * g->stride);
        GJ_CUDA_TRY(cudaMemcpy(h.data(), g->pop_rows + (size_t)island * pop * g->stride, h.size() * 4, cudaMemcpyDeviceToHost));
        for (int k = 0; k < pop; ++k)
            for (int i = 0; i < n; ++i) rows[(size_t)k * n + i] = (double)h[(size_t)k * g->stride + i];
    }
    if (scores) {
        std::vector<double> h((size_t)pop * GJ_MAX_LEVELS);
        GJ_CUDA_TRY(cudaMemcpy(h.data(), g->pop_scores + (size_t)island * pop * GJ_MAX_LEVELS, h.size() * 8, cudaMemcpyDeviceToHost));
        for (int k = 0; k < pop; ++k)
            for (int l = 0; l < g->levels; ++l) scores[(size_t)k * g->levels + l] = h[(size_t)k * GJ_MAX_LEVELS + l];
    }
    if (order) GJ_CUDA_TRY(cudaMemcpy(order, g->order + (size_t)island * pop, (size_t)pop * 4, cudaMemcpyDeviceToHost));
    return GJ_OK;
}

extern "C" gj_status gj_islands_ga_trace_generation(gj_islands* g, int32_t island, gj_ga_trace* out) {
    if (!g || !out || g->prm.agent != GJ_AGENT_GENETIC_ALGORITHM || island < 0 || island >= g->I)
        return gj_fail(GJ_ERR_INVALID, "bad GA island");
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    gj_status rc;
    if (!g->ga_trace_sel) {
        if ((rc = ga_alloc(g, (size_t)g->I * g->half * 8, &g->ga_trace_sel))) return rc;
        if ((rc = ga_alloc(g, (size_t)g->I * g->pop * 3, &g->ga_trace_rep))) return rc;
    }
    GJ_CUDA_TRY(cudaDeviceSynchronize());
    const int pop = g->pop, n = g->n_vars, nc = g->n_cand;
    if (out->order_before)
        GJ_CUDA_TRY(cudaMemcpy(out->order_before, g->order + (size_t)island * pop, (size_t)pop * 4, cudaMemcpyDeviceToHost));
    cudaStream_t st = g->p->stream;
    if ((rc = gj_ga_step(g, 1, st))) return rc;
    GJ_CUDA_TRY(cudaStreamSynchronize(st));
    if (out->pairs)
        GJ_CUDA_TRY(cudaMemcpy(out->pairs, g->ga_trace_sel + (size_t)island * g->half * 8, (size_t)g->half * 8 * 8, cudaMemcpyDeviceToHost));
    if (out->replace)
        GJ_CUDA_TRY(cudaMemcpy(out->replace, g->ga_trace_rep + (size_t)island * pop * 3, (size_t)pop * 3 * 8, cudaMemcpyDeviceToHost));
    if (out->src)
        GJ_CUDA_TRY(cudaMemcpy(out->src, g->ga_src + (size_t)island * pop, (size_t)pop * 4, cudaMemcpyDeviceToHost));
    if (out->move_desc) {
        std::vector<GjMove> mv(nc);
        GJ_CUDA_TRY(cudaMemcpy(mv.data(), g->moves + (size_t)island * nc, (size_t)nc * sizeof(GjMove), cudaMemcpyDeviceToHost));
        for (int j = 0; j < nc; ++j) {
            int32_t* d = out->move_desc + (size_t)j * 20;
            d[0] = mv[j].kind; d[1] = mv[j].group; d[2] = mv[j].k; d[3] = 0;
            for (int i = 0; i < GJ_MOVE_MAXK; ++i) { d[4 + i] = mv[j].a[i]; d[12 + i] = mv[j].v[i]; }
        }
    }
    if (out->cand_rows) {
        std::vector<int32_t> h((size_t)nc * g->stride);
        GJ_CUDA_TRY(cudaMemcpy(h.data(), g->cand_rows + (size_t)island * nc * g->stride, h.size() * 4, cudaMemcpyDeviceToHost));
        for (int k = 0; k < nc; ++k)
            for (int i = 0; i < n; ++i) out->cand_rows[(size_t)k * n + i] = (double)h[(size_t)k * g->stride + i];
    }
    if (out->cand_scores)
        GJ_CUDA_TRY(cudaMemcpy(out->cand_scores, g->cand_scores + (size_t)island * nc * g->levels,
                               (size_t)nc * g->levels * 8, cudaMemcpyDeviceToHost));
    // switch the trace writes off again (the buffers stay allocated)
    return GJ_OK;
}
