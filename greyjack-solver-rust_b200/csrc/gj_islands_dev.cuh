// gj_islands_dev.cuh -- device-side pieces shared by the island kernels of every translation unit
// (gj_islands.cu: per-step kernels, migration, global top; gj_islands_fused.cu: the fused step;
// gj_islands_chain.cu / gj_islands_vrp_chain.cu: many-steps-per-launch chains): the argument
// blocks, acceptance helpers, update_top_individual, the tabu table and update_global_top's adopt half.
#pragma once

#include "gj_eval.cuh"
#include "gj_islands.hpp"

// x mod m for 0 <= x < 2m (ring indices), without the integer division
__device__ __forceinline__ int gj_wrap_once(int x, int m) { return x >= m ? x - m : x; }

__device__ __forceinline__ int gj_vrp_tw_mode(const GjProblemDev& P) {
    return P.kind == GJ_VRP_SERVICE ? GJ_TW_ISC_SERVICE : GJ_TW_ISC_FILE;      // islands score with the ISC
}

struct GjSelectArgs {
    int agent;                  // GJ_AGENT_TABU_SEARCH / GJ_AGENT_LATE_ACCEPTANCE
    int K, stride, levels, n_vars;
    int late_size;
    int noop;
    int n_groups;
    const GjMove* moves;        // stored moves, or nullptr: regenerate from the counter RNG
    GjMoverParams M; uint64_t seed; uint64_t step; int island_base;
    const double* cand_scores;
    int32_t* cur; double* cur_score;
    int32_t* best; double* best_score;
    int* dirty;
    double* late; int* late_head; int* late_len;      // LA: circular deque per island
    unsigned long long* counters;                     // [0] candidates [1] steps [2] accepted
    // tabu state: rank-indexed deques (slot 0 = newest), read from _old, written to _new
    uint32_t* tabu_bits; int tabu_words_per_island; const int32_t* tabu_word_off;
    const int32_t* tabu_ring_old; int32_t* tabu_ring_new; int tabu_ring_per_island; const int32_t* tabu_ring_off;
    const int32_t* tabu_size; int* tabu_fill;
    // delta scoring: the island's cached state goes stale when cur changes; update_top_individual
    // is deferred to k_refresh (after the exact re-score of the accepted neighbour)
    int* stale; int defer_top; int* work_count;
    // SimulatedAnnealing: temperatures per island [I][GJ_MAX_LEVELS], schedule
    double* sa_temp; GjSaParams sa;
    // fused islands adopt the published global top at the START of their next step (P0)
    int compare_to_global;
    int chain_mode;             // 1: the group runs chain kernels (gseen marks "solution is gbest of version v")
    int32_t* gbest; double* gbest_score; int* gver; int* gseen;
    // trace
    long long* selected_out; int* accepted_out; double* aux_out;
};

// uniform [0, 1) of the acceptance rule of (island, step): its own RNG stream
__device__ __forceinline__ double gj_accept_uniform(uint64_t seed, uint32_t island_global, uint64_t step) {
    GjPhilox rng;
    gj_rng_init(rng, seed, island_global, (uint32_t)step, (uint32_t)(step >> 32), 0xFFFFFFF0u);
    return gj_rng_f64(rng);
}

// SimulatedAnnealing acceptance of one island's single neighbour (thread 0 of its CTA / lane 0)
__device__ __forceinline__ bool gj_sa_step_accept(const GjSelectArgs& A, int island, const GjScore& b,
                                                  const GjScore& cur) {
    double* temp = A.sa_temp + (size_t)island * GJ_MAX_LEVELS;
    double t[GJ_MAX_LEVELS] = {temp[0], temp[1], temp[2]};
    const double u = gj_accept_uniform(A.seed, (uint32_t)(A.island_base + island), A.step);
    double proba;
    const bool accept = gj_sa_accept(b, cur, A.levels, t, A.sa, u, &proba);
    for (int l = 0; l < GJ_MAX_LEVELS; ++l) temp[l] = t[l];
    if (A.aux_out) {
        double* o = A.aux_out + (size_t)island * 5;
        o[0] = u; o[1] = proba; o[2] = t[0]; o[3] = t[1]; o[4] = t[2];
    }
    return accept;
}

__device__ __forceinline__ GjScore gj_load_score(const double* p, int levels) {
    GjScore s;
    for (int l = 0; l < GJ_MAX_LEVELS; ++l) s.v[l] = (l < levels) ? p[l] : 0.0;
    return s;
}

// positions a move selected (what select_non_tabu_ids pushed into the tabu deque)
__device__ __forceinline__ int gj_move_selected(const GjMove& m, int* out) {
    if (m.kind == GJ_MOVE_NULL) return 0;
    if (m.kind == 3) { out[0] = m.a[0]; return 1; }
    const int k = (m.kind >= 4) ? 2 : m.k;
#pragma unroll
    for (int i = 0; i < GJ_MOVE_MAXK; ++i) out[i] = m.a[i];
    return k;
}

// update_top_individual (agent_base.rs:220-224): population[0] <= agent_top -> replace.
// Cooperative over the CTA; `dirty` marks islands whose population[0] changed.
__device__ __forceinline__ void gj_update_top(int island, int levels, int stride, int n_vars,
                                              const int32_t* cur, const double* cur_score,
                                              int32_t* best, double* best_score, int* dirty) {
    if (dirty[island]) {
        GjScore c = gj_load_score(cur_score + (size_t)island * GJ_MAX_LEVELS, levels);
        GjScore top = gj_load_score(best_score + (size_t)island * GJ_MAX_LEVELS, levels);
        if (gj_score_le(c, top, levels)) {
            const int32_t* cur_row = cur + (size_t)island * stride;
            int32_t* best_row = best + (size_t)island * stride;
            for (int i = threadIdx.x; i < n_vars; i += blockDim.x) best_row[i] = cur_row[i];
            if (threadIdx.x == 0)
                for (int l = 0; l < GJ_MAX_LEVELS; ++l) best_score[(size_t)island * GJ_MAX_LEVELS + l] = c.v[l];
        }
        __syncthreads();
        if (threadIdx.x == 0) dirty[island] = 0;
    }
}

// Rebuilds one group's tabu table (membership bits, then the exclusive prefix count of free
// positions per word; layout in gj_moves.cuh) from its deque.  Cooperative over the CTA;
// `scan` holds max(blockDim, 33) ints of shared memory.
__device__ __forceinline__ void gj_tabu_table_rebuild(uint32_t* table, int glen, const int32_t* ring,
                                                      int fill, int* scan) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int W = (glen + 31) >> 5;
    int32_t* prefix = (int32_t*)(table + W + 1);
    __syncthreads();
    for (int w = tid; w <= W; w += nthr) table[w] = 0u;
    __syncthreads();
    for (int i = tid; i < fill; i += nthr) {
        const int pos = ring[i];
        atomicOr(&table[pos >> 5], 1u << (pos & 31));
    }
    __syncthreads();
    int carry = 0;
    for (int base = 0; base < W; base += nthr) {
        const int w = base + tid;
        int f = 0;
        if (w < W) {
            const int rem = glen - 32 * w;
            const uint32_t valid = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
            f = __popc(~table[w] & valid);
        }
        int chunk_total;
        const int incl = gj_block_scan_incl(f, scan, &chunk_total);
        if (w < W) prefix[w] = carry + incl - f;
        carry += chunk_total;
        __syncthreads();
    }
    if (tid == 0) prefix[W] = carry;
    __syncthreads();
    // compact the free positions, ascending: position p lands at prefix[word] + (free bits below it)
    int32_t* free_list = (int32_t*)(table + 2 * (W + 1));
    for (int pos = tid; pos < glen; pos += nthr) {
        const int w = pos >> 5, b = pos & 31;
        const uint32_t fm = ~table[w];
        if ((fm >> b) & 1u) free_list[prefix[w] + __popc(fm & ((1u << b) - 1u))] = pos;
    }
    __syncthreads();
}

// Tabu deque update of one island after a step (Mover::select_non_tabu_ids, mover.rs:75-96): every id
// selected by the step's K moves is pushed to the front in candidate order; ids beyond the deque's
// size fall off the back.  The deque is stored by recency rank (slot 0 = newest), so only the newest
// `size` ids of the step are needed: walk the candidates backwards and stop once the deque is full.
// Cooperative over the CTA; `scan` holds blockDim ints of shared memory; load_move(j) yields move j.
template <class LoadMove>
__device__ __forceinline__ void gj_tabu_deque_advance(uint32_t* bits_rw, const int32_t* ring_old_island,
                                                      int32_t* ring_new_island, const int32_t* ring_off,
                                                      const int32_t* tabu_size, const int32_t* word_off,
                                                      int* fill_island, int n_groups, const GjGroups& G, int K,
                                                      LoadMove load_move, int* scan) {
    const int tid = threadIdx.x;
    const int n_chunks = (K + blockDim.x - 1) / blockDim.x;
    for (int g = 0; g < n_groups; ++g) {
        const int T = tabu_size[g];
        const int32_t* ring_old = ring_old_island + ring_off[g];
        int32_t* ring_new = ring_new_island + ring_off[g];
        const int glen = G.offsets[g + 1] - G.offsets[g];
        const int fill_old = fill_island[g];
        int collected = 0;
        for (int chunk = n_chunks - 1; chunk >= 0 && collected < T; --chunk) {
            const int j = chunk * blockDim.x + tid;
            int sel[GJ_MOVE_MAXK]; int cnt = 0;
            if (j < K) {
                const GjMove m = load_move(j);
                if (m.kind != GJ_MOVE_NULL && m.group == g) cnt = gj_move_selected(m, sel);
            }
            // ids pushed by later candidates of this chunk (exclusive suffix sum)
            int total;
            const int incl = gj_block_scan_incl(cnt, scan, &total);
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
        // older ids keep their order behind the new ones
        for (int r = collected + tid; r < T; r += blockDim.x) {
            const int rho = r - collected;
            if (rho < fill_old) ring_new[r] = ring_old[rho];
        }
        const int fill = min(T, fill_old + collected);
        __syncthreads();
        if (tid == 0) fill_island[g] = fill;
        gj_tabu_table_rebuild(bits_rw + word_off[g], glen, ring_new, fill, scan);
    }
}

// update_global_top, adopt half (agent_base.rs:465-489).  The reference re-evaluates
// `global.score < agent_top.score` after EVERY step and re-assigns population[0] = global each time
// it holds (LateAcceptance also pushes the score it leaves behind on late_scores, every time).
// TabuSearch (only with compare_to_global): after one step population[0] <= global and
// update_top_individual brings agent_top down with it, so "once per published version" is the same
// thing and saves the row copy.  LateAcceptance / SimulatedAnnealing may accept a WORSE neighbour
// right after adopting; the reference then resets them to the global top again (and again) until a
// step rejects or improves -- so for those agents the test is made every step, un-gated.
// Decision and bookkeeping (score, late list, flags) by ONE thread; returns whether the island's
// solution is to be replaced by the gbest row.
__device__ __forceinline__ bool gj_adopt_decide(const GjSelectArgs& A, int island) {
    const int ver = *A.gver;
    if (ver == 0) return false;                              // nothing published yet (stub score)
    if (A.chain_mode) {
        // chains (k_la_chains / k_vrp_chains) make this test themselves when a launch starts; a host
        // read between launches lands here first.  gseen[island] == version <=> the stored solution
        // is that version's gbest row (see gj_islands_chain.cuh).
        if (A.gseen[island] == ver && A.dirty[island] != 1) return false;
    } else if (A.agent == GJ_AGENT_TABU_SEARCH) {
        if (ver == A.gseen[island]) return false;
        A.gseen[island] = ver;
    } else {
        // once per step index: the hook may be reached twice for the same step (a host read or a
        // trace applies pending adoptions before the step's own staging phase does)
        const int epoch = (int)(A.step & 0x3fffffffu) + 1;
        if (A.gseen[island] == epoch) return false;
        A.gseen[island] = epoch;
    }
    const GjScore g = gj_load_score(A.gbest_score, A.levels);
    const GjScore top = gj_load_score(A.best_score + (size_t)island * GJ_MAX_LEVELS, A.levels);
    const bool take = !gj_score_le(top, g, A.levels) && A.compare_to_global;     // global < agent_top
    if (!take) return false;
    if (A.agent == GJ_AGENT_LATE_ACCEPTANCE) {
        double* lt = A.late + (size_t)island * A.late_size * GJ_MAX_LEVELS;
        const int head = (A.late_head[island] + A.late_size - 1) % A.late_size;
        for (int l = 0; l < GJ_MAX_LEVELS; ++l)
            lt[(size_t)head * GJ_MAX_LEVELS + l] = A.cur_score[(size_t)island * GJ_MAX_LEVELS + l];
        A.late_head[island] = head; A.late_len[island] = min(A.late_len[island] + 1, A.late_size);
    }
    for (int l = 0; l < GJ_MAX_LEVELS; ++l) A.cur_score[(size_t)island * GJ_MAX_LEVELS + l] = g.v[l];
    if (A.chain_mode) {
        A.gseen[island] = ver;
        A.dirty[island] = 2;        // replaced, and already adopted (1 = replaced by a migrant)
    } else {
        A.dirty[island] = 1;
    }
    if (A.stale) A.stale[island] = 1;
    return true;
}

__device__ __forceinline__ size_t gj_slot_bytes(int stride) { return (size_t)stride * 4 + GJ_MAX_LEVELS * 8; }

struct GjChainArgs {
    int agent;                  // GJ_AGENT_LATE_ACCEPTANCE / GJ_AGENT_SIMULATED_ANNEALING
    double* sa_temp; GjSaParams sa; double* trace_aux;
    int I, stride, n_vars, late_size, noop, n_groups, symmetric, island_base;
    GjMoverParams M;
    uint64_t seed, step0;
    int n_steps;
    int32_t* cur; double* cur_score;
    int32_t* best; double* best_score;
    int* dirty;
    double* late; int* late_head; int* late_len;
    unsigned long long* counters;
    // chain tabu state, per island and group (global, persistent):
    //   bits [W + 1] words | ring [T] ints | head, fill
    uint32_t* ctabu; int ctabu_words_per_island; const int32_t* ctabu_off; const int32_t* tabu_size;
    // published global top (one-CTA k_global_top); adopted here, at the start of a launch
    const int32_t* gbest; const double* gbest_score; const int* gver; int* gseen;
    // trace (tests): [n_steps][I]
    GjMove* trace_moves; double* trace_scores; int* trace_accept;
};

// update_global_top, publish half (agent_base.rs:451-461), ONE CTA.  Agent tops only ever improve
// and the global top is refreshed from them after every step, so the new global top is simply the
// best agent top (first index on ties); it replaces the published one when strictly better (:451)
// and bumps the version that gj_adopt_decide (the adopt half) watches.
// Cooperative over ONE CTA (k_global_top, or the last CTA of a fused step to finish -- see
// gj_islands_tsfast.cuh); sh_s / sh_i: 32 entries of shared memory each.
__device__ __forceinline__ void gj_global_top_cta(int I, int levels, int stride, int n_vars,
                                                  const int32_t* __restrict__ best, const double* __restrict__ best_score,
                                                  int32_t* gbest, double* gbest_score, int* gver,
                                                  GjScore* sh_s, int* sh_i, int* sh_publish_p) {
    int& sh_publish = *sh_publish_p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    GjScore mine; int mine_idx = -1;
    mine.v[0] = mine.v[1] = mine.v[2] = 0.0;
    for (int i = tid; i < I; i += blockDim.x) {
        GjScore s;                              // L2 loads: another SM may have written the score this launch
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) s.v[l] = (l < levels) ? __ldcg(best_score + (size_t)i * GJ_MAX_LEVELS + l) : 0.0;
        if (mine_idx < 0 || gj_score_cmp(s, mine, levels) < 0) { mine = s; mine_idx = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        GjScore other; const int oidx = __shfl_xor_sync(GJ_FULL_MASK, mine_idx, o);
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) other.v[l] = __shfl_xor_sync(GJ_FULL_MASK, mine.v[l], o);
        if (oidx >= 0) {
            const int c = (mine_idx < 0) ? 1 : gj_score_cmp(mine, other, levels);
            if (c > 0 || (c == 0 && oidx < mine_idx)) { mine = other; mine_idx = oidx; }
        }
    }
    if (lane == 0) { sh_s[warp] = mine; sh_i[warp] = mine_idx; }
    __syncthreads();
    if (tid == 0) {
        GjScore b = sh_s[0]; int bi = sh_i[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            if (sh_i[w] < 0) continue;
            const int c = (bi < 0) ? 1 : gj_score_cmp(b, sh_s[w], levels);
            if (c > 0 || (c == 0 && sh_i[w] < bi)) { b = sh_s[w]; bi = sh_i[w]; }
        }
        sh_i[0] = bi;
        sh_publish = 0;
        const GjScore g = gj_load_score(gbest_score, levels);
        if (bi >= 0 && !gj_score_le(g, b, levels)) {            // strict: agent_top < global (:451)
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) gbest_score[l] = b.v[l];
            sh_publish = 1;
            *gver += 1;
        }
    }
    __syncthreads();
    if (sh_publish) {
        const int32_t* win_row = best + (size_t)sh_i[0] * stride;
        for (int i = tid; i < n_vars; i += blockDim.x) gbest[i] = __ldcg(&win_row[i]);
    }
}

// ---- shared-memory plans / launch constants the host side needs when it picks a step path ----------
__host__ __device__ inline size_t gj_fused_smem_bytes_lean(int n_vars, int words) {
    const size_t n_pad = ((size_t)n_vars + 3) & ~(size_t)3;
    return (((n_pad + 8) * 4 + (size_t)words * 4) + 15) & ~(size_t)15;
}

__host__ __device__ inline size_t gj_fused_smem_bytes(int n_vars, int cnt_stride, int tabu_words,
                                                      int words, int n_clone) {
    const size_t n_pad = ((size_t)n_vars + 3) & ~(size_t)3;
    size_t b = (n_pad + 8) * 4 + (size_t)cnt_stride * 4 + (((size_t)tabu_words + 3) & ~(size_t)3) * 4;
    b += (size_t)n_clone * (size_t)words * 4 + (size_t)n_clone * n_pad * 4;
    b = (b + 15) & ~(size_t)15;
    b += ((size_t)n_vars + 1) * 8;
    return b;
}

__host__ __device__ inline size_t gj_chain_smem_bytes(int n_vars, int words, int ctabu_words, int late_size,
                                                      bool tsp) {
    const size_t n_pad = ((size_t)n_vars + 3) & ~(size_t)3;
    const size_t scratch = n_pad < 32 ? 32 : n_pad;      // also holds <= 17 ints of a small move's columns
    size_t b = (n_pad + 8) * 4 + (size_t)32 * words * 4 + scratch * 4 + (size_t)words * 4;
    b += (((size_t)ctabu_words + 3) & ~(size_t)3) * 4;
    b = (b + 15) & ~(size_t)15;
    if (tsp) b += (((size_t)n_vars + 2) & ~(size_t)1) * 8;
    b += (size_t)late_size * (tsp ? 2 : 1) * 8;          // the late list, one entry per level of the model
    return (b + 15) & ~(size_t)15;
}

static constexpr int kChainWarps = 4;          // k_la_chains: chains (warps) per CTA, narrow variant
static constexpr int kChainWarpsWide = 28;     // ... wide variant: one CTA per SM holds every chain of the SM (<= 72 registers)
#define GJ_VRPC_DIFF 512          // stops the agent's top may trail the chain by before a whole-row copy

// static + dynamic shared memory beyond 48 KB needs the opt-in; kernels here carry up to ~19 KB of
// static shared memory
template <class Kern>
static gj_status opt_in_smem(Kern kernel, size_t bytes) {
    if (bytes > 24 * 1024)
        GJ_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return GJ_OK;
}

// host-side builders / launchers that cross translation units
GjSelectArgs gj_make_select_args(gj_islands* g, bool trace, bool stored_moves);
gj_status gj_launch_fused_step(gj_islands* g, cudaStream_t st, bool trace);                  // gj_islands_fused.cu
gj_status gj_launch_tsfast_step(gj_islands* g, cudaStream_t st, bool trace);                 // gj_islands_tsfast.cu
size_t gj_tsfast_smem(const gj_islands* g);
gj_status gj_launch_la_chains(gj_islands* g, const GjChainArgs& A, cudaStream_t st);          // gj_islands_chain.cu
gj_status gj_launch_vrp_chains(gj_islands* g, const GjChainArgs& A, cudaStream_t st);         // gj_islands_vrp_chain.cu
gj_status gj_launch_vrp_gindex(gj_islands* g, cudaStream_t st);
