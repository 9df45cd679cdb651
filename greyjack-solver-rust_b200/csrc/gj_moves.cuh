// gj_moves.cuh -- the reference's Mover (greyjack/src/agents/metaheuristic_bases/mover.rs)
// on the device: a move is a small descriptor (kind, semantic group, chosen group
// positions, new values) generated from a counter-based RNG; it is expanded on the fly,
// either into the shared-memory clone of the base that the scorer evaluates or into the
// (column, value) list the reference's incremental form would have emitted.
#pragma once

#include "gj_device.cuh"

#define GJ_MOVE_MAXK 8          // max chosen positions of a change / swap / swap_edges move
#define GJ_MOVE_MAXPAIRS 16     // max (column, value) pairs of a small move

struct GjMove {
    uint8_t kind;               // GJ_MOVE_*; 255 = null move (reference returns None)
    uint8_t group;              // semantic group index
    uint8_t k;                  // number of chosen positions
    uint8_t pad;
    int32_t a[GJ_MOVE_MAXK];    // chosen group positions (scramble: a[0] = start;
                                // insertion: get_out, put_in; inverse: the two ends)
    int32_t v[GJ_MOVE_MAXK];    // change: new decoded values; scramble: permutation
};

struct GjGroups {               // VariablesManager::semantic_groups_map on the device
    int n_groups;
    const int32_t* offsets;     // [n_groups + 1]
    const int32_t* ids;         // variable ids, group after group
    const int4* info;           // per group {first, step, uniform, 0}: ids[k] == first + k*step when
                                // step != 0 (affine group); uniform = every column of the group has
                                // the same integer bounds (moving values inside it never clamps)
};

enum { GJ_MOVE_NULL = 255 };

// ---- small moves: (column, value) pairs in the reference's emission order ----------------
// `rd(var_id)` reads the base.  incremental = the ISC form (TS / LA); otherwise the result
// of the plain form's sequential swaps, expressed as final (column, value) pairs.
// Returns the number of pairs; later pairs win on repeated columns.
// ids of an affine semantic group (GjGroups::info): position k is variable first + k * step
struct GjAffineIds {
    int first, step;
    __device__ __forceinline__ int operator[](int k) const { return first + k * step; }
};

// `g[k]`: variable id of group position k (the group's id list, or GjAffineIds)
template <class Ids, class Rd>
__device__ __forceinline__ int gj_small_move_pairs(const GjMove& m, Ids g,
                                                   bool incremental, bool noop_quirk, Rd rd,
                                                   int* cols, int* vals) {
    const int k = m.k;
    switch (m.kind) {
        case 0: {   // change_move, mover.rs:145-178
            for (int i = 0; i < k; ++i) { cols[i] = g[m.a[i]]; vals[i] = m.v[i]; }
            return k;
        }
        case 1: {   // swap_move, mover.rs:180-219: cyclic rotation over the chosen columns
            for (int i = 0; i < k; ++i) cols[i] = g[m.a[i]];
            if (incremental) {
                for (int i = 0; i < k; ++i) vals[i] = rd(cols[i]);
                for (int i = 1; i < k; ++i) { int t = vals[i - 1]; vals[i - 1] = vals[i]; vals[i] = t; }
                return k;
            }
            // plain: candidate.swap(c[i-1], c[i]) sequentially == the same rotation, because
            // the chosen columns are distinct (choice without replacement / non-tabu ids)
            for (int i = 0; i < k; ++i) vals[i] = rd(cols[i]);
            for (int i = 1; i < k; ++i) { int t = vals[i - 1]; vals[i - 1] = vals[i]; vals[i] = t; }
            return k;
        }
        case 2: {   // swap_edges_move, mover.rs:221-277
            int e0[GJ_MOVE_MAXK], e1[GJ_MOVE_MAXK];
            for (int i = 0; i < k; ++i) {
                e0[i] = g[m.a[i]]; e1[i] = g[m.a[i] + 1];
                cols[2 * i] = e0[i]; cols[2 * i + 1] = e1[i];
            }
            // edges.rotate_left(1)
            { int f0 = e0[0], f1 = e1[0];
              for (int i = 0; i + 1 < k; ++i) { e0[i] = e0[i + 1]; e1[i] = e1[i + 1]; }
              e0[k - 1] = f0; e1[k - 1] = f1; }
            if (incremental && noop_quirk) {
                for (int i = 0; i < k; ++i) { vals[2 * i] = rd(e0[i]); vals[2 * i + 1] = rd(e1[i]); }
                for (int i = 1; i < k; ++i) {
                    int t = vals[2 * (i - 1)]; vals[2 * (i - 1)] = vals[2 * i]; vals[2 * i] = t;
                    t = vals[2 * (i - 1) + 1]; vals[2 * (i - 1) + 1] = vals[2 * i + 1]; vals[2 * i + 1] = t;
                }
                return 2 * k;
            }
            // plain form: sequential swaps over the rotated edge list, replayed on the window
            for (int i = 0; i < 2 * k; ++i) vals[i] = rd(cols[i]);
            auto sw = [&](int ca, int cb) {
                int pa = -1, pb = -1;
                for (int j = 0; j < 2 * k; ++j) { if (cols[j] == ca) pa = j; if (cols[j] == cb) pb = j; }
                int t = vals[pa]; vals[pa] = vals[pb]; vals[pb] = t;
                for (int j = 0; j < 2 * k; ++j) {
                    if (cols[j] == ca) vals[j] = vals[pa];
                    if (cols[j] == cb) vals[j] = vals[pb];
                }
            };
            for (int i = 1; i < k; ++i) { sw(e0[i - 1], e0[i]); sw(e1[i - 1], e1[i]); }
            return 2 * k;
        }
        case 3: {   // scramble_move, mover.rs:279-317
            const int start = m.a[0];
            int native[GJ_MOVE_MAXK], scr[GJ_MOVE_MAXK];
            for (int i = 0; i < k; ++i) native[i] = g[start + i];
            for (int i = 0; i < k; ++i) scr[i] = native[m.v[i]];
            if (incremental && noop_quirk) {     // :306-309: every column keeps its own value
                for (int i = 0; i < k; ++i) { cols[i] = scr[i]; vals[i] = rd(scr[i]); }
                return k;
            }
            for (int i = 0; i < k; ++i) { cols[i] = native[i]; vals[i] = rd(native[i]); }
            for (int i = 0; i < k; ++i) {        // candidate.swap(native[i], scrambled[i])
                const int pa = i, pb = m.v[i];
                int t = vals[pa]; vals[pa] = vals[pb]; vals[pb] = t;
            }
            return k;
        }
        default: return 0;
    }
}

// ---- segment moves (insertion / inverse): value of segment slot t after the move ----------
// The segment is group positions [lo, hi]; src_slot(t) gives the slot whose BASE value
// lands in slot t.  mover.rs:319-376 (insertion) and :378-420 (inverse).
__device__ __forceinline__ void gj_segment_bounds(const GjMove& m, int& lo, int& hi) {
    lo = min(m.a[0], m.a[1]);
    hi = max(m.a[0], m.a[1]);
}

__device__ __forceinline__ int gj_segment_src_slot(const GjMove& m, bool incremental, int t, int len) {
    if (m.kind == 5) return len - 1 - t;                     // inverse: both forms agree
    const bool left_rotate = m.a[0] < m.a[1];                // get_out < put_in
    if (incremental) {                                       // rotate by one
        return left_rotate ? ((t + 1 == len) ? 0 : t + 1) : ((t == 0) ? len - 1 : t - 1);
    }
    // plain form = a chain of swaps over (old_ids, shifted_ids):
    //   left : [v0, v2, v3, .., v_{m-1}, v1]
    //   right: [v1, v2, .., v_{m-2}, v0, v_{m-1}]
    if (len == 2) return t;
    if (left_rotate) {
        if (t == 0) return 0;
        if (t == len - 1) return 1;
        return t + 1;
    }
    if (t == len - 1) return len - 1;
    if (t == len - 2) return 0;
    return t + 1;
}

// Applies a move: dst(var_id, value) receives the changed columns; rd(var_id) reads the
// base.  Cooperative over `nthr` threads with rank `tid` (one warp or one CTA); small
// moves are expanded by tid 0 (<= 16 pairs), segment moves in parallel.  The caller
// synchronises afterwards.
// VariablesManager::fix_deltas / fix_variables on a changed column (variables_manager.rs:
// 187-220): the value is clamped into the destination variable's own bounds (matters when a
// semantic group mixes variables with different domains, e.g. the VRP "common" group).
__device__ __forceinline__ int gj_fix_column(const GjProblemDev& P, int col, int v) {
    return min(max(v, P.lbi[col]), P.ubi[col]);
}

template <class Rd, class Wr>
__device__ __forceinline__ void gj_apply_move(const GjProblemDev& P, const GjMove& m, const GjGroups& G,
                                              bool incremental, bool noop_quirk, int tid, int nthr,
                                              Rd rd, Wr wr) {
    if (m.kind == GJ_MOVE_NULL) return;
    const int32_t* g = G.ids + G.offsets[m.group];
    if (m.kind <= 3) {
        if (tid == 0) {
            int cols[GJ_MOVE_MAXPAIRS], vals[GJ_MOVE_MAXPAIRS];
            const int n = gj_small_move_pairs(m, g, incremental, noop_quirk, rd, cols, vals);
            for (int i = 0; i < n; ++i) wr(cols[i], gj_fix_column(P, cols[i], vals[i]));
        }
        return;
    }
    int lo, hi;
    gj_segment_bounds(m, lo, hi);
    const int len = hi - lo + 1;
    for (int t = tid; t < len; t += nthr) {
        const int s = gj_segment_src_slot(m, incremental, t, len);
        const int col = g[lo + t];
        wr(col, gj_fix_column(P, col, rd(g[lo + s])));
    }
}

// ---- generation -----------------------------------------------------------------------------

struct GjMoverParams {
    double thresholds[6];       // cumulative move probabilities (mover.rs:36-62)
    double tabu_entity_rate;
    double mutation_rate_multiplier;   // 0 when None
    int tabu_layout;            // 0: free-position table (gj_tabu_view), 1: chain bitmap
    int pad;
};

// Binomial(n, p) by CDF inversion, truncated at kmax (p*n is O(1) in every configuration:
// group_mutation_rate = multiplier / group_size, mover.rs:138-140).
__device__ __forceinline__ int gj_binomial_small(GjPhilox& rng, int n, double p, int kmax) {
    if (p <= 0.0) return 0;
    if (p >= 1.0) return min(n, kmax);
    const double u = gj_rng_f64(rng);
    double pk = pow(1.0 - p, (double)n);
    double cdf = pk;
    const double ratio = p / (1.0 - p);
    int k = 0;
    while (u > cdf && k < kmax && k < n) {
        pk = pk * ratio * (double)(n - k) / (double)(k + 1);
        cdf += pk;
        ++k;
    }
    return k;
}

// ---- tabu table ---------------------------------------------------------------------------------
// Per island and semantic group (W = ceil(group_len / 32)):
//     bits   [W + 1] words   membership (bit set = position is in the tabu deque)
//     prefix [W + 1] ints    exclusive count of FREE positions per word; prefix[W] = all of them
//     free   [group_len] ints the free positions in ascending order (prefix[W] valid entries)
// The list turns "draw until the id is not tabu" (Mover::select_non_tabu_ids :75-96) into one
// draw and one load: the r-th free position, r ~ U[0, free) -- the same distribution without a
// data-dependent retry loop.  bits / prefix are what the list is compacted from once per step.
struct GjTabuView {
    const int32_t* free;        // free-position list (TabuSearch islands); nullptr otherwise
    int n_free;
    int glen;
    const uint32_t* bits;       // membership bits only (LateAcceptance chains, whose deque moves
                                // every step: ids are drawn by rejection); nullptr otherwise
};

__host__ __device__ inline int gj_tabu_region_words(int glen) {
    return 2 * (((glen + 31) >> 5) + 1) + glen;
}

__device__ __forceinline__ GjTabuView gj_tabu_view(const uint32_t* table, int glen, int layout) {
    GjTabuView v;
    const int W = (glen + 31) >> 5;
    v.free = nullptr; v.n_free = 0; v.bits = nullptr; v.glen = glen;
    if (table && layout == 0) {
        v.free = (const int32_t*)(table + 2 * (W + 1));
        v.n_free = ((const int32_t*)table)[2 * W + 1];              // prefix[W]
    } else if (table) {
        v.bits = table;
    }
    return v;
}

// number of free positions below right_end (right_end is within a few positions of group_len)
__device__ __forceinline__ int gj_tabu_free_below(const GjTabuView& tv, int right_end) {
    int f = tv.n_free;
    if (right_end >= tv.glen) return f;           // the whole group: nothing to trim
    while (f > 0 && tv.free[f - 1] >= right_end) --f;
    return f;
}

// k distinct positions in [0, right_end) outside the island's tabu snapshot
// (Mover::select_non_tabu_ids :75-96; with tabu_entity_rate == 0 this is math_utils::choice
// without replacement, :43-45): the i-th pick draws a rank among the free positions not yet
// taken.  Loops are unrolled over the fixed maximum so that everything stays in registers.
__device__ __forceinline__ void gj_pick_positions(GjPhilox& rng, int right_end, int k,
                                                  const GjTabuView& tv, int32_t* out) {
    if (tv.bits) {
        // rejection against the bitmap: draw until the id is neither tabu nor already chosen
        // (select_non_tabu_ids :75-96 verbatim; the tabu test is dropped after 48 tries so that a
        // nearly-full deque cannot stall the chain)
#pragma unroll
        for (int i = 0; i < GJ_MOVE_MAXK; ++i) {
            if (i < k) {
                int pos = 0;
                for (int tries = 0; tries < 64; ++tries) {
                    pos = (int)gj_rng_below(rng, (uint32_t)right_end);
                    bool clash = false;
#pragma unroll
                    for (int j = 0; j < GJ_MOVE_MAXK; ++j)
                        if (j < i) clash |= (out[j] == pos);
                    if (!clash && ((tv.bits[pos >> 5] >> (pos & 31)) & 1u) && tries < 48) clash = true;
                    if (!clash) break;
                }
                out[i] = pos;
            }
        }
        return;
    }
    int F = right_end;
    bool use_tabu = false;
    if (tv.free) {
        const int f = gj_tabu_free_below(tv, right_end);
        if (f >= k) { F = f; use_tabu = true; }     // a fully tabu group falls back to plain choice
    }
    if (k == 2) {                                  // the common case, straight-line
        const int r0 = (int)gj_rng_below(rng, (uint32_t)F);
        int r1 = (int)gj_rng_below(rng, (uint32_t)(F - 1));
        if (r1 >= r0) r1 += 1;
        out[0] = use_tabu ? tv.free[r0] : r0;
        out[1] = use_tabu ? tv.free[r1] : r1;
        return;
    }
    int sorted[GJ_MOVE_MAXK];
#pragma unroll
    for (int i = 0; i < GJ_MOVE_MAXK; ++i) sorted[i] = 0;
#pragma unroll
    for (int i = 0; i < GJ_MOVE_MAXK; ++i) {
        if (i < k) {
            int r = (int)gj_rng_below(rng, (uint32_t)(F - i));
#pragma unroll
            for (int j = 0; j < GJ_MOVE_MAXK; ++j)
                if (j < i && r >= sorted[j]) r += 1;          // skip the ranks already taken (ascending)
            int ins = r;
#pragma unroll
            for (int j = 0; j < GJ_MOVE_MAXK; ++j)
                if (j < i && sorted[j] > ins) { const int t = sorted[j]; sorted[j] = ins; ins = t; }
            sorted[i] = ins;
            out[i] = use_tabu ? tv.free[r] : r;
        }
    }
}

// Mover::do_move (mover.rs:98-128) for candidate `cand` of island `island` at step `step`.
// COMPACT_RNG: same stream, Philox refills out of line (smaller code; see GjPhilox::compact).
template <bool COMPACT_RNG = false>
__device__ __forceinline__ GjMove gj_generate_move(const GjProblemDev& P, const GjGroups& G,
                                                   const GjMoverParams& M, uint64_t seed,
                                                   uint32_t island, uint64_t step, uint32_t cand,
                                                   const uint32_t* tabu_island_bits,
                                                   const int32_t* tabu_word_off) {
    GjPhilox rng;
    gj_rng_init(rng, seed, island, (uint32_t)step, (uint32_t)(step >> 32), cand);
    rng.compact = COMPACT_RNG ? 1 : 0;
    GjMove m;
    m.pad = 0;
#pragma unroll
    for (int i = 0; i < GJ_MOVE_MAXK; ++i) { m.a[i] = 0; m.v[i] = 0; }
    const double u = (double)gj_rng_u32(rng) * (1.0 / 4294967296.0);
    // first i with u <= thresholds[i] (mover.rs:105-121); thresholds are cumulative, so that is
    // the number of thresholds below u -- no branches
    int kind = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) kind += (u > M.thresholds[i]) ? 1 : 0;
    const int grp = (int)gj_rng_below(rng, (uint32_t)G.n_groups);
    const int glen = G.offsets[grp + 1] - G.offsets[grp];
    const int32_t* g = G.ids + G.offsets[grp];
    const GjTabuView tabu = gj_tabu_view((M.tabu_entity_rate != 0.0 && tabu_island_bits)
                                             ? tabu_island_bits + tabu_word_off[grp] : nullptr, glen,
                                         M.tabu_layout);
    m.kind = (uint8_t)kind;
    m.group = (uint8_t)grp;
    m.k = 0;
    const double rate = (glen > 0 && M.mutation_rate_multiplier != 0.0)
                            ? M.mutation_rate_multiplier * (1.0 / (double)glen) : 0.0;
    // Per kind: how many positions to draw and from which range.  ONE gj_pick_positions call site
    // serves every kind, so a warp whose lanes drew different kinds does not replay the sampler.
    int k = 2, right_end = glen, count = 0;
    bool null_move = false;
    if (kind <= 2) {
        // get_necessary_info_for_move: change count ~ Binomial(n_vars, group rate)
        k = gj_binomial_small(rng, P.n_vars, rate, GJ_MOVE_MAXK);
        if (kind == 0) {                    // change_move, mover.rs:145-178
            if (k < 1) k = 1;
            null_move = glen < k;
        } else if (kind == 1) {             // swap_move, :180-219
            if (k < 2) k = 2;
            null_move = glen < k;
        } else {                            // swap_edges_move, :221-277
            if (k < 2) k = 2;
            if (k > glen - 1) k = glen - 1;
            null_move = (glen == 0) || (k < 1);
            right_end = glen - 1;
        }
    } else if (kind == 3) {                 // scramble_move, :279-317: a window of 3..6
        count = 3 + (int)gj_rng_below(rng, 4);
        null_move = (glen < count - 1) || (glen - count <= 0);
        right_end = glen - count;
        k = 1;
    } else {                                // insertion_move :319-376 / inverse_move :378-420
        null_move = glen <= 1;
    }
    if (null_move) { m.kind = GJ_MOVE_NULL; return m; }
    gj_pick_positions(rng, right_end, k, tabu, m.a);
    m.k = (uint8_t)k;
    if (kind == 0) {
#pragma unroll
        for (int i = 0; i < GJ_MOVE_MAXK; ++i) {
            if (i < k) {
                // get_column_random_value: Uniform::new(lb, ub), then fix_deltas (clamp + rint)
                const int var = g[m.a[i]];
                const double lb = P.lb[var], ub = P.ub[var];
                const double x = lb + gj_rng_f64(rng) * (ub - lb);
                m.v[i] = gj_decode(P, var, x);
            }
        }
    } else if (kind == 3) {
#pragma unroll
        for (int i = 0; i < GJ_MOVE_MAXK; ++i) m.v[i] = (i < count) ? i : 0;
#pragma unroll
        for (int i = GJ_MOVE_MAXK - 1; i > 0; --i) {   // Fisher-Yates over v[0..count)
            if (i < count) {
                const int r = (int)gj_rng_below(rng, (uint32_t)(i + 1));
                const int vi = m.v[i];
                int vr = vi;
#pragma unroll
                for (int j = 0; j < GJ_MOVE_MAXK; ++j)
                    if (j == r) { vr = m.v[j]; m.v[j] = vi; }
                m.v[i] = vr;
            }
        }
        m.k = (uint8_t)count;
    }
    return m;
}
