// gj_delta.cuh -- delta evaluation of a move against an island's cached base state.
//
// The reference's "incremental" score calculators are pseudo-incremental: every neighbour is
// materialised and fully re-scored (tsp ISC :58-86, nqueens ISC :36-59).  Here an
// island keeps, next to its current solution,
//     cnt[key]  occurrence count of every counted value (TSP: location ids; N-Queens: rows,
//               col+row and col-row diagonals in three disjoint key ranges),
//     raw[]     the unrounded constraint terms of the current solution (TSP: dup count, tour
//               length), produced by the FULL evaluator in the reference's summation order,
// and one thread scores one neighbour from the O(k) terms its move touches:
//     integer levels : uniq' = uniq + sum over touched keys [(cnt+net > 0) - (cnt > 0)]  (exact)
//     TSP distance   : dist' = dist + (sum of new edges) - (sum of removed edges)        (f64)
// The float level therefore differs from a full re-evaluation only by summation order
// (<= 1e-12 relative before ScoreTrait::round, one 10^-precision quantum after); the integer
// levels are bit-exact.  Every function returns false for a move it does not cover; the caller
// queues that neighbour for the full evaluator.
#pragma once

#include "gj_eval.cuh"
#include "gj_moves.cuh"

struct GjDeltaState {
    int32_t* cnt;        // [I][cnt_stride]
    int cnt_stride;
    double* raw;         // [I][GJ_MAX_LEVELS] unrounded, unweighted terms of the current solution
    int* stale;          // [I] 1 = cur changed since the state was built
    double* edge;        // [I][n + 1] TSP: scratch for the exact-order fold of k_refresh
};

// Change of the number of distinct keys when, for every i < m with ko[i] != kn[i], one
// occurrence moves from key ko[i] to key kn[i] (negative key = skip).  O(m^2), m <= 16.
__device__ __forceinline__ int gj_uniq_delta(const int32_t* __restrict__ cnt, const int* ko,
                                             const int* kn, int m) {
    int d = 0;
    for (int j = 0; j < 2 * m; ++j) {
        const int key = (j < m) ? ko[j] : kn[j - m];
        if (key < 0) continue;
        bool seen = false;
        int net = 0;
        for (int q = 0; q < 2 * m; ++q) {
            const int other = (q < m) ? ko[q] : kn[q - m];
            if (other == key) {
                if (q < j) { seen = true; break; }
                net += (q < m) ? -1 : 1;
            }
        }
        if (seen || net == 0) continue;
        const int before = cnt[key];
        d += ((before + net) > 0 ? 1 : 0) - (before > 0 ? 1 : 0);
    }
    return d;
}

// Same for exactly two moved occurrences (a swap of two entities), everything in registers.
__device__ __forceinline__ int gj_uniq_delta2(const int32_t* __restrict__ cnt, int o0, int o1,
                                              int n0, int n1) {
    if (o0 == n0 && o1 == n1) return 0;
    if (o0 == n1 && o1 == n0) return 0;               // the two keys just trade places
    int key[4] = {o0, o1, n0, n1};
    int d = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bool seen = false;
        int net = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (key[q] == key[j]) {
                if (q < j) seen = true;
                net += (q < 2) ? -1 : 1;
            }
        }
        if (!seen && net != 0) {
            const int before = cnt[key[j]];
            d += ((before + net) > 0 ? 1 : 0) - (before > 0 ? 1 : 0);
        }
    }
    return d;
}

// Drops pairs overridden by a later pair on the same column ("later pairs win").
__device__ __forceinline__ unsigned gj_live_mask(const int* cols, int m) {
    unsigned live = 0u;
    for (int i = 0; i < m; ++i) {
        bool dead = false;
        for (int j = i + 1; j < m; ++j) dead |= (cols[j] == cols[i]);
        if (!dead) live |= 1u << i;
    }
    return live;
}

// ---- TSP -------------------------------------------------------------------------------------
struct GjTspBase {
    const int32_t* __restrict__ t;     // current tour (decoded location ids per stop)
    int n;
    const double* __restrict__ D;
    size_t L;
    const double* edge;                // optional [n + 1]: edge[i] = D[at(i-1)][at(i)] of the current
                                       // tour (edge[n] closes it); nullptr -> gathered from D
    bool padded;                       // t[-1] and t[n] hold the depot (0): no bounds checks
    __device__ __forceinline__ double tour_edge(int i) const {
        return edge ? edge[i] : d(at(i - 1), at(i));
    }
    __device__ __forceinline__ int at(int q) const {
        if (padded) return t[q];
        return (q < 0 || q >= n) ? 0 : t[q];
    }
    __device__ __forceinline__ double d(int a, int b) const {
        // 32-bit index arithmetic while the matrix has fewer than 2^31 entries (L <= 46340)
        if (L <= 46340) return __ldg(&D[(unsigned)(a * (int)L + b)]);
        return __ldg(&D[(size_t)a * L + (size_t)b]);
    }
};

// (column, value) pairs of a small move -> change of the tour length and of the distinct count.
__device__ __forceinline__ void gj_tsp_pairs_delta(const GjProblemDev& P, const GjTspBase& B,
                                                   const int32_t* __restrict__ cnt, const int* cols,
                                                   const int* vals, int m, int& d_uniq,
                                                   double& removed, double& added) {
    const unsigned live = gj_live_mask(cols, m);
    auto in_set = [&](int q) {
        for (int i = 0; i < m; ++i) if (((live >> i) & 1u) && cols[i] == q) return true;
        return false;
    };
    auto nv = [&](int q) {
        for (int i = 0; i < m; ++i) if (((live >> i) & 1u) && cols[i] == q) return vals[i];
        return B.at(q);
    };
    removed = 0.0; added = 0.0;
    int ko[GJ_MOVE_MAXPAIRS], kn[GJ_MOVE_MAXPAIRS];
    for (int i = 0; i < m; ++i) {
        ko[i] = -1; kn[i] = -1;
        if (!((live >> i) & 1u)) continue;
        const int c = cols[i];
        const int oc = B.at(c), nc = vals[i];
        // every edge touching a changed column is counted once: (c-1, c) always, (c, c+1) only
        // when c+1 is not itself changed (it is then the left edge of c+1)
        const int ol = B.at(c - 1), nl = nv(c - 1);
        removed += B.d(ol, oc);
        added += B.d(nl, nc);
        if (!in_set(c + 1)) {
            const int r = B.at(c + 1);
            removed += B.d(oc, r);
            added += B.d(nc, r);
        }
        if (oc != nc) { ko[i] = oc - P.val_lo; kn[i] = nc - P.val_lo; }
    }
    d_uniq = gj_uniq_delta(cnt, ko, kn, m);
}

// Returns false when the move needs the full evaluator.
//   gi = {first, step, uniform_bounds, _} of the move's semantic group (affine groups only for
//   segment moves); symmetric = D[i][j] == D[j][i] bitwise (2-opt interior cancels).
__device__ __forceinline__ bool gj_tsp_move_delta(const GjProblemDev& P, const GjGroups& G,
                                                  const GjMove& m, bool noop_quirk, bool symmetric,
                                                  const GjTspBase& B, const int32_t* __restrict__ cnt,
                                                  int& d_uniq, double& d_dist) {
    d_uniq = 0; d_dist = 0.0;
    if (m.kind == GJ_MOVE_NULL) return true;
    // the reference's incremental scramble / swap_edges(k = 2) assign every column its own value
    // (SURVEY.md Q8): the neighbour IS the base
    if (noop_quirk && (m.kind == 3 || (m.kind == 2 && m.k == 2))) return true;
    const int32_t* g = G.ids + G.offsets[m.group];
    const int4 gi = G.info[m.group];
    // Fast path, branch-free over the move kind: a swap of two stops, a 2-opt reversal and an
    // insertion all replace at most four tour edges by four others.  Needs a group with uniform
    // bounds (no fix_deltas clamp, stop multiset unchanged); the segment moves also a group of
    // consecutive columns, the reversal a symmetric matrix (its interior edges then cancel).
    const bool two_swap = (m.kind == 1 && m.k == 2);
    const bool seg = (m.kind >= 4) && gi.y == 1 && (m.kind == 4 || symmetric);
    if (gi.z != 0 && (two_swap || seg)) {
        const int c0 = (gi.y != 0) ? gi.x + m.a[0] * gi.y : g[m.a[0]];
        const int c1 = (gi.y != 0) ? gi.x + m.a[1] * gi.y : g[m.a[1]];
        const int p = min(c0, c1), q = max(c0, c1);
        const int pm = B.at(p - 1), tp = B.at(p), pn = B.at(p + 1);
        const int qm = B.at(q - 1), tq = B.at(q), qp = B.at(q + 1);
        const bool adj = (q == p + 1);
        const bool inv = (m.kind == 5);
        const bool ins_l = (m.kind == 4) && (m.a[0] < m.a[1]);     // t[p] travels to the end
        const bool ins_r = (m.kind == 4) && !ins_l;                 // t[q] travels to the front
        const bool swp_far = two_swap && !adj;
        // removed: (pm,tp) (tq,qp) always; third: (tp,pn) [= (tp,tq) when adjacent] or, for a right
        // insertion, (qm,tq); fourth (far swap): (qm,tq)
        // added: (pm, x0) (y1, qp) always; third / fourth edge by kind
        const int x0 = ins_l ? pn : tq;
        const int y1 = ins_r ? qm : tp;
        const int a2b = swp_far ? pn : tp;
        // the removed edges are edges of the current tour: read from the island's edge cache
        // (the same f64 values the matrix holds); only the added edges gather from D
        double removed = B.tour_edge(p) + B.tour_edge(q + 1);
        double added = B.d(pm, x0) + B.d(y1, qp);
        if (!inv) { removed += B.tour_edge(ins_r ? q : p + 1); added += B.d(tq, a2b); }
        if (swp_far) { removed += B.tour_edge(q); added += B.d(qm, tp); }
        d_dist = added - removed;
        return true;
    }
    if (m.kind <= 3) {
        // general small move.  Works on a copy: the expansion indexes the descriptor dynamically,
        // and the caller's `m` must stay in registers for the fast path above.
        const GjMove ms = m;
        int cols[GJ_MOVE_MAXPAIRS], vals[GJ_MOVE_MAXPAIRS];
        const int np = gj_small_move_pairs(ms, g, true, noop_quirk,
                                           [&](int id) { return B.t[id]; }, cols, vals);
        for (int i = 0; i < np; ++i) vals[i] = gj_fix_column(P, cols[i], vals[i]);
        double removed, added;
        gj_tsp_pairs_delta(P, B, cnt, cols, vals, np, d_uniq, removed, added);
        d_dist = added - removed;
        return true;
    }
    // segment move outside the fast path (non-consecutive columns, per-column bounds, or a
    // reversal on an asymmetric matrix): O(segment) edges change -> full evaluator
    return false;
}

// ---- N-Queens --------------------------------------------------------------------------------
// key ranges inside cnt: rows [0, 32*bm_words), desc [.., +32*desc_words), asc [.., +32*asc_words)
__device__ __forceinline__ bool gj_nqueens_move_delta(const GjProblemDev& P, const GjGroups& G,
                                                      const GjMove& m, bool noop_quirk,
                                                      const int32_t* __restrict__ rows,
                                                      const int32_t* __restrict__ cnt, int& d_uniq) {
    d_uniq = 0;
    if (m.kind == GJ_MOVE_NULL) return true;
    if (noop_quirk && (m.kind == 3 || (m.kind == 2 && m.k == 2))) return true;   // see gj_tsp_move_delta
    if (m.kind > 3) return false;                      // O(segment) changes: full evaluator
    const int32_t* g = G.ids + G.offsets[m.group];
    const int4 gi = G.info[m.group];
    if (m.kind == 1 && m.k == 2 && gi.z != 0) {
        // swap of two queens' rows: the row multiset is unchanged, both diagonals move
        const int c0 = (gi.y != 0) ? gi.x + m.a[0] * gi.y : g[m.a[0]];
        const int c1 = (gi.y != 0) ? gi.x + m.a[1] * gi.y : g[m.a[1]];
        const int r0 = rows[c0], r1 = rows[c1];
        if (r0 == r1) return true;
        const int k0 = P.column_id[c0], k1 = P.column_id[c1];
        const int off_d = 32 * P.bm_words - P.desc_lo, off_a = 32 * (P.bm_words + P.desc_words) - P.asc_lo;
        d_uniq = gj_uniq_delta2(cnt, off_d + k0 + r0, off_d + k1 + r1, off_d + k0 + r1, off_d + k1 + r0) +
                 gj_uniq_delta2(cnt, off_a + k0 - r0, off_a + k1 - r1, off_a + k0 - r1, off_a + k1 - r0);
        return true;
    }
    const GjMove ms = m;                               // see gj_tsp_move_delta
    int cols[GJ_MOVE_MAXPAIRS], vals[GJ_MOVE_MAXPAIRS];
    const int np = gj_small_move_pairs(ms, g, true, noop_quirk,
                                       [&](int id) { return rows[id]; }, cols, vals);
    for (int i = 0; i < np; ++i) vals[i] = gj_fix_column(P, cols[i], vals[i]);
    const unsigned live = gj_live_mask(cols, np);
    const int off_desc = 32 * P.bm_words, off_asc = 32 * (P.bm_words + P.desc_words);
    int ko[GJ_MOVE_MAXPAIRS], kn[GJ_MOVE_MAXPAIRS];
    int d = 0;
#pragma unroll 1
    for (int space = 0; space < 3; ++space) {
        for (int i = 0; i < np; ++i) {
            ko[i] = -1; kn[i] = -1;
            if (!((live >> i) & 1u)) continue;
            const int c = cols[i], col = __ldg(&P.column_id[c]);
            const int o = rows[c], v = vals[i];
            if (o == v) continue;
            if (space == 0) { ko[i] = o - P.val_lo; kn[i] = v - P.val_lo; }
            else if (space == 1) { ko[i] = off_desc + (col + o - P.desc_lo); kn[i] = off_desc + (col + v - P.desc_lo); }
            else { ko[i] = off_asc + (col - o - P.asc_lo); kn[i] = off_asc + (col - v - P.asc_lo); }
        }
        d += gj_uniq_delta(cnt, ko, kn, np);
    }
    d_uniq = d;
    return true;
}
