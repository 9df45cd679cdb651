// gj_delta.cuh -- delta evaluation of a move against an island's cached base state.
//
// The reference's "incremental" score calculators are pseudo-incremental: every neighbour is
// materialised and fully re-scored (tsp ISC :58-86, nqueens ISC :36-59).  Here an
// island keeps, next to its current solution,
//     cnt[key]  occurrence count of every counted value (TSP: location ids; N-Queens: rows,
//               col+row and col-row diagonals in three disjoint key ranges),
//     uniq      number of keys with cnt > 0,
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
    int* uniq;           // [I]
    int* stale;          // [I] 1 = cur changed since the state was built
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
        const int before = __ldg(&cnt[key]);
        d += ((before + net) > 0 ? 1 : 0) - (before > 0 ? 1 : 0);
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
    __device__ __forceinline__ int at(int q) const { return (q < 0 || q >= n) ? 0 : __ldg(&t[q]); }
    __device__ __forceinline__ double d(int a, int b) const { return __ldg(&D[(size_t)a * L + (size_t)b]); }
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
    const int32_t* g = G.ids + G.offsets[m.group];
    if (m.kind <= 3) {
        int cols[GJ_MOVE_MAXPAIRS], vals[GJ_MOVE_MAXPAIRS];
        const int np = gj_small_move_pairs(m, g, true, noop_quirk,
                                           [&](int id) { return __ldg(&B.t[id]); }, cols, vals);
        for (int i = 0; i < np; ++i) vals[i] = gj_fix_column(P, cols[i], vals[i]);
        double removed, added;
        gj_tsp_pairs_delta(P, B, cnt, cols, vals, np, d_uniq, removed, added);
        d_dist = added - removed;
        return true;
    }
    const int4 gi = G.info[m.group];
    if (gi.y != 1 || gi.z == 0) return false;          // segment must be contiguous tour positions
    int lo, hi;
    gj_segment_bounds(m, lo, hi);
    const int a = gi.x + lo, b = gi.x + hi;
    const int ta = B.at(a), tb = B.at(b), pm = B.at(a - 1), pp = B.at(b + 1);
    if (m.kind == 5) {                                 // inverse_move (2-opt), mover.rs:378-420
        if (!symmetric) return false;
        const double removed = B.d(pm, ta) + B.d(tb, pp);
        const double added = B.d(pm, tb) + B.d(ta, pp);
        d_dist = added - removed;
        return true;
    }
    // insertion_move, incremental form (mover.rs:339-369): rotate the segment by one
    if (m.a[0] < m.a[1]) {                             // t[a] travels to the end
        const int tn = B.at(a + 1);
        const double removed = B.d(pm, ta) + B.d(ta, tn) + B.d(tb, pp);
        const double added = B.d(pm, tn) + B.d(tb, ta) + B.d(ta, pp);
        d_dist = added - removed;
    } else {                                           // t[b] travels to the front
        const int tq = B.at(b - 1);
        const double removed = B.d(pm, ta) + B.d(tq, tb) + B.d(tb, pp);
        const double added = B.d(pm, tb) + B.d(tb, ta) + B.d(tq, pp);
        d_dist = added - removed;
    }
    return true;
}

// ---- N-Queens --------------------------------------------------------------------------------
// key ranges inside cnt: rows [0, 32*bm_words), desc [.., +32*desc_words), asc [.., +32*asc_words)
__device__ __forceinline__ bool gj_nqueens_move_delta(const GjProblemDev& P, const GjGroups& G,
                                                      const GjMove& m, bool noop_quirk,
                                                      const int32_t* __restrict__ rows,
                                                      const int32_t* __restrict__ cnt, int& d_uniq) {
    d_uniq = 0;
    if (m.kind == GJ_MOVE_NULL) return true;
    if (m.kind > 3) return false;                      // O(segment) changes: full evaluator
    const int32_t* g = G.ids + G.offsets[m.group];
    int cols[GJ_MOVE_MAXPAIRS], vals[GJ_MOVE_MAXPAIRS];
    const int np = gj_small_move_pairs(m, g, true, noop_quirk,
                                       [&](int id) { return __ldg(&rows[id]); }, cols, vals);
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
            const int o = __ldg(&rows[c]), v = vals[i];
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
