// gj_eval.cuh -- full-evaluation device functions, one per constraint model of the
// reference's examples.  Each is written against a `Src` functor that yields the decoded
// value of planning variable i of ONE candidate, so the same arithmetic serves
//   * request_score_plain  (f64 rows from the caller, int32 rows of a device population)
//   * request_score_incremental (base + delta list materialised in shared memory)
//   * the island kernels (base + move descriptor).
//
// N-Queens / TSP: one warp per candidate; distinct counts via a shared-memory bitmap,
// distances gathered from the (L2-resident) matrix, warp-shuffle reductions.
// VRP: one CTA per candidate; stable counting sort of the stops by vehicle in shared
// memory, then one thread walks one route in the reference's own order.
#pragma once

#include "gj_device.cuh"

// ---- candidate sources ----------------------------------------------------------------

struct GjSrcF64 {           // a row of the caller's Vec<Vec<f64>>, decoded on the fly
    const double* row;
    const GjProblemDev* P;
    __device__ __forceinline__ int operator()(int i) const { return gj_decode(*P, i, row[i]); }
};

struct GjSrcI32 {           // a row of an int32 device population, or a smem copy
    const int32_t* row;
    __device__ __forceinline__ int operator()(int i) const { return row[i]; }
};

// ---- N-Queens -----------------------------------------------------------------------
// examples/nqueens/src/score/incremental_score_calculator.rs:44-56 (ISC) and
// plain_score_calculator.rs:37-59 (PSC): (N - |rows|) + (N - |col+row|) + (N - |col-row|).
// bm: rows | desc | asc bitmaps, P.bm_words + P.desc_words + P.asc_words words.
template <class Src>
__device__ __forceinline__ double gj_nqueens_eval_warp(const GjProblemDev& P, const Src& src,
                                                       uint32_t* bm, int lane) {
    const int n = P.n_vars;
    const int total_words = P.bm_words + P.desc_words + P.asc_words;
    for (int w = lane; w < total_words; w += 32) bm[w] = 0u;
    __syncwarp();
    uint32_t* bm_desc = bm + P.bm_words;
    uint32_t* bm_asc = bm_desc + P.desc_words;
    for (int i = lane; i < n; i += 32) {
        int row = src(i);
        int col = P.column_id[i];
        unsigned r = (unsigned)(row - P.val_lo);
        unsigned d = (unsigned)(col + row - P.desc_lo);
        unsigned a = (unsigned)(col - row - P.asc_lo);
        atomicOr(&bm[r >> 5], 1u << (r & 31));
        atomicOr(&bm_desc[d >> 5], 1u << (d & 31));
        atomicOr(&bm_asc[a >> 5], 1u << (a & 31));
    }
    __syncwarp();
    int u_rows = 0, u_desc = 0, u_asc = 0;
    for (int w = lane; w < P.bm_words; w += 32) u_rows += __popc(bm[w]);
    for (int w = lane; w < P.desc_words; w += 32) u_desc += __popc(bm_desc[w]);
    for (int w = lane; w < P.asc_words; w += 32) u_asc += __popc(bm_asc[w]);
    u_rows = gj_warp_sum(u_rows);
    u_desc = gj_warp_sum(u_desc);
    u_asc = gj_warp_sum(u_asc);
    __syncwarp();
    double a = (double)(n - u_rows), b = (double)(n - u_desc), c = (double)(n - u_asc);
    return a + b + c;
}

// ---- TSP -----------------------------------------------------------------------------
// examples/tsp/src/score/incremental_score_calculator.rs:71-80 and
// plain_score_calculator.rs:34-43, 70-84: hard = n_stops - |{location ids}|,
// soft = ((0.0 + D[0][s0]) + D[s_last][0]) + fold_{i>=1}(D[s_{i-1}][s_i]).
//
// Two summation modes (P.exact_sums):
//   exact : the 32 distances a warp gathers per step are folded strictly in stop order
//           (shuffle + DADD chain, every lane redundantly), reproducing the reference's
//           sequential f64 fold bit for bit.  This matters more than it looks: distances
//           are multiples of 1e-3, so exact tour lengths sit ON the truncation boundaries
//           of ScoreTrait::round -- the third decimal of the reference's rounded score is
//           decided by its summation order.
//   fast  : per-lane partial sums + tree reduction (documented tolerance: 1e-12 relative
//           before rounding, one 1e-3 quantum after).
template <class Src>
__device__ __forceinline__ void gj_tsp_eval_warp(const GjProblemDev& P, const Src& src,
                                                 uint32_t* bm, int lane, double& dup,
                                                 double& dist) {
    const int n = P.n_vars;
    const size_t L = (size_t)P.n_locations;
    const double* __restrict__ D = P.D;
    const bool exact = P.exact_sums != 0;
    for (int w = lane; w < P.bm_words; w += 32) bm[w] = 0u;
    __syncwarp();
    double acc0 = 0.0, acc1 = 0.0, fold = 0.0, head = 0.0;
    int carry = 0;      // the depot (location 0) precedes stop 0
    for (int base = 0; base < n; base += 64) {
        const int i0 = base + lane, i1 = base + 32 + lane;
        const int v0 = (i0 < n) ? src(i0) : 0;
        const int v1 = (i1 < n) ? src(i1) : 0;
        int p0 = __shfl_up_sync(GJ_FULL_MASK, v0, 1);
        int p1 = __shfl_up_sync(GJ_FULL_MASK, v1, 1);
        const int last0 = __shfl_sync(GJ_FULL_MASK, v0, 31);
        if (lane == 0) { p0 = carry; p1 = last0; }
        carry = __shfl_sync(GJ_FULL_MASK, v1, 31);
        double d0 = 0.0, d1 = 0.0;
        if (i0 < n) d0 = __ldg(&D[(size_t)p0 * L + (size_t)v0]);
        if (i1 < n) d1 = __ldg(&D[(size_t)p1 * L + (size_t)v1]);
        if (i0 < n) { unsigned b = (unsigned)(v0 - P.val_lo); atomicOr(&bm[b >> 5], 1u << (b & 31)); }
        if (i1 < n) { unsigned b = (unsigned)(v1 - P.val_lo); atomicOr(&bm[b >> 5], 1u << (b & 31)); }
        if (exact) {
            if (base == 0) {                      // D[0][s0] is added outside the fold
                head = __shfl_sync(GJ_FULL_MASK, d0, 0);
                if (lane == 0) d0 = 0.0;
            }
#pragma unroll
            for (int l = 0; l < 32; ++l) fold = fold + __shfl_sync(GJ_FULL_MASK, d0, l);
            if (base + 32 < n) {
#pragma unroll
                for (int l = 0; l < 32; ++l) fold = fold + __shfl_sync(GJ_FULL_MASK, d1, l);
            }
        } else {
            acc0 += d0;
            acc1 += d1;
        }
    }
    // closing edge D[s_last][0]
    double closing = 0.0;
    if (lane == 0) closing = __ldg(&D[(size_t)src(n - 1) * L]);
    closing = __shfl_sync(GJ_FULL_MASK, closing, 0);
    __syncwarp();
    int uniq = 0;
    for (int w = lane; w < P.bm_words; w += 32) uniq += __popc(bm[w]);
    uniq = gj_warp_sum(uniq);
    if (exact) {
        double sample_distance = 0.0;
        sample_distance += head;
        sample_distance += closing;
        sample_distance += fold;
        dist = sample_distance;
    } else {
        dist = gj_warp_sum(acc0 + acc1) + closing;
    }
    dup = (double)(n - uniq);
    __syncwarp();
}

// ---- VRP -----------------------------------------------------------------------------
// examples/vrp/src/score/incremental_score_calculator.rs:58-137 (ISC, file variant)
// examples/vrp_service/src/score/incremental_score_calculator.rs:58-138 (ISC, service)
// examples/vrp/src/score/plain_score_calculator.rs:51-233 (PSC)
enum { GJ_TW_ISC_FILE = 0, GJ_TW_ISC_SERVICE = 1, GJ_TW_PSC = 2 };

// warps of the CTA that evaluates one VRP candidate (k_plain_vrp / k_incr_vrp / k_score_moves_vrp)
#ifndef GJ_VRP_WARPS
#define GJ_VRP_WARPS 8
#endif
static constexpr int kVrpWarps = GJ_VRP_WARPS;
// development aid (-DGJ_VRP_PHASE_CLOCKS): thread 0 of every CTA adds the cycles between phase marks
#ifdef GJ_VRP_PHASE_CLOCKS
static __device__ unsigned long long gj_vrp_phase_cycles[16];
#define GJ_PHASE_DECL long long gj_ph_t0 = clock64(); (void)gj_ph_t0
#define GJ_PHASE_MARK(k) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&gj_vrp_phase_cycles[k], (unsigned long long)(t_ - gj_ph_t0)); gj_ph_t0 = t_; } } while (0)
#else
#define GJ_PHASE_MARK(k) do { } while (0)
#endif



// Shared-memory plan of one VRP candidate (one CTA).  `legs` (models without time windows): the decoded
// columns are dead once the stops are bucketed, so their space -- padded to 8 bytes per stop -- is
// reused for the leg lengths of the bucketed stops, and `rl` collects the per-route demand.
struct GjVrpSmem {
    uint32_t* bm;            // customer-id bitmap, P.bm_words
    int* cnt;                // [n_warps][K] per-warp-block vehicle counts -> offsets
    int* start;              // [K + 1] route starts in `bucket`
    double* vdist;           // [K] per-vehicle distance
    unsigned long long* acc; // [0] capacity penalty, [1] lateness penalty
    uint16_t* veh;           // [n_stops] decoded vehicle ids (relative to veh_lo)
    uint16_t* pos;           // [n_stops] rank of the stop among its warp block's stops of the same vehicle
    int32_t* cust;           // [n_stops] decoded customer ids
    uint16_t* bucket;        // [n_stops] customers grouped by vehicle, stop order kept (location ids < 65536:
                             // checked by gj_problem_create)
    double* wfold;           // [n_warps][32] leg lengths of the chunk a warp is folding
    double* leg;             // legs: [n_stops] D[bucket[i-1]][bucket[i]], over cust / veh
    uint32_t* rl;            // legs: [2][K] per-route demand, low and high 16-bit halves summed apart
    double* dfl;             // legs: [2][K] depot -> first stop, last stop -> depot
};

__host__ __device__ inline size_t gj_vrp_smem_bytes(int n_stops, int K, int bm_words, int n_warps, bool legs) {
    size_t b = 0;
    b += (size_t)bm_words * 4;
    b += (size_t)n_warps * (size_t)K * 4;
    b += (size_t)(K + 1) * 4;
    if (legs) b += (size_t)K * 8;
    b = (b + 7) & ~(size_t)7;
    if (legs) b += (size_t)K * 16;
    b += (size_t)K * 8;
    b += 32;
    b = (b + 15) & ~(size_t)15;
    if (!legs) b += (size_t)n_warps * 32 * 8;
    b += ((size_t)n_stops * 2 + 7) & ~(size_t)7;
    b += legs ? (size_t)n_stops * 8 : (size_t)n_stops * 4 + 2 * (((size_t)n_stops * 2 + 7) & ~(size_t)7);
    return b;
}

__device__ __forceinline__ GjVrpSmem gj_vrp_carve(unsigned char* smem, int n_stops, int K,
                                                  int bm_words, int n_warps, bool legs) {
    GjVrpSmem s;
    size_t o = 0;
    s.bm = (uint32_t*)(smem + o); o += (size_t)bm_words * 4;
    s.cnt = (int*)(smem + o); o += (size_t)n_warps * (size_t)K * 4;
    s.start = (int*)(smem + o); o += (size_t)(K + 1) * 4;
    s.rl = (uint32_t*)(smem + o);
    if (legs) o += (size_t)K * 8;
    o = (o + 7) & ~(size_t)7;
    s.dfl = (double*)(smem + o);
    if (legs) o += (size_t)K * 16;
    s.vdist = (double*)(smem + o); o += (size_t)K * 8;
    s.acc = (unsigned long long*)(smem + o); o += 32;      // capacity, lateness, distinct customers, -
    o = (o + 15) & ~(size_t)15;
    s.wfold = (double*)(smem + o);
    if (!legs) o += (size_t)n_warps * 32 * 8;
    s.bucket = (uint16_t*)(smem + o); o += ((size_t)n_stops * 2 + 7) & ~(size_t)7;
    s.leg = (double*)(smem + o);                     // [n_stops] f64 over cust (4 B / stop) + veh (2 B) + pos (2 B)
    s.cust = (int32_t*)(smem + o); o += (size_t)n_stops * 4;
    s.veh = (uint16_t*)(smem + o);
    s.pos = legs ? s.veh + n_stops : (uint16_t*)(smem + o + (((size_t)n_stops * 2 + 7) & ~(size_t)7));
    return s;
}

// Optional per-route by-products of an evaluation (global memory; the delta evaluator's base state).
struct GjVrpOut {
    int32_t* bstop;              // [n_stops] stop index of every bucket slot
    unsigned long long* rload;   // [K] demand carried per vehicle
    unsigned long long* rlate;   // [K] lateness per vehicle
};

// Clears the counters of an evaluation.  A caller that has a barrier of its own between filling
// s.veh / s.cust and the evaluation calls this before that barrier and passes zeroed = true.
__device__ __forceinline__ void gj_vrp_eval_zero(const GjProblemDev& P, const GjVrpSmem& s) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int K = P.n_vehicles, n_warps = nthr >> 5;
    for (int w = tid; w < P.bm_words; w += nthr) s.bm[w] = 0u;
    for (int w = tid; w < n_warps * K; w += nthr) s.cnt[w] = 0;
    if (!P.time_windowed) for (int w = tid; w < 2 * K; w += nthr) s.rl[w] = 0u;
    if (tid < 3) s.acc[tid] = 0ull;
}

// Evaluates the candidate whose decoded (vehicle, customer) columns already sit in
// s.veh / s.cust.  All threads of the CTA must call it.  Results valid in thread 0.
__device__ __forceinline__ void gj_vrp_eval_cta(const GjProblemDev& P, const GjVrpSmem& s,
                                                int tw_mode, double& dup1000, double& cap,
                                                double& dist, double& late,
                                                const GjVrpOut* out = nullptr, bool zeroed = false) {
    const int n = P.n_entities;
    const int K = P.n_vehicles;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, n_warps = nthr >> 5;
    const size_t L = (size_t)P.n_locations;
    const double* __restrict__ D = P.D;
#ifdef GJ_VRP_PHASE_CLOCKS
    GJ_PHASE_DECL;
#endif

    if (!zeroed) {
        gj_vrp_eval_zero(P, s);
        __syncthreads();
    GJ_PHASE_MARK(3);
    }

    // pass 1: customer bitmap + per-warp-block vehicle histogram.  Warp w owns the contiguous block of
    // stops [w*blk, (w+1)*blk) so that block order == stop order, and walks it 32 stops at a time: lanes
    // with the same vehicle are ranked by lane id (= stop order), the first of them advances the block's
    // counter of that vehicle, and every stop keeps its rank among the block's stops of its vehicle
    // (pos) -- pass 2 then needs neither a match nor a serial chain.  Four chunks' loads (and, without time
    // windows, their demand gathers) are issued before the first ranked chunk.
    const int blk = (n + n_warps - 1) / n_warps;
    const int lo = warp * blk;
    const int hi = min(n, lo + blk);
    int* mycnt = s.cnt + warp * K;
    const bool legs = !P.time_windowed;
    for (int i0 = lo + lane; i0 < hi + lane; i0 += 4 * 32) {
        int c[4], v[4];
        unsigned dem[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32;
            c[u] = 0; v[u] = 0x10000 + lane; dem[u] = 0u;
            if (i < hi) {
                c[u] = s.cust[i]; v[u] = s.veh[i];
                // the demand a route carries does not depend on the stop order: summed here
                if (legs) dem[u] = P.cust[c[u]].x;
            }
        }
        // the four matches are independent of each other: issued back to back, ahead of the counter chain
        unsigned grp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) grp[u] = __match_any_sync(GJ_FULL_MASK, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 - lane + u * 32 >= hi) break;                          // uniform over the warp
            const bool on = i0 + u * 32 < hi;
            const int rank = __popc(grp[u] & ((1u << lane) - 1u));
            int old = 0;
            if (on && rank == 0) { old = mycnt[v[u]]; mycnt[v[u]] = old + __popc(grp[u]); }
            __syncwarp();                                                 // the counters, before the next chunk reads them
            old = __shfl_sync(GJ_FULL_MASK, old, __ffs(grp[u]) - 1);
            if (on) s.pos[i0 + u * 32] = (uint16_t)(old + rank);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u * 32 >= hi) continue;
            const unsigned b = (unsigned)(c[u] - P.val_lo);
            atomicOr(&s.bm[b >> 5], 1u << (b & 31));
        }
        if (legs) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u * 32 >= hi) continue;
                // 16-bit halves apart, so that 32-bit shared atomics cannot overflow
                atomicAdd(&s.rl[v[u]], dem[u] & 0xffffu);
                if (dem[u] >> 16) atomicAdd(&s.rl[K + v[u]], dem[u] >> 16);
            }
        }
    }
    __syncthreads();
    GJ_PHASE_MARK(4);

    // exclusive scan over (vehicle major, warp minor): cnt[w][v] -> first slot of warp w's stops of
    // vehicle v; start[v] = route start.  Warp 0 alone, 32 vehicles per round (column sums, warp scan,
    // offsets written back) -- no CTA barrier inside; meanwhile the last warp counts the distinct
    // customers, which nothing but the final score needs.
    if (warp == 0) {
        int carry = 0;
        if (lane == 0) s.start[0] = 0;
        for (int base = 0; base < K; base += 32) {
            const int v = base + lane;
            int tot = 0;
            if (v < K) for (int w = 0; w < n_warps; ++w) tot += s.cnt[w * K + v];
            int x = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(GJ_FULL_MASK, x, o);
                if (lane >= o) x += y;
            }
            if (v < K) {
                s.start[v + 1] = x + carry;
                int off = x + carry - tot;
                for (int w = 0; w < n_warps; ++w) {
                    const int c = s.cnt[w * K + v];
                    s.cnt[w * K + v] = off;
                    off += c;
                }
            }
            carry += __shfl_sync(GJ_FULL_MASK, x, 31);
        }
    }
    if (warp == n_warps - 1) {
        int uniq = 0;
        for (int w = lane; w < P.bm_words; w += 32) uniq += __popc(s.bm[w]);
        uniq = gj_warp_sum(uniq);
        if (lane == 0) s.acc[2] = (unsigned long long)uniq;
    }
    __syncthreads();
    GJ_PHASE_MARK(8);

    // pass 2: stable scatter: slot = first slot of (warp block, vehicle) + rank inside the block
    for (int i = lo + lane; i < hi; i += 32) {
        const int slot = mycnt[s.veh[i]] + (int)s.pos[i];
        s.bucket[slot] = (uint16_t)s.cust[i];
        if (out) out->bstop[slot] = i;
    }
    __syncthreads();
    GJ_PHASE_MARK(9);

    // route walks: one WARP per vehicle.  The 32 stops of a chunk gather their customer facts and leg
    // lengths in parallel (one L2 round trip per chunk instead of one per stop); demand and lateness are
    // integer warp sums; the arrival-time recurrence arrival' = max(arrival, start) + service is a max-plus
    // map a -> max(a + A, B), composed associatively by a warp scan; only the f64 distance fold stays
    // sequential, in the reference's own order -- so every level is bit-identical to a one-thread-per-route
    // walk (which this replaces: it paid one L2 round trip per stop).
    unsigned long long my_cap = 0ull, my_late = 0ull;
    if (legs) {
        // Without time windows a route is its demand (summed in pass 1) and its distance fold.  Every
        // thread gathers the leg lengths of its share of the bucketed stops -- all 2 000 matrix gathers of
        // a candidate in flight at once -- into the space the decoded columns no longer need; then ONE
        // THREAD per route adds its legs strictly in stop order (a shared-memory load and a DADD per stop).
        // The warp-per-route walk below spends 32 lanes on that serial chain: 3x the instructions.
        // (a route's first stop gets a leg nobody reads.)  leg[] lies over cust / veh, last read before the
        // barrier that closed pass 2.  Eight gathers per thread in flight before the first store.
        {
            // the two depot legs of a route ride along with the first batch of leg gathers (threads < K):
            // the fold below then reads shared memory only
            double d_first = 0.0, d_last = 0.0;
            const bool has_route = tid < K;
            if (has_route) {
                const int b = s.start[tid], e = s.start[tid + 1];
                if (e != b) {
                    const size_t depot = (size_t)P.veh_depot[tid];
                    d_first = __ldg(&D[depot * L + (size_t)s.bucket[b]]);
                    d_last = __ldg(&D[(size_t)s.bucket[e - 1] * L + depot]);
                }
            }
            constexpr int U = 8;
            for (int i0 = tid; i0 < n; i0 += U * nthr) {
                double d[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * nthr;
                    d[u] = 0.0;
                    if (i < n) d[u] = __ldg(&D[(size_t)s.bucket[i > 0 ? i - 1 : 0] * L + (size_t)s.bucket[i]]);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * nthr;
                    if (i < n) s.leg[i] = d[u];
                }
            }
            if (has_route) { s.dfl[tid] = d_first; s.dfl[K + tid] = d_last; }
            for (int v = nthr + tid; v < K; v += nthr) {          // K > CTA width: the remaining routes
                const int b = s.start[v], e = s.start[v + 1];
                double f = 0.0, l = 0.0;
                if (e != b) {
                    const size_t depot = (size_t)P.veh_depot[v];
                    f = __ldg(&D[depot * L + (size_t)s.bucket[b]]);
                    l = __ldg(&D[(size_t)s.bucket[e - 1] * L + depot]);
                }
                s.dfl[v] = f; s.dfl[K + v] = l;
            }
        }
        __syncthreads();
        GJ_PHASE_MARK(10);
        for (int v = tid; v < K; v += nthr) {
            const int b = s.start[v], e = s.start[v + 1];
            double current_distance = 0.0;
            const unsigned long long route_load = (unsigned long long)s.rl[v] + ((unsigned long long)s.rl[K + v] << 16);
            if (e != b) {
                const double d_first = s.dfl[v], d_last = s.dfl[K + v];
                // the loads run ahead of the dependent DADD chain, four legs at a time
                double fold = 0.0;
                int i = b + 1;
                for (; i + 4 <= e; i += 4) {
                    const double x0 = s.leg[i], x1 = s.leg[i + 1], x2 = s.leg[i + 2], x3 = s.leg[i + 3];
                    fold = fold + x0; fold = fold + x1; fold = fold + x2; fold = fold + x3;
                }
                for (; i < e; ++i) fold = fold + s.leg[i];
                current_distance += d_first;
                current_distance += d_last;
                current_distance += fold;
                const unsigned long long capv = P.veh_capacity[v];
                if (route_load > capv) my_cap += route_load - capv;
            }
            s.vdist[v] = current_distance;
            if (out) { out->rload[v] = route_load; out->rlate[v] = 0ull; }
        }
    } else
    for (int v = warp; v < K; v += n_warps) {
        const int b = s.start[v], e = s.start[v + 1];
        const int len = e - b;
        double current_distance = 0.0;
        unsigned long long route_load = 0ull, route_late = 0ull;
        if (len != 0) {
            const uint16_t* st = s.bucket + b;
            const size_t depot = (size_t)P.veh_depot[v];
            const int upto = (tw_mode == GJ_TW_PSC) ? len - 1 : len;      // the PSC walk skips the last stop (Q3)
            unsigned long long arrival = P.day_start[v];
            double fold = 0.0;
            for (int base = 0; base < len; base += 32) {
                const int i = base + lane;
                const bool on = i < len;
                const int c = on ? st[i] : 0;
                uint4 f = make_uint4(0u, 0u, 0u, 0u);
                double d = 0.0;
                if (on) {
                    f = P.cust[c];
                    if (i > 0) d = __ldg(&D[(size_t)st[i - 1] * L + (size_t)c]);
                }
                route_load += (unsigned long long)__reduce_add_sync(GJ_FULL_MASK, f.x & 0xffffu) +
                              ((unsigned long long)__reduce_add_sync(GJ_FULL_MASK, f.x >> 16) << 16);
                if (P.time_windowed) {
                    const bool tw_on = i < upto;
                    const unsigned long long ws = f.y, we = f.z, sv = tw_on ? (unsigned long long)f.w : 0ull;
                    unsigned long long SA = sv, SB = tw_on ? ws + sv : 0ull;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned long long pa = __shfl_up_sync(GJ_FULL_MASK, SA, o);
                        const unsigned long long pb = __shfl_up_sync(GJ_FULL_MASK, SB, o);
                        if (lane >= o) { SB = max(pb + SA, SB); SA = pa + SA; }
                    }
                    const unsigned long long after = max(arrival + SA, SB);          // leaving stop i
                    unsigned long long before = __shfl_up_sync(GJ_FULL_MASK, after, 1);
                    if (lane == 0) before = arrival;
                    const unsigned long long t = max(before, ws);                     // service start at stop i
                    unsigned long long lt = 0ull;
                    if (tw_on) {
                        if (tw_mode == GJ_TW_ISC_FILE) { if (t + sv > we) lt = (t + sv) - we; }
                        else { if (t > we + sv) lt = t - (we + sv); }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) lt += __shfl_xor_sync(GJ_FULL_MASK, lt, o);
                    route_late += lt;
                    arrival = __shfl_sync(GJ_FULL_MASK, after, 31);
                }
                // fold_{i >= 1} D[s_{i-1}][s_i], strictly in stop order.  The chunk's 32 leg lengths go through
                // shared memory and are added by a chain of broadcast loads (two values per LDS.128, one DADD
                // per stop) -- a shuffle per stop cost ~12 instructions and was 60 % of the kernel's
                // instruction count.  Lane 0 of the first chunk and the lanes past the route's end hold 0.0:
                // x + 0.0 == x, so the full 32-step chain yields the same bits as stopping at the last stop.
                double* wf = s.wfold + warp * 32;
                wf[lane] = d;
                __syncwarp();
                const double2* wf2 = reinterpret_cast<const double2*>(wf);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const double2 x = wf2[k];
                    fold = fold + x.x;
                    fold = fold + x.y;
                }
                __syncwarp();
            }
            if (lane == 0) {
                current_distance += __ldg(&D[depot * L + (size_t)st[0]]);
                current_distance += __ldg(&D[(size_t)st[len - 1] * L + depot]);
                current_distance += fold;
                const unsigned long long capv = P.veh_capacity[v];
                if (route_load > capv) my_cap += route_load - capv;
                if (P.time_windowed) {
                    if (arrival > P.day_end[v]) route_late += arrival - P.day_end[v];
                    my_late += route_late;
                }
            }
        }
        if (lane == 0) {
            s.vdist[v] = current_distance;
            if (out) { out->rload[v] = route_load; out->rlate[v] = route_late; }
        }
    }
    if (my_cap) atomicAdd(&s.acc[0], my_cap);
    if (my_late) atomicAdd(&s.acc[1], my_late);
    __syncthreads();
    GJ_PHASE_MARK(11);

    if (tid == 0) {
        const int uniq = (int)s.acc[2];
        // vehicle_distances.iter().sum(): sequential, vehicle order (ISC :132)
        double sum_distance = 0.0;
        int v = 0;
        for (; v + 4 <= K; v += 4) {
            const double x0 = s.vdist[v], x1 = s.vdist[v + 1], x2 = s.vdist[v + 2], x3 = s.vdist[v + 3];
            sum_distance += x0; sum_distance += x1; sum_distance += x2; sum_distance += x3;
        }
        for (; v < K; ++v) sum_distance += s.vdist[v];
        dist = sum_distance;
        dup1000 = 1000.0 * (double)(n - uniq);
        cap = (double)s.acc[0];
        late = (double)s.acc[1];
    }
    GJ_PHASE_MARK(12);
}

// ---- weighted combination -------------------------------------------------------------
// score_calculators/plain_score_calculator.rs:79-90 / incremental_score_calculator.rs:84-95:
// sum = null_score; for each constraint: sum += score_i.mul(weight_i).
__device__ __forceinline__ void gj_combine_nqueens(const GjProblemDev& P, double v, double* out) {
    double acc = 0.0;
    acc += P.w[0] * v;
    out[0] = acc;
}

__device__ __forceinline__ void gj_combine_tsp(const GjProblemDev& P, bool isc, double dup,
                                               double dist, double* out) {
    double h = 0.0, f = 0.0;
    if (isc) {
        h += P.w[0] * dup; f += P.w[0] * dist;
    } else {
        h += P.w[0] * dup; f += P.w[0] * 0.0;
        h += P.w[1] * 0.0; f += P.w[1] * dist;
    }
    out[0] = h; out[1] = f;
}

__device__ __forceinline__ void gj_combine_vrp(const GjProblemDev& P, bool isc, double dup1000,
                                               double cap, double dist, double late, double* out) {
    double h = 0.0, m = 0.0, f = 0.0;
    if (isc) {
        const double hard = dup1000 + cap;
        h += P.w[0] * hard; m += P.w[0] * late; f += P.w[0] * dist;
    } else {
        h += P.w[0] * dup1000; m += P.w[0] * 0.0; f += P.w[0] * 0.0;
        h += P.w[1] * cap;     m += P.w[1] * 0.0; f += P.w[1] * 0.0;
        h += P.w[2] * 0.0;     m += P.w[2] * 0.0; f += P.w[2] * dist;
        if (P.time_windowed) { h += P.w[3] * 0.0; m += P.w[3] * late; f += P.w[3] * 0.0; }
    }
    out[0] = h; out[1] = m; out[2] = f;
}
