// gj_islands_vrp_chain.cuh -- LateAcceptance / SimulatedAnnealing chains on the VRP models: many
// steps per launch, one warp per chain, route-level delta evaluation (included by gj_islands.cu).
//
// A single-neighbour agent on VRPTW-5000 (BASELINE config C4) used to pay one 5000-stop full
// evaluation per step (late_acceptance_base.rs:116-141 -> vrp_service ISC :32-143).  A change / swap
// move re-labels a handful of stops, i.e. touches 2-4 of the 125 routes.  Every chain therefore keeps
// a ROUTE INDEX of its current solution in HBM (L2-resident between the steps of a launch):
//     rs   [K][n]     the stops of every route, ascending (= the reference's visiting order: stops are
//                     bucketed by vehicle in stop order, ISC :87-98), each entry (stop << 16) | customer --
//                     both below 65536 (gj_problem_create) -- so that a walk needs no second, dependent
//                     read of the solution row per stop
//     rlen / rdist / rload / rlate [K]   length, distance, demand, lateness of every route
//     cnt  [locations]                   customer occurrence counts;  tot: duplicates, capacity, lateness
// and a step re-walks only the touched routes: the warp loads 32 stops of the old route at a time,
// drops the departures, merges the arrivals (both sorted by stop index), gathers customer facts and
// leg lengths for the merged run in parallel, and folds them in order -- the reference's own
// summation order, so hard, medium AND the float soft level are bit-identical to the full
// evaluation.  The merged stop lists go to a spare buffer and are copied over the route slots only
// when the neighbour is accepted.  Segment moves (insertion / inverse) are not generated on this
// path (gj_islands_create picks it only when move_probas excludes them).
#pragma once


static constexpr int kVrpChainWarps = 4;        // prepare kernel
#ifndef GJ_VRPC_STEP_WARPS
#define GJ_VRPC_STEP_WARPS 28     // 28 x 32 threads leave 72 registers per thread (32 warps: 64, 500 B more spills)
#endif
static constexpr int kVrpStepWarps = GJ_VRPC_STEP_WARPS;   // step kernel: warps of a CTA re-align every step
#define GJ_VRPC_Q 48             // 32 stops of the old route + <= 16 arrivals

struct GjVrpcScratch {
    alignas(16) double qd[32];   // a chunk of leg / route lengths on its way through the sequential fold
    int32_t qc[GJ_VRPC_Q];       // customers
    int32_t qs[GJ_VRPC_Q];       // stop indices
    // the move: changed stops (new / old labels), arrivals of the route being walked
    int cs_stop[GJ_VRP_MAXCS], cs_v[GJ_VRP_MAXCS], cs_c[GJ_VRP_MAXCS], cs_ov[GJ_VRP_MAXCS], cs_oc[GJ_VRP_MAXCS];
    int arr_stop[GJ_VRP_MAXCS], arr_c[GJ_VRP_MAXCS];
    int av[GJ_VRP_MAXAV], noff[GJ_VRP_MAXAV], nlen[GJ_VRP_MAXAV];
    double nd[GJ_VRP_MAXAV];
    unsigned long long nl[GJ_VRP_MAXAV], nt[GJ_VRP_MAXAV];
    int ncs, nav, na, d_uniq;
    GjMove mv;
};

// what a walk without a move needs (rebuilds): the merged-run buffers only
struct GjVrpcScratchLite {
    alignas(16) double qd[32];
    int32_t qc[GJ_VRPC_Q];
    int32_t qs[GJ_VRPC_Q];
    int cs_stop[1], cs_v[1], cs_c[1], arr_stop[1], arr_c[1];
};

struct GjRouteStat { double dist; unsigned long long load, late; int len; };

// Walks route v of the chain: old stop list `rs_v[len]` minus the changed stops that leave, plus the
// arrivals q.arr_* (sorted by stop), changed stops that stay take their new customer.  Every lane
// returns the same result.  `out` (nullable) receives the merged stop list.
template <class Scratch>
__device__ __forceinline__ GjRouteStat gj_vrpc_walk(const GjProblemDev& P, int tw_mode, int v, const int32_t* row,
                                                    const int32_t* rs_v, int len, int ncs, int na,
                                                    int32_t* out, Scratch& q, int lane) {
    const size_t L = (size_t)P.n_locations;
    const double* __restrict__ D = P.D;
    int first = -1, last = -1, outn = 0, ia = 0;
    double fold = 0.0;
    unsigned long long load = 0ull, lateness = 0ull, arrival = P.day_start[v];
    const int nchunks = max(1, (len + 31) >> 5);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int idx = (ch << 5) + lane;
        bool keep = idx < len;
        const int e = keep ? rs_v[idx] : 0;
        const int s = keep ? (int)((unsigned)e >> 16) : 0x7fffffff;
        int c = e & 0xffff;
        for (int j = 0; j < ncs; ++j)
            if (q.cs_stop[j] == s) {
                if (q.cs_v[j] != v) keep = false; else c = q.cs_c[j];
            }
        // arrivals that belong before the next chunk's first stop
        const int limit = (ch == nchunks - 1) ? 0x7fffffff : __shfl_sync(GJ_FULL_MASK, s, 31);
        int ib = ia;
        while (ib < na && q.arr_stop[ib] < limit) ++ib;
        const unsigned keepmask = __ballot_sync(GJ_FULL_MASK, keep);
        int pos = __popc(keepmask & ((1u << lane) - 1u));
        for (int a = ia; a < ib; ++a) {
            const int st = q.arr_stop[a];
            if (st < s) ++pos;
            const unsigned before = __ballot_sync(GJ_FULL_MASK, keep && s < st);
            if (lane == 0) {
                const int p = __popc(before) + (a - ia);
                q.qc[p] = q.arr_c[a]; q.qs[p] = st;
            }
        }
        if (keep) { q.qc[pos] = c; q.qs[pos] = s; }
        const int m = __popc(keepmask) + (ib - ia);
        ia = ib;
        __syncwarp();
        // the merged run, 32 stops at a time: facts and legs gathered in parallel; demand and lateness
        // are integer sums (order-free), the arrival-time recurrence arrival' = max(arrival, start) +
        // service is a max-plus map a -> max(a + A, B) and composes associatively (warp scan); only the
        // float distance fold stays sequential, in the reference's order
        for (int b2 = 0; b2 < m; b2 += 32) {
            const int i = b2 + lane;
            const bool on = i < m;
            const int ci = on ? q.qc[i] : 0;
            const int prev = i > 0 ? q.qc[on ? i - 1 : 0] : last;
            uint4 f = make_uint4(0u, 0u, 0u, 0u);
            double d = 0.0;
            if (on) {
                f = P.cust[ci];
                if (prev >= 0) d = __ldg(&D[(size_t)prev * L + (size_t)ci]);
                if (out) out[outn + i] = (q.qs[i] << 16) | ci;
            }
            load += (unsigned long long)__reduce_add_sync(GJ_FULL_MASK, f.x & 0xffffu) +
                    ((unsigned long long)__reduce_add_sync(GJ_FULL_MASK, f.x >> 16) << 16);
            if (P.time_windowed) {
                const unsigned long long ws = f.y, we = f.z, sv = f.w;
                unsigned long long SA = sv, SB = on ? ws + sv : 0ull;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned long long pa = __shfl_up_sync(GJ_FULL_MASK, SA, o);
                    const unsigned long long pb = __shfl_up_sync(GJ_FULL_MASK, SB, o);
                    if (lane >= o) { SB = max(pb + SA, SB); SA = pa + SA; }
                }
                const unsigned long long after = max(arrival + SA, SB);         // leaving stop i
                unsigned long long before = __shfl_up_sync(GJ_FULL_MASK, after, 1);
                if (lane == 0) before = arrival;
                const unsigned long long t = max(before, ws);                    // service start at stop i
                unsigned long long lt = 0ull;
                if (on) {
                    if (tw_mode == GJ_TW_ISC_FILE) { if (t + sv > we) lt = (t + sv) - we; }
                    else { if (t > we + sv) lt = t - (we + sv); }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) lt += __shfl_xor_sync(GJ_FULL_MASK, lt, o);
                lateness += lt;
                arrival = __shfl_sync(GJ_FULL_MASK, after, 31);
            }
            // the chunk's leg lengths are added strictly in stop order: through shared memory, two per
            // broadcast LDS.128 and one DADD each (a shuffle per stop cost twice the instructions).  Lanes
            // past the run's end hold 0.0, and x + 0.0 == x.
            const int cntk = min(32, m - b2);
            q.qd[lane] = d;
            __syncwarp();
            const double2* qd2 = reinterpret_cast<const double2*>(q.qd);
            for (int k = 0; k < cntk; k += 2) {
                const double2 x = qd2[k >> 1];
                fold = fold + x.x;
                fold = fold + x.y;
            }
            __syncwarp();
        }
        if (m > 0) {
            if (first < 0) first = q.qc[0];
            last = q.qc[m - 1];
        }
        outn += m;
        __syncwarp();
    }
    GjRouteStat r;
    double current_distance = 0.0;
    if (first >= 0) {
        const size_t depot = (size_t)P.veh_depot[v];
        current_distance += __ldg(&D[depot * L + (size_t)first]);
        current_distance += __ldg(&D[(size_t)last * L + depot]);
        current_distance += fold;
        if (P.time_windowed && arrival > P.day_end[v]) lateness += arrival - P.day_end[v];
    }
    r.dist = current_distance; r.load = load; r.late = lateness; r.len = outn;
    return r;
}

// sum over the K routes in the reference's order (vehicle_distances.iter().sum()), touched routes
// substituted by their re-walked value
__device__ __forceinline__ double gj_vrpc_sum_routes(const double* rdist, int K, GjVrpcScratch& q, int nav, int lane) {
    double sum = 0.0;
    const double2* qd2 = reinterpret_cast<const double2*>(q.qd);
    for (int v0 = 0; v0 < K; v0 += 32) {
        const int v = v0 + lane;
        double x = v < K ? rdist[v] : 0.0;
        for (int a = 0; a < nav; ++a) if (q.av[a] == v) x = q.nd[a];
        q.qd[lane] = x;                                   // 0.0 past the last route: x + 0.0 == x
        __syncwarp();
        const int m = min(32, K - v0);
        for (int i = 0; i < m; i += 2) {
            const double2 y = qd2[i >> 1];
            sum += y.x;
            sum += y.y;
        }
        __syncwarp();
    }
    return sum;
}

#define GJ_VRPC_KSM 512           // route-length counters kept in shared memory while bucketing

// Bucket phase of a rebuild, one warp: stop lists rs[K][n] (stop order kept: rank among the
// same-vehicle lanes of every 32-stop chunk), route lengths, customer counts.  `sh_rlen`: K <= 
// GJ_VRPC_KSM counters in shared memory (nullptr: the counters in HBM are used directly).
__device__ __forceinline__ void gj_vrpc_bucket(const GjProblemDev& P, const int32_t* row, const GjVrpChainState& V,
                                               int island, int* sh_rlen, int lane) {
    const int n = P.n_entities, K = P.n_vehicles;
    int32_t* rs = V.rs + (size_t)island * K * n;
    int32_t* rlen_g = V.rlen + (size_t)island * K;
    int32_t* cnt = V.cnt + (size_t)island * V.cnt_stride;
    int* rlen = sh_rlen ? sh_rlen : rlen_g;
    for (int i = lane; i < V.cnt_stride; i += 32) cnt[i] = 0;
    for (int v = lane; v < K; v += 32) rlen[v] = 0;
    __syncwarp();
    // four 32-stop chunks in flight: the (vehicle, customer) loads are issued together, the ordered
    // placement then runs out of registers
    for (int s0 = 0; s0 < n; s0 += 128) {
        int2 vc4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int s = s0 + 32 * u + lane;
            vc4[u] = s < n ? *reinterpret_cast<const int2*>(row + 2 * s) : make_int2(0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int s = s0 + 32 * u + lane;
            if (s0 + 32 * u >= n) break;                      // uniform
            const bool on = s < n;
            const int2 vc = vc4[u];
            const int v = on ? vc.x : -1 - lane;
            const unsigned grp = __match_any_sync(GJ_FULL_MASK, v);
            if (on) {
                const int rank = __popc(grp & ((1u << lane) - 1u));
                rs[(size_t)v * n + rlen[v] + rank] = (s << 16) | vc.y;
                atomicAdd(&cnt[vc.y - P.val_lo], 1);
            }
            __syncwarp();
            if (on && lane == 31 - __clz(grp)) rlen[v] += __popc(grp);
            __syncwarp();
        }
    }
    if (sh_rlen) for (int v = lane; v < K; v += 32) rlen_g[v] = sh_rlen[v];
    __syncwarp();
}

// Rebuilds the route index of a chain from its solution row (creation, migrant), one warp.
__device__ __forceinline__ void gj_vrpc_rebuild(const GjProblemDev& P, int tw_mode, const int32_t* row,
                                                const GjVrpChainState& V, int island, GjVrpcScratchLite& q,
                                                int* sh_rlen, int lane) {
    const int n = P.n_entities, K = P.n_vehicles;
    const int32_t* rs = V.rs + (size_t)island * K * n;
    const int32_t* rlen = V.rlen + (size_t)island * K;
    const int32_t* cnt = V.cnt + (size_t)island * V.cnt_stride;
    gj_vrpc_bucket(P, row, V, island, K <= GJ_VRPC_KSM ? sh_rlen : nullptr, lane);
    unsigned long long cap_pen = 0ull, late_pen = 0ull;
    for (int v = 0; v < K; ++v) {
        const GjRouteStat r = gj_vrpc_walk(P, tw_mode, v, row, rs + (size_t)v * n, rlen[v], 0, 0, nullptr, q, lane);
        if (lane == 0) {
            V.rdist[(size_t)island * K + v] = r.dist;
            V.rload[(size_t)island * K + v] = r.load;
            V.rlate[(size_t)island * K + v] = r.late;
        }
        const unsigned long long capv = P.veh_capacity[v];
        if (r.load > capv) cap_pen += r.load - capv;
        late_pen += r.late;
    }
    int distinct = 0;
    for (int i = lane; i < V.cnt_stride; i += 32) distinct += cnt[i] > 0 ? 1 : 0;
    distinct = gj_warp_sum(distinct);
    distinct = __shfl_sync(GJ_FULL_MASK, distinct, 0);
    if (lane == 0) {
        unsigned long long* tot = V.tot + (size_t)island * 4;
        tot[0] = (unsigned long long)(n - distinct);
        tot[1] = cap_pen;
        tot[2] = late_pen;
        V.stale[island] = 0;
    }
    __syncwarp();
}

// Route index of the published global top (slot I), once per published version.  One CTA: warp 0
// buckets the stops, then the warps share the K route walks and the flattening of the stop lists.
static constexpr int kGindexWarps = 32;
__global__ void __launch_bounds__(kGindexWarps * 32)
k_vrp_chain_gindex(GjProblemDev P, int I, const int32_t* gbest, const int* gver, GjVrpChainState V) {
    __shared__ GjVrpcScratchLite sh_q[kGindexWarps];
    __shared__ int sh_rlen[GJ_VRPC_KSM];
    __shared__ unsigned long long sh_pen[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ver = *gver;
    if (ver == *V.gidx_ver) return;                   // uniform over the CTA
    const int n = P.n_entities, K = P.n_vehicles;
    const int tw_mode = gj_vrp_tw_mode(P);
    const int32_t* rs = V.rs + (size_t)I * K * n;
    int32_t* rlen = V.rlen + (size_t)I * K;
    if (threadIdx.x < 2) sh_pen[threadIdx.x] = 0ull;
    if (warp == 0) gj_vrpc_bucket(P, gbest, V, I, K <= GJ_VRPC_KSM ? sh_rlen : nullptr, lane);
    __syncthreads();
    unsigned long long cap_pen = 0ull, late_pen = 0ull;
    for (int v = warp; v < K; v += kGindexWarps) {
        const GjRouteStat r = gj_vrpc_walk(P, tw_mode, v, gbest, rs + (size_t)v * n, rlen[v], 0, 0, nullptr, sh_q[warp], lane);
        if (lane == 0) {
            V.rdist[(size_t)I * K + v] = r.dist;
            V.rload[(size_t)I * K + v] = r.load;
            V.rlate[(size_t)I * K + v] = r.late;
        }
        const unsigned long long capv = P.veh_capacity[v];
        if (r.load > capv) cap_pen += r.load - capv;
        late_pen += r.late;
    }
    if (lane == 0) { atomicAdd(&sh_pen[0], cap_pen); atomicAdd(&sh_pen[1], late_pen); }
    // flattened stop lists: offsets = exclusive prefix of the route lengths (warp 0, into goff)
    if (warp == 0) {
        int run = 0;
        for (int v0 = 0; v0 < K; v0 += 32) {
            const int v = v0 + lane;
            const int len = v < K ? rlen[v] : 0;
            int inc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(GJ_FULL_MASK, inc, o);
                if (lane >= o) inc += t;
            }
            if (v < K) V.goff[v] = run + inc - len;
            run += __shfl_sync(GJ_FULL_MASK, inc, 31);
        }
    }
    __syncthreads();
    for (int v = warp; v < K; v += kGindexWarps) {
        const int len = rlen[v], p = V.goff[v];
        for (int i = lane; i < len; i += 32) { V.gstop[p + i] = rs[(size_t)v * n + i]; V.gdst[p + i] = v * n + i; }
    }
    if (warp == 0) {
        const int32_t* cnt = V.cnt + (size_t)I * V.cnt_stride;
        int distinct = 0;
        for (int i = lane; i < V.cnt_stride; i += 32) distinct += cnt[i] > 0 ? 1 : 0;
        distinct = gj_warp_sum(distinct);
        distinct = __shfl_sync(GJ_FULL_MASK, distinct, 0);
        if (lane == 0) {
            unsigned long long* tot = V.tot + (size_t)I * 4;
            tot[0] = (unsigned long long)(n - distinct);
            tot[1] = sh_pen[0];
            tot[2] = sh_pen[1];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *V.gidx_ver = ver;
}

// The same route index by ONE full evaluation of the CTA evaluator (gj_vrp_eval_cta: all warps bucket the
// stops, one warp per route) -- its by-products are exactly the index: the stop of every bucket slot is
// the flattened stop list, start[] the offsets, vdist / rload / rlate the per-route statistics (the same
// bits as the walk's: both fold in the reference's order).  Used whenever the evaluator's shared-memory
// plan fits; k_vrp_chain_gindex above (one warp buckets 32 stops at a time) took 5x as long on C4.
// slot: the chain (or I = the global top) whose index is built from `row`; flat: also the flattened lists
__device__ __forceinline__ void gj_vrpc_index_cta(const GjProblemDev& P, unsigned char* smem_raw, const int32_t* __restrict__ row,
                                                  int slot, const GjVrpChainState& V, int32_t* bstop, bool flat) {
    const int n = P.n_entities, K = P.n_vehicles;
    GjVrpSmem s = gj_vrp_carve(smem_raw, n, K, P.bm_words, kVrpWarps, !P.time_windowed);
    int32_t* cnt = V.cnt + (size_t)slot * V.cnt_stride;
    for (int i = threadIdx.x; i < V.cnt_stride; i += blockDim.x) cnt[i] = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int2 pr = *reinterpret_cast<const int2*>(row + 2 * i);
        s.veh[i] = (uint16_t)pr.x;
        s.cust[i] = pr.y;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&cnt[s.cust[i] - P.val_lo], 1);
    GjVrpOut out{bstop, V.rload + (size_t)slot * K, V.rlate + (size_t)slot * K};
    double dup1000 = 0, cap = 0, dist = 0, late = 0;
    gj_vrp_eval_cta(P, s, gj_vrp_tw_mode(P), dup1000, cap, dist, late, &out);
    __syncthreads();
    int32_t* rs = V.rs + (size_t)slot * K * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int v = warp; v < K; v += n_warps) {
        const int b = s.start[v], len = s.start[v + 1] - b;
        for (int i = lane; i < len; i += 32) {
            const int e = (bstop[b + i] << 16) | (int)s.bucket[b + i];
            rs[(size_t)v * n + i] = e;
            if (flat) { bstop[b + i] = e; V.gdst[b + i] = v * n + i; }     // bstop == V.gstop: the flattened lists
        }
        if (lane == 0) {
            V.rlen[(size_t)slot * K + v] = len;
            V.rdist[(size_t)slot * K + v] = s.vdist[v];
            if (flat) V.goff[v] = b;
        }
    }
    if (threadIdx.x == 0) {
        unsigned long long* tot = V.tot + (size_t)slot * 4;
        tot[0] = (unsigned long long)llrint(dup1000 / 1000.0);
        tot[1] = s.acc[0];
        tot[2] = s.acc[1];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kVrpWarps * 32)
k_vrp_chain_gindex_cta(GjProblemDev P, int I, const int32_t* __restrict__ gbest, const int* gver, GjVrpChainState V) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ver = *gver;
    if (ver == *V.gidx_ver) return;                   // uniform over the CTA
    gj_vrpc_index_cta(P, smem_raw, gbest, I, V, V.gstop, true);
    if (threadIdx.x == 0) *V.gidx_ver = ver;
}

// Route index of the chains k_vrp_chain_prepare flagged (creation, an accepted migrant, an adopted global
// top whose index was not ready): one CTA per chain, the others leave at once.  The merged-route spare list
// of the chain, idle between launches, takes the evaluator's stop-per-slot by-product.
__global__ void __launch_bounds__(kVrpWarps * 32)
k_vrp_chain_rebuild_cta(GjProblemDev P, int stride, const int32_t* __restrict__ cur, GjVrpChainState V) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int island = blockIdx.x;
    if (!V.stale[island]) return;                     // uniform over the CTA
    gj_vrpc_index_cta(P, smem_raw, cur + (size_t)island * stride, island, V, V.spare + (size_t)island * P.n_entities, false);
    if (threadIdx.x == 0) V.stale[island] = 0;
}

// Between launches (cold path): update_global_top adopt half (agent_base.rs:465-489), route index of
// chains whose solution was replaced (adopted global top: copy of slot I; migrant / creation:
// rebuild), update_top_individual for the replaced solution.  One warp per chain; chains with
// nothing pending leave after three loads.
__global__ void __launch_bounds__(kVrpChainWarps * 32)
k_vrp_chain_prepare(GjProblemDev P, GjChainArgs A, GjVrpChainState V, int cta_rebuild) {
    __shared__ GjVrpcScratchLite sh_q[kVrpChainWarps];
    __shared__ int sh_rlen[kVrpChainWarps][GJ_VRPC_KSM];
    constexpr int LV = 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int island = blockIdx.x * kVrpChainWarps + warp;
    if (island >= A.I) return;
    GjVrpcScratchLite& q = sh_q[warp];
    const int n = P.n_entities, K = P.n_vehicles;
    int32_t* row = A.cur + (size_t)island * A.stride;
    int32_t* best_row = A.best + (size_t)island * A.stride;
    const bool is_la = A.agent == GJ_AGENT_LATE_ACCEPTANCE;
    // update_global_top, adopt half: the test the reference makes at the end of every iteration
    // (agent_base.rs:465-489) -- here for the last iteration of the previous launch; the step kernel
    // makes the later ones itself ("holding").  gseen[island] == version <=> the stored solution IS
    // that version's gbest row; dirty == 1: a migrant replaced it since.
    int adopted = 0;
    const int stale = V.stale[island], dirty = A.dirty[island];
    if (lane == 0 && A.gver) {
        const int ver = *A.gver;
        const bool cur_is_g = ver != 0 && A.gseen[island] == ver && dirty != 1;
        if (ver != 0 && !cur_is_g) {
            const GjScore g = gj_load_score(A.gbest_score, LV);
            const GjScore mytop = gj_load_score(A.best_score + (size_t)island * GJ_MAX_LEVELS, LV);
            adopted = gj_score_le(mytop, g, LV) ? 0 : 1;                 // global < agent_top
            if (adopted && *V.gidx_ver == ver) adopted = 2;              // ... and its route index is ready
            A.gseen[island] = adopted ? ver : 0;
        }
    }
    adopted = __shfl_sync(GJ_FULL_MASK, adopted, 0);
    if (!adopted && !stale && !dirty) return;
    GjScore cur = gj_load_score(A.cur_score + (size_t)island * GJ_MAX_LEVELS, LV);
    if (adopted) {
        {   // stride is a multiple of 4 ints, rows are 16-byte aligned
            const int4* g4 = reinterpret_cast<const int4*>(A.gbest);
            int4* r4 = reinterpret_cast<int4*>(row);
            for (int i0 = lane; i0 < A.stride / 4; i0 += 128) {
                int4 x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int i = i0 + 32 * u; x[u] = i < A.stride / 4 ? g4[i] : make_int4(0, 0, 0, 0); }
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int i = i0 + 32 * u; if (i < A.stride / 4) r4[i] = x[u]; }
            }
        }
        if (is_la && lane == 0) {   // LateAcceptance remembers the score it leaves behind (agent_base.rs:467-471)
            double* late_g = A.late + (size_t)island * A.late_size * GJ_MAX_LEVELS;
            const int head = (A.late_head[island] + A.late_size - 1) % A.late_size;
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) late_g[(size_t)head * GJ_MAX_LEVELS + l] = cur.v[l];
            A.late_head[island] = head;
            A.late_len[island] = min(A.late_len[island] + 1, A.late_size);
        }
        cur = gj_load_score(A.gbest_score, LV);
        if (lane == 0)
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) A.cur_score[(size_t)island * GJ_MAX_LEVELS + l] = cur.v[l];
        __syncwarp();
    }
    if (adopted == 2) {
        // copy the global top's route index (slot I) instead of re-walking all K routes
        const size_t gI = (size_t)A.I;
        int32_t* rs = V.rs + (size_t)island * K * n;
        // (four independent loads in flight per lane: the copy is latency-, not bandwidth-bound)
        for (int p0 = lane; p0 < n; p0 += 128) {
            int d[4], x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = p0 + 32 * u;
                d[u] = p < n ? V.gdst[p] : -1;
                x[u] = p < n ? V.gstop[p] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) if (d[u] >= 0) rs[d[u]] = x[u];
        }
        for (int v = lane; v < K; v += 32) {
            V.rlen[(size_t)island * K + v] = V.rlen[gI * K + v]; V.rdist[(size_t)island * K + v] = V.rdist[gI * K + v];
            V.rload[(size_t)island * K + v] = V.rload[gI * K + v]; V.rlate[(size_t)island * K + v] = V.rlate[gI * K + v];
        }
        {
            // cnt_stride is a multiple of 32 ints: 16-byte vectors
            const int4* csrc = reinterpret_cast<const int4*>(V.cnt + gI * V.cnt_stride);
            int4* cdst = reinterpret_cast<int4*>(V.cnt + (size_t)island * V.cnt_stride);
            const int c4 = V.cnt_stride / 4;
            for (int i0 = lane; i0 < c4; i0 += 128) {
                int4 x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int i = i0 + 32 * u; x[u] = i < c4 ? csrc[i] : make_int4(0, 0, 0, 0); }
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int i = i0 + 32 * u; if (i < c4) cdst[i] = x[u]; }
            }
        }
        if (lane < 3) V.tot[(size_t)island * 4 + lane] = V.tot[gI * 4 + lane];
        if (lane == 0) V.stale[island] = 0;
    } else if (stale || adopted) {
        // (one warp re-walking 125 routes took as long as everything else in this kernel together: the
        // flagged chains go to k_vrp_chain_rebuild_cta, one CTA each, when the evaluator's plan fits)
        if (cta_rebuild) { if (lane == 0) V.stale[island] = 1; }
        else gj_vrpc_rebuild(P, gj_vrp_tw_mode(P), row, V, island, q, sh_rlen[warp], lane);
    }
    // A solution that arrived between steps meets the agent's top only after the next step
    // (update_top_individual runs once per iteration, agent_base.rs:149-152): the step kernel owes
    // that comparison.  row and best_row are unrelated vectors from here on.
    if ((adopted || dirty) && lane == 0) {
        V.pend[island] = 1;
        V.ndiff[island] = -1;
        A.dirty[island] = 0;
    }
}

#ifndef GJ_VRPC_MINBLOCKS
#define GJ_VRPC_MINBLOCKS 8
#endif
#ifndef GJ_VRPC_SYNC_EVERY
#define GJ_VRPC_SYNC_EVERY 1    // development knob: re-align only every n-th step
#endif
#ifndef GJ_VRPC_SYNC
#define GJ_VRPC_SYNC 1          // bit 0 = barrier at the top of a step, bit 1 = before the totals (measured on
                                // C4: 1 -> 49.7 us per step, 3 -> 51.6, none -> 54; a barrier only every 2nd / 4th /
                                // 8th step -> 55 / 61 / 63; one more before every route walk -> 53.3)
#endif
#if GJ_VRPC_SYNC & 1
#define GJ_VRPC_REALIGN_TOP() __syncthreads()
#else
#define GJ_VRPC_REALIGN_TOP() do { } while (0)
#endif
#if GJ_VRPC_SYNC & 2
#define GJ_VRPC_REALIGN_MID() __syncthreads()
#else
#define GJ_VRPC_REALIGN_MID() do { } while (0)
#endif
template <int AGENT>            // GJ_AGENT_LATE_ACCEPTANCE / GJ_AGENT_SIMULATED_ANNEALING: one rule per instantiation
__global__ void __launch_bounds__(kVrpStepWarps * 32, 1)
k_vrp_chains(GjProblemDev P, GjGroups G, GjChainArgs A, GjVrpChainState V) {
    extern __shared__ __align__(16) unsigned char vrpc_smem[];
    GjVrpcScratch* sh_q = reinterpret_cast<GjVrpcScratch*>(vrpc_smem);
    constexpr int LV = 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int island = blockIdx.x * (blockDim.x >> 5) + warp;    // 4..kVrpStepWarps warps per CTA (host picks)
    if (island >= A.I) return;                       // whole warps only: the barrier below counts live warps
    GjVrpcScratch& q = sh_q[warp];
    const int n = P.n_entities, K = P.n_vehicles;
    const int tw_mode = gj_vrp_tw_mode(P);
    int32_t* row = A.cur + (size_t)island * A.stride;
    int32_t* best_row = A.best + (size_t)island * A.stride;
    int32_t* rs = V.rs + (size_t)island * K * n;
    int32_t* rlen = V.rlen + (size_t)island * K;
    double* rdist = V.rdist + (size_t)island * K;
    unsigned long long* rload = V.rload + (size_t)island * K;
    unsigned long long* rlate = V.rlate + (size_t)island * K;
    unsigned long long* tot = V.tot + (size_t)island * 4;
    int32_t* cnt = V.cnt + (size_t)island * V.cnt_stride;
    int32_t* spare = V.spare + (size_t)island * n;
    int32_t* diff = V.diff + (size_t)island * GJ_VRPC_DIFF;
    uint32_t* tabu_g = A.ctabu ? A.ctabu + (size_t)island * A.ctabu_words_per_island : nullptr;
    double* late_g = A.late ? A.late + (size_t)island * A.late_size * GJ_MAX_LEVELS : nullptr;
    constexpr bool is_la = AGENT == GJ_AGENT_LATE_ACCEPTANCE;
    int late_head = is_la ? A.late_head[island] : 0, late_len = is_la ? A.late_len[island] : 0;
    double temp[GJ_MAX_LEVELS] = {1.0, 1.0, 1.0};
    if (!is_la)
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) temp[l] = A.sa_temp[(size_t)island * GJ_MAX_LEVELS + l];
    GjScore cur = gj_load_score(A.cur_score + (size_t)island * GJ_MAX_LEVELS, LV);
    GjScore top = gj_load_score(A.best_score + (size_t)island * GJ_MAX_LEVELS, LV);
    // best_row (the agent's top) trails the chain: `diff` lists the stops whose labels changed since
    // best_row last equalled the solution row (ndiff < 0: too many -> whole-row copy)
    int ndiff = V.ndiff[island];
    auto materialize_top = [&]() {
        if (ndiff < 0) {
            // whole-row copy (once per adopted / migrated solution): 16-byte vectors, four loads in flight --
            // the other chains of the CTA wait for this one at the next step's barrier
            const int4* r4 = reinterpret_cast<const int4*>(row);
            int4* b4 = reinterpret_cast<int4*>(best_row);
            const int n4 = A.stride / 4;
            for (int i0 = lane; i0 < n4; i0 += 128) {
                int4 x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int i = i0 + 32 * u; x[u] = i < n4 ? r4[i] : make_int4(0, 0, 0, 0); }
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int i = i0 + 32 * u; if (i < n4) b4[i] = x[u]; }
            }
        } else {
            for (int i = lane; i < ndiff; i += 32) {
                const int st = diff[i];
                best_row[2 * st] = row[2 * st];
                best_row[2 * st + 1] = row[2 * st + 1];
            }
        }
        ndiff = 0;
        __syncwarp();
    };
    int accepted_total = 0;
    // update_top_individual copies are lazy: while the top IS the current solution, best_row is only
    // written when the chain is about to leave it (or at the end of the launch)
    bool top_pending = false;
    // update_global_top inside a launch (see k_vrp_chain_prepare): G = the global top published before
    // the launch; while the chain stands on it and it still beats the chain's own top, a worse
    // neighbour that gets accepted "bounces" (the reference moves and is reset within the same iteration)
    const int g_ver = A.gver ? *A.gver : 0;
    const bool have_g = g_ver != 0;
    bool cur_is_g = have_g && A.gseen[island] == g_ver;
    GjScore gsc = cur;
    if (have_g) gsc = gj_load_score(A.gbest_score, LV);
    bool check_pending = V.pend[island] != 0 || (cur_is_g && !gj_score_le(top, gsc, LV));

    for (int it = 0; it < A.n_steps; ++it) {
        // the warps of a CTA run the same code on different chains; re-aligning them every step keeps
        // them in the same instruction-cache lines (instruction fetch was the top stall without it)
#if GJ_VRPC_SYNC_EVERY > 1
        if ((it % GJ_VRPC_SYNC_EVERY) == 0) { GJ_VRPC_REALIGN_TOP(); }
#else
        GJ_VRPC_REALIGN_TOP();
#endif
        const uint64_t step = A.step0 + (uint64_t)it;
        // the chain's totals are needed after the route walks: requested now, the round trip hides under them
        const unsigned long long tot0 = tot[0], tot1 = tot[1], tot2 = tot[2];
        // ---- generate (every lane computes the same move) -----------------------------------------------
        GjMoverParams M = A.M;
        const GjMove m = gj_generate_move<true>(P, G, M, A.seed, (uint32_t)(A.island_base + island), step, 0u,
                                                tabu_g, A.ctabu_off);
        const bool identity = m.kind == GJ_MOVE_NULL || (A.noop != 0 && (m.kind == 3 || (m.kind == 2 && m.k == 2)));
        // ---- changed stops, touched routes (lane 0) --------------------------------------------------
        if (lane == 0) {
            int ncs = 0, nav = 0, d_uniq = 0;
            if (!identity) {
                q.mv = m;
                const GjMove ms = q.mv;
                const int32_t* g = G.ids + G.offsets[ms.group];
                int cols[GJ_MOVE_MAXPAIRS], vals[GJ_MOVE_MAXPAIRS];
                const int np = gj_small_move_pairs(ms, g, true, A.noop != 0, [&](int id) { return row[id]; }, cols, vals);
                for (int i = 0; i < np; ++i) {
                    const int col = cols[i], val = gj_fix_column(P, col, vals[i]);
                    const int stop = col >> 1;
                    int j = 0;
                    for (; j < ncs; ++j) if (q.cs_stop[j] == stop) break;
                    if (j == ncs) {
                        q.cs_stop[j] = stop;
                        q.cs_v[j] = q.cs_ov[j] = row[2 * stop];
                        q.cs_c[j] = q.cs_oc[j] = row[2 * stop + 1];
                        ++ncs;
                    }
                    if (col & 1) q.cs_c[j] = val; else q.cs_v[j] = val;
                }
                // customer multiset -> duplicates
                for (int j = 0; j < ncs; ++j) {
                    if (q.cs_oc[j] == q.cs_c[j]) continue;
                    for (int side = 0; side < 2; ++side) {
                        const int key = (side ? q.cs_c[j] : q.cs_oc[j]) - P.val_lo;
                        bool seen = false;
                        int net = 0;
                        for (int t = 0; t < ncs; ++t) {
                            if (q.cs_oc[t] == q.cs_c[t]) continue;
                            for (int ts = 0; ts < 2; ++ts) {
                                const int other = (ts ? q.cs_c[t] : q.cs_oc[t]) - P.val_lo;
                                if (other != key) continue;
                                if (t < j || (t == j && ts < side)) seen = true;
                                net += ts ? 1 : -1;
                            }
                        }
                        if (seen || net == 0) continue;
                        const int before = cnt[key];
                        d_uniq += ((before + net) > 0 ? 1 : 0) - (before > 0 ? 1 : 0);
                    }
                }
                // routes that gain, lose or re-label a stop
                for (int j = 0; j < ncs; ++j) {
                    if (q.cs_ov[j] == q.cs_v[j] && q.cs_oc[j] == q.cs_c[j]) continue;
                    for (int side = 0; side < 2; ++side) {
                        const int v = side ? q.cs_v[j] : q.cs_ov[j];
                        int a = 0;
                        for (; a < nav; ++a) if (q.av[a] == v) break;
                        if (a == nav) q.av[nav++] = v;
                    }
                }
            }
            q.ncs = ncs; q.nav = nav; q.d_uniq = d_uniq;
        }
        __syncwarp();
        const int ncs = q.ncs, nav = q.nav;
        // ---- re-walk the touched routes -----------------------------------------------------------------
        int off = 0;
        for (int a = 0; a < nav; ++a) {
            const int v = q.av[a];
            if (lane == 0) {
                int na = 0;                                  // arrivals, sorted by stop index
                for (int j = 0; j < ncs; ++j)
                    if (q.cs_v[j] == v && q.cs_ov[j] != v) {
                        int p = na++;
                        while (p > 0 && q.arr_stop[p - 1] > q.cs_stop[j]) {
                            q.arr_stop[p] = q.arr_stop[p - 1]; q.arr_c[p] = q.arr_c[p - 1]; --p;
                        }
                        q.arr_stop[p] = q.cs_stop[j]; q.arr_c[p] = q.cs_c[j];
                    }
                q.na = na;
            }
            __syncwarp();
            const GjRouteStat r = gj_vrpc_walk(P, tw_mode, v, row, rs + (size_t)v * n, rlen[v], ncs, q.na,
                                               spare + off, q, lane);
            if (lane == 0) { q.nd[a] = r.dist; q.nl[a] = r.load; q.nt[a] = r.late; q.noff[a] = off; q.nlen[a] = r.len; }
            off += r.len;
            __syncwarp();
        }
        GJ_VRPC_REALIGN_MID();                          // re-align (see the top of the loop)
        // ---- totals -----------------------------------------------------------------------------------
        unsigned long long cap_pen = tot1, late_pen = tot2;
        for (int a = 0; a < nav; ++a) {
            const int v = q.av[a];
            const unsigned long long capv = P.veh_capacity[v];
            const unsigned long long ol = rload[v], nl = q.nl[a];
            if (ol > capv) cap_pen -= ol - capv;
            if (nl > capv) cap_pen += nl - capv;
            late_pen -= rlate[v];
            late_pen += q.nt[a];
        }
        const double sum_distance = gj_vrpc_sum_routes(rdist, K, q, nav, lane);
        const long long dups = (long long)tot0 - (long long)q.d_uniq;
        GjScore sc;
        gj_combine_vrp(P, true, 1000.0 * (double)dups, (double)cap_pen, sum_distance, (double)late_pen, sc.v);
        gj_score_round(sc, P);                          // agent_base.rs:311-314
        // ---- acceptance ---------------------------------------------------------------------------------
        bool accept;
        if constexpr (is_la) {
            GjScore late_native = cur;                  // late_acceptance_base.rs:196-213
            if (late_len > 0) late_native = gj_load_score(late_g + (size_t)gj_wrap_once(late_head + late_len - 1, A.late_size) * GJ_MAX_LEVELS, LV);
            accept = gj_score_le(sc, late_native, LV) || gj_score_le(sc, cur, LV);
        } else {
            const double u = gj_accept_uniform(A.seed, (uint32_t)(A.island_base + island), step);
            double proba;
            accept = gj_sa_accept(sc, cur, LV, temp, A.sa, u, &proba);      // simulated_annealing_base.rs:198-233
            if (A.trace_aux && lane == 0) {
                double* o = A.trace_aux + (size_t)island * 5;
                o[0] = u; o[1] = proba; o[2] = temp[0]; o[3] = temp[1]; o[4] = temp[2];
            }
        }
        if (A.trace_moves && lane == 0) {
            A.trace_moves[(size_t)it * A.I + island] = m;
            for (int l = 0; l < LV; ++l) A.trace_scores[((size_t)it * A.I + island) * LV + l] = sc.v[l];
            A.trace_accept[(size_t)it * A.I + island] = accept ? 1 : 0;
        }
        const bool holding = cur_is_g && !gj_score_le(top, gsc, LV);
        const bool bounced = accept && holding && !gj_score_le(sc, cur, LV);
        if (bounced) {
            // late_acceptance_base.rs:207-211 pushes the accepted score, agent_base.rs:465-471 pushes
            // population[0].score (the same neighbour) once more and restores the global top
            accepted_total += 1;
            if (is_la) {
                for (int rep = 0; rep < 2; ++rep) {
                    late_head = (late_head == 0 ? A.late_size : late_head) - 1;
                    if (lane == 0)
                        for (int l = 0; l < GJ_MAX_LEVELS; ++l) late_g[(size_t)late_head * GJ_MAX_LEVELS + l] = sc.v[l];
                    late_len = min(late_len + 1, A.late_size);
                }
            }
            if (gj_score_le(sc, top, LV)) {
                // agent_top = the neighbour: best_row := row with the move's stops re-labelled
                materialize_top();
                if (lane < ncs) {
                    const int st = q.cs_stop[lane];
                    best_row[2 * st] = q.cs_v[lane];
                    best_row[2 * st + 1] = q.cs_c[lane];
                    diff[lane] = st;
                }
                ndiff = ncs;
                top = sc;
                top_pending = false;
            }
            __syncwarp();
        } else if (accept) {
            cur_is_g = false;
            if (top_pending && !gj_score_le(sc, top, LV)) {
                materialize_top();
                top_pending = false;
            }
            if (ndiff >= 0) {
                if (ndiff + ncs <= GJ_VRPC_DIFF) {
                    if (lane < ncs) diff[ndiff + lane] = q.cs_stop[lane];
                    ndiff += ncs;
                } else {
                    ndiff = -1;
                }
            }
            // ---- apply: labels, counts, route slots, totals ------------------------------------------------
            if (lane == 0) {
                for (int j = 0; j < ncs; ++j) {
                    const int stop = q.cs_stop[j];
                    row[2 * stop] = q.cs_v[j];
                    if (q.cs_oc[j] != q.cs_c[j]) {
                        row[2 * stop + 1] = q.cs_c[j];
                        cnt[q.cs_oc[j] - P.val_lo] -= 1;
                        cnt[q.cs_c[j] - P.val_lo] += 1;
                    }
                }
                tot[0] = (unsigned long long)dups; tot[1] = cap_pen; tot[2] = late_pen;
            }
            for (int a = 0; a < nav; ++a) {
                const int v = q.av[a], len = q.nlen[a];
                const int32_t* src = spare + q.noff[a];
                int32_t* dst = rs + (size_t)v * n;
                for (int i = lane; i < len; i += 32) dst[i] = src[i];
                if (lane == 0) { rlen[v] = len; rdist[v] = q.nd[a]; rload[v] = q.nl[a]; rlate[v] = q.nt[a]; }
            }
            cur = sc;
            accepted_total += 1;
            if (is_la) {
                late_head = (late_head == 0 ? A.late_size : late_head) - 1;
                if (lane == 0)
                    for (int l = 0; l < GJ_MAX_LEVELS; ++l) late_g[(size_t)late_head * GJ_MAX_LEVELS + l] = sc.v[l];
                late_len = min(late_len + 1, A.late_size);
            }
            __syncwarp();
            // update_top_individual (agent_base.rs:220-224)
            if (gj_score_le(cur, top, LV)) { top = cur; top_pending = true; }
            check_pending = false;
        } else if (check_pending) {
            // nothing moved: the solution that arrived between steps meets the agent's top now
            check_pending = false;
            if (gj_score_le(cur, top, LV)) { top = cur; top_pending = true; }
        }
        // ---- tabu deque (mover.rs:75-96): every id the move selected enters, the oldest leave ------
        if (tabu_g && m.kind != GJ_MOVE_NULL && lane == 0) {
            int sel[GJ_MOVE_MAXK];
            const int cntsel = gj_move_selected(m, sel);
            const int glen = G.offsets[m.group + 1] - G.offsets[m.group];
            const int W = (glen + 31) >> 5;
            const int T = A.tabu_size[m.group];
            uint32_t* bits = tabu_g + A.ctabu_off[m.group];
            int32_t* ring = (int32_t*)(bits + W + 1);
            int head = ring[T], fill = ring[T + 1];
#pragma unroll 1
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

    // ---- finish -----------------------------------------------------------------------------------------
    if (top_pending) materialize_top();
    if (lane == 0) {
        V.ndiff[island] = ndiff;
        V.pend[island] = check_pending ? 1 : 0;
        if (A.gseen) A.gseen[island] = cur_is_g ? g_ver : 0;
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) {
            A.cur_score[(size_t)island * GJ_MAX_LEVELS + l] = cur.v[l];
            A.best_score[(size_t)island * GJ_MAX_LEVELS + l] = top.v[l];
        }
        if (is_la) { A.late_head[island] = late_head; A.late_len[island] = late_len; }
        else for (int l = 0; l < GJ_MAX_LEVELS; ++l) A.sa_temp[(size_t)island * GJ_MAX_LEVELS + l] = temp[l];
        atomicAdd(&A.counters[0], (unsigned long long)A.n_steps);
        if (island == 0) atomicAdd(&A.counters[1], (unsigned long long)A.n_steps);
        if (accepted_total) atomicAdd(&A.counters[2], (unsigned long long)accepted_total);
    }
}
