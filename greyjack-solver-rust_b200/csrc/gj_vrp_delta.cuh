// gj_vrp_delta.cuh -- route-level delta evaluation for the VRP models (included by gj_islands.cu).
//
// The reference re-scores every neighbour from scratch (vrp ISC :58-137: clone both columns, hash
// the customers, bucket the stops by vehicle, walk every route).  A change / swap / swap_edges /
// scramble move touches at most 16 stops, i.e. a handful of routes.  An island therefore keeps
// the bucketed form of its current solution in HBM,
//     bucket / bstop [n]   customers and stop indices grouped by vehicle, stop order kept
//     start [K + 1]        route boundaries
//     rdist / rload / rlate [K]   every route's distance fold, demand and lateness
//     cnt [locations]      customer occurrence counts;  totals: duplicates, capacity, lateness
// and ONE THREAD scores a neighbour by re-walking only the routes that gain, lose or re-label a
// stop -- in the reference's own order (route folds, then vehicle_distances.iter().sum() over all
// K vehicles with the re-walked routes substituted), so every level, the float one included, is
// bit-identical to the full evaluation.  Segment moves (insertion / inverse) re-label O(segment)
// stops and go to the full evaluator.
#pragma once

struct GjVrpState {
    int32_t* bucket;                 // [I][n]
    int32_t* bstop;                  // [I][n]
    int32_t* start;                  // [I][K + 1]
    double* rdist;                   // [I][K]
    unsigned long long* rload;       // [I][K]
    unsigned long long* rlate;       // [I][K]
    unsigned long long* tot;         // [I][4]: duplicates, capacity penalty, lateness, -
};

// Route index of a single-neighbour chain (gj_islands_vrp_chain.cuh)
struct GjVrpChainState {
    int32_t* rs;                 // [I + 1][K][n]
    int32_t* rlen;               // [I][K]
    double* rdist;               // [I][K]
    unsigned long long* rload;   // [I][K]
    unsigned long long* rlate;   // [I][K]
    unsigned long long* tot;     // [I][4]: duplicates, capacity penalty, lateness, -
    int32_t* spare;              // [I][n]
    int32_t* cnt; int cnt_stride;
    int* stale;                  // [I + 1] route index out of date (creation, migrant)
    // slot I of every array above = the route index of the published global top, built once per
    // published version (k_vrp_chain_gindex); adopting chains copy it instead of re-walking K routes
    int32_t* gstop; int32_t* gdst;   // [n] its stop lists flattened: rs[gdst[p]] = gstop[p]
    int* gidx_ver; int32_t* goff;    // [K] offsets of the routes in the flattened lists
    int32_t* diff; int* ndiff;       // [I][GJ_VRPC_DIFF], [I]: stops where the chain differs from its top row
    int* pend;                       // [I] update_top_individual still owes a comparison for a solution that
                                     // arrived between steps (migrant / adopted global top)
};

#define GJ_VRP_MAXCS 16              // changed stops of a small move
#define GJ_VRP_MAXAV 32              // routes they can touch

// Returns false when the move needs the full evaluator.  On success the four unweighted terms of
// the neighbour (same meaning as gj_vrp_eval_cta's outputs).
__device__ __forceinline__ bool gj_vrp_move_delta(const GjProblemDev& P, const GjGroups& G, const GjMove& m,
                                                  bool noop_quirk, int tw_mode,
                                                  const int32_t* __restrict__ row,          // cur [2n]
                                                  const int32_t* __restrict__ bucket, const int32_t* __restrict__ bstop,
                                                  const int32_t* __restrict__ start, const double* __restrict__ rdist,
                                                  const unsigned long long* __restrict__ rload,
                                                  const unsigned long long* __restrict__ rlate,
                                                  const unsigned long long* __restrict__ tot,
                                                  const int32_t* __restrict__ cnt,
                                                  double& dup1000, double& cap, double& dist, double& late) {
    const int n = P.n_entities, K = P.n_vehicles;
    const size_t L = (size_t)P.n_locations;
    const double* __restrict__ D = P.D;
    int ncs = 0, nav = 0;
    int cs_stop[GJ_VRP_MAXCS], cs_v[GJ_VRP_MAXCS], cs_c[GJ_VRP_MAXCS];
    int av[GJ_VRP_MAXAV];
    double nd[GJ_VRP_MAXAV];
    unsigned long long nl[GJ_VRP_MAXAV], nt[GJ_VRP_MAXAV];
    int d_uniq = 0;
    const bool identity = m.kind == GJ_MOVE_NULL || (noop_quirk && (m.kind == 3 || (m.kind == 2 && m.k == 2)));
    if (!identity) {
        if (m.kind > 3) return false;
        const GjMove ms = m;
        const int32_t* g = G.ids + G.offsets[ms.group];
        int cols[GJ_MOVE_MAXPAIRS], vals[GJ_MOVE_MAXPAIRS];
        const int np = gj_small_move_pairs(ms, g, true, noop_quirk, [&](int id) { return row[id]; }, cols, vals);
        // (column, value) pairs in emission order -> changed stops (var-wise application, later wins)
        for (int i = 0; i < np; ++i) {
            const int col = cols[i], val = gj_fix_column(P, col, vals[i]);
            const int stop = col >> 1;
            int j = 0;
            for (; j < ncs; ++j) if (cs_stop[j] == stop) break;
            if (j == ncs) { cs_stop[j] = stop; cs_v[j] = row[2 * stop]; cs_c[j] = row[2 * stop + 1]; ++ncs; }
            if (col & 1) cs_c[j] = val; else cs_v[j] = val;
        }
        // customer multiset -> duplicates
        int ko[GJ_VRP_MAXCS], kn[GJ_VRP_MAXCS];
        for (int j = 0; j < ncs; ++j) {
            const int oc = row[2 * cs_stop[j] + 1];
            ko[j] = -1; kn[j] = -1;
            if (oc != cs_c[j]) { ko[j] = oc - P.val_lo; kn[j] = cs_c[j] - P.val_lo; }
        }
        d_uniq = gj_uniq_delta(cnt, ko, kn, ncs);
        // routes that gain, lose or re-label a stop
        auto touch = [&](int v) {
            for (int a = 0; a < nav; ++a) if (av[a] == v) return;
            av[nav++] = v;
        };
        for (int j = 0; j < ncs; ++j) {
            const int ov = row[2 * cs_stop[j]], oc = row[2 * cs_stop[j] + 1];
            if (ov == cs_v[j] && oc == cs_c[j]) continue;
            touch(ov);
            touch(cs_v[j]);
        }
        // re-walk every touched route in stop order: its surviving stops merged with the arrivals
        for (int a = 0; a < nav; ++a) {
            const int v = av[a];
            int ins[GJ_VRP_MAXCS], ni = 0;                   // arrivals (changed stops coming from another route)
            for (int j = 0; j < ncs; ++j)
                if (cs_v[j] == v && row[2 * cs_stop[j]] != v) {
                    int q = ni++;
                    while (q > 0 && cs_stop[ins[q - 1]] > cs_stop[j]) { ins[q] = ins[q - 1]; --q; }
                    ins[q] = j;
                }
            int s = start[v], ip = 0;
            const int e = start[v + 1];
            int first = -1, last = -1;
            double fold = 0.0;
            unsigned long long load = 0ull, lateness = 0ull, arrival = P.day_start[v];
            auto visit = [&](int c) {
                if (first < 0) first = c; else fold = fold + __ldg(&D[(size_t)last * L + (size_t)c]);
                last = c;
                const uint4 f = P.cust[c];
                load += (unsigned long long)f.x;
                if (P.time_windowed) {
                    const unsigned long long ws = f.y, we = f.z, sv = f.w;
                    if (arrival < ws) arrival = ws;
                    if (tw_mode == GJ_TW_ISC_FILE) {
                        if (arrival + sv > we) lateness += (arrival + sv) - we;
                    } else {
                        if (arrival > we + sv) lateness += arrival - (we + sv);
                    }
                    arrival += sv;
                }
            };
            while (s < e || ip < ni) {
                const int so = (s < e) ? bstop[s] : 0x7fffffff;
                const int si = (ip < ni) ? cs_stop[ins[ip]] : 0x7fffffff;
                if (so < si) {
                    int j = 0;
                    for (; j < ncs; ++j) if (cs_stop[j] == so) break;
                    if (j == ncs) visit(bucket[s]);
                    else if (cs_v[j] == v) visit(cs_c[j]);   // stays on the route (maybe another customer)
                    ++s;
                } else {
                    visit(cs_c[ins[ip]]);
                    ++ip;
                }
            }
            double current_distance = 0.0;
            if (first >= 0) {
                const size_t depot = (size_t)P.veh_depot[v];
                current_distance += __ldg(&D[depot * L + (size_t)first]);
                current_distance += __ldg(&D[(size_t)last * L + depot]);
                current_distance += fold;
                if (P.time_windowed && arrival > P.day_end[v]) lateness += arrival - P.day_end[v];
            }
            nd[a] = current_distance; nl[a] = load; nt[a] = lateness;
        }
    }
    // totals: integers by difference, the distance by the reference's own sequential vehicle sum
    unsigned long long cap_pen = tot[1], late_pen = tot[2];
    for (int a = 0; a < nav; ++a) {
        const int v = av[a];
        const unsigned long long capv = P.veh_capacity[v];
        const unsigned long long ol = rload[v];
        if (ol > capv) cap_pen -= ol - capv;
        if (nl[a] > capv) cap_pen += nl[a] - capv;
        late_pen -= rlate[v];
        late_pen += nt[a];
    }
    double sum_distance = 0.0;
    for (int v = 0; v < K; ++v) {
        double x = rdist[v];
        for (int a = 0; a < nav; ++a) if (av[a] == v) x = nd[a];
        sum_distance += x;
    }
    const long long dups = (long long)tot[0] - (long long)d_uniq;
    dup1000 = 1000.0 * (double)dups;
    cap = (double)cap_pen;
    late = (double)late_pen;
    dist = sum_distance;
    (void)n;
    return true;
}
