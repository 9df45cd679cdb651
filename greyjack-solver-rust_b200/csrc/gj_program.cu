// gj_program.cu -- constraint programs: the device-side counterpart of the reference's constraint
// REGISTRY (greyjack/src/score_calculation/score_calculators/plain_score_calculator.rs:20-94).
//
// A reference user registers named constraints -- closures over Polars frames -- with
// add_constraint / remove_constraint / set_constraint_weights, and get_score sums
// weight_i * score_i over them (:60-94).  A closure cannot cross to the GPU; what the examples'
// closures DO is a handful of relational patterns over the planning columns:
//
//   GJ_OP_DISTINCT_DEFICIT   count(col) - n_unique(expr)       nqueens all_different (:37-59 of the
//                            expr = cv * value + ci * index     example PSC), tsp / vrp no_duplicating_stops
//   GJ_OP_GATHER_FOLD        sequential fold of table[prev][cur] over a column in row order, closed
//                            through a depot                     tsp minimize_distance (:70-84),
//                            (per segment, summed in segment order, when a segment column is given)
//                                                                vrp minimize_distance (:142-167)
//   GJ_OP_SEGMENT_OVER_CAP   per segment: sum of a fact of the column's values against the segment's
//                            capacity, overflow summed          vrp capacity constraint (:95-107)
//   GJ_OP_MAXPLUS_LATENESS   per segment: arrival = max(arrival, start) + service along the route,
//                            lateness against the window summed vrp late_arrival_penalty (:191-230)
//
// so a constraint here is a NAME, a score LEVEL, a WEIGHT and up to four TERMS, each one of those
// primitives with its column selectors; the program is interpreted by one kernel per call:
//   * un-segmented terms (N-Queens, TSP): one warp per candidate, distinct counts through a
//     shared-memory bitmap, gathers + warp reductions / in-order folds (the arithmetic of gj_eval.cuh);
//   * segmented terms (the VRP family): one CTA per candidate shares ONE bucketing of the stops by
//     segment between all the terms (gj_vrp_eval_cta) -- they must use the problem's (vehicle,
//     customer) column layout and its fact tables.
// Scores are sum_i weight_i * (sum_t scale_t * term_t) per level, constraints in insertion order
// (the reference iterates a HashMap: its order is unspecified).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "gj_eval.cuh"
#include "gj_internal.hpp"

static constexpr int kProgMaxConstraints = 16;
static constexpr int kProgWarps = 4;

struct GjTermDev {
    int op, value_offset, value_stride, n_rows;
    int key_value_coef, key_index_coef, key_lo, key_words;   // DISTINCT_DEFICIT
    int variant;
    double scale;
};
struct GjConstraintDev {
    int level, n_terms;
    double weight;
    GjTermDev terms[GJ_PROGRAM_MAX_TERMS];
};
struct GjProgramDev {
    int n_constraints, levels, bm_words;
    GjConstraintDev c[kProgMaxConstraints];
};

struct gj_program {
    gj_problem* p = nullptr;
    int levels = 1;
    std::vector<gj_constraint> constraints;
    std::vector<double> weights;
    bool segmented = false;
};

// value of row r of a term's column for one candidate (decoded like every planning variable)
__device__ __forceinline__ int gj_prog_value(const GjProblemDev& P, const double* row, const GjTermDev& t, int r) {
    const int var = t.value_offset + r * t.value_stride;
    return gj_decode(P, var, row[var]);
}

// ---- un-segmented programs: one warp per candidate ---------------------------------------------------
__global__ void __launch_bounds__(kProgWarps * 32)
k_program_warp(GjProblemDev P, const __grid_constant__ GjProgramDev G, const double* __restrict__ samples, int64_t S,
               double* __restrict__ scores) {
    extern __shared__ uint32_t smem_u32[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* bm = smem_u32 + (size_t)warp * G.bm_words;
    const int64_t j = (int64_t)blockIdx.x * kProgWarps + warp;
    if (j >= S) return;
    const double* row = samples + j * (int64_t)P.n_vars;
    double level_sum[GJ_MAX_LEVELS] = {0.0, 0.0, 0.0};
    for (int ci = 0; ci < G.n_constraints; ++ci) {
        const GjConstraintDev& C = G.c[ci];
        double acc = 0.0;
        for (int ti = 0; ti < C.n_terms; ++ti) {
            const GjTermDev& T = C.terms[ti];
            double result = 0.0;
            if (T.op == GJ_OP_DISTINCT_DEFICIT) {
                for (int w = lane; w < T.key_words; w += 32) bm[w] = 0u;
                __syncwarp();
                for (int r = lane; r < T.n_rows; r += 32) {
                    const int v = gj_prog_value(P, row, T, r);
                    const int idx = P.column_id ? P.column_id[T.value_offset + r * T.value_stride] : r;
                    const unsigned k = (unsigned)(T.key_value_coef * v + T.key_index_coef * idx - T.key_lo);
                    atomicOr(&bm[k >> 5], 1u << (k & 31));
                }
                __syncwarp();
                int u = 0;
                for (int w = lane; w < T.key_words; w += 32) u += __popc(bm[w]);
                u = gj_warp_sum(u);
                result = (double)(T.n_rows - u);
                __syncwarp();
            } else if (T.op == GJ_OP_GATHER_FOLD) {
                // tsp PSC :70-84: ((0 + D[depot][s0]) + D[s_last][depot]) + fold_{i>=1} D[s_{i-1}][s_i]
                const size_t L = (size_t)P.n_locations;
                const double* __restrict__ D = P.D;
                const int n = T.n_rows;
                double fold = 0.0, tree = 0.0, head = 0.0;
                int carry = 0;
                for (int base = 0; base < n; base += 32) {
                    const int i = base + lane;
                    const int v = (i < n) ? gj_prog_value(P, row, T, i) : 0;
                    int prev = __shfl_up_sync(GJ_FULL_MASK, v, 1);
                    if (lane == 0) prev = carry;
                    carry = __shfl_sync(GJ_FULL_MASK, v, 31);
                    double d = 0.0;
                    if (i < n) d = __ldg(&D[(size_t)prev * L + (size_t)v]);
                    if (P.exact_sums) {
                        if (base == 0) { head = __shfl_sync(GJ_FULL_MASK, d, 0); if (lane == 0) d = 0.0; }
                        const int m = min(32, n - base);
                        for (int l = 0; l < m; ++l) fold = fold + __shfl_sync(GJ_FULL_MASK, d, l);
                    } else {
                        tree += d;
                    }
                }
                double closing = 0.0;
                if (lane == 0) closing = __ldg(&D[(size_t)gj_prog_value(P, row, T, n - 1) * L]);
                closing = __shfl_sync(GJ_FULL_MASK, closing, 0);
                if (P.exact_sums) {
                    double sample_distance = 0.0;
                    sample_distance += head;
                    sample_distance += closing;
                    sample_distance += fold;
                    result = sample_distance;
                } else {
                    result = gj_warp_sum(tree) + closing;
                }
            }
            acc += T.scale * result;
        }
        level_sum[C.level] += C.weight * acc;
    }
    if (lane == 0)
        for (int l = 0; l < G.levels; ++l) scores[j * G.levels + l] = level_sum[l];
}

// ---- segmented programs (VRP family): one CTA per candidate, one shared bucketing ------------------------
__global__ void __launch_bounds__(kVrpWarps * 32)
k_program_vrp(GjProblemDev P, const __grid_constant__ GjProgramDev G, const double* __restrict__ samples, int64_t S,
              int tw_mode, double* __restrict__ scores) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = P.n_entities;
    GjVrpSmem s = gj_vrp_carve(smem_raw, n, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
    const int64_t j = blockIdx.x;
    const double* row = samples + j * (int64_t)P.n_vars;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double2 pr = *reinterpret_cast<const double2*>(row + 2 * i);
        s.veh[i] = (uint16_t)gj_decode(P, 2 * i, pr.x);
        s.cust[i] = gj_decode(P, 2 * i + 1, pr.y);
    }
    __syncthreads();
    double dup1000 = 0, cap = 0, dist = 0, late = 0;
    gj_vrp_eval_cta(P, s, tw_mode, dup1000, cap, dist, late);
    if (threadIdx.x == 0) {
        double level_sum[GJ_MAX_LEVELS] = {0.0, 0.0, 0.0};
        for (int ci = 0; ci < G.n_constraints; ++ci) {
            const GjConstraintDev& C = G.c[ci];
            double acc = 0.0;
            for (int ti = 0; ti < C.n_terms; ++ti) {
                const GjTermDev& T = C.terms[ti];
                double result = 0.0;
                if (T.op == GJ_OP_DISTINCT_DEFICIT) result = dup1000 / 1000.0;
                else if (T.op == GJ_OP_GATHER_FOLD) result = dist;
                else if (T.op == GJ_OP_SEGMENT_OVER_CAP) result = cap;
                else if (T.op == GJ_OP_MAXPLUS_LATENESS) result = late;
                acc += T.scale * result;
            }
            level_sum[C.level] += C.weight * acc;
        }
        for (int l = 0; l < G.levels; ++l) scores[j * G.levels + l] = level_sum[l];
    }
}

// ---- host ----------------------------------------------------------------------------------------------
extern "C" gj_status gj_program_create(gj_problem* p, int32_t levels, gj_program** out) {
    if (!p || !out || levels < 1 || levels > GJ_MAX_LEVELS) return gj_fail(GJ_ERR_INVALID, "bad argument");
    std::unique_ptr<gj_program> g(new gj_program());
    g->p = p;
    g->levels = levels;
    *out = g.release();
    return GJ_OK;
}

extern "C" void gj_program_destroy(gj_program* g) { delete g; }

static int find_constraint(const gj_program* g, const char* name) {
    for (size_t i = 0; i < g->constraints.size(); ++i)
        if (std::strncmp(g->constraints[i].name, name, sizeof(g->constraints[i].name)) == 0) return (int)i;
    return -1;
}

// PlainScoreCalculator::add_constraint (:29-34): a new name gets weight 1.0, a known name keeps its weight
extern "C" gj_status gj_program_add_constraint(gj_program* g, const gj_constraint* c) {
    if (!g || !c) return gj_fail(GJ_ERR_INVALID, "null argument");
    if (c->n_terms < 1 || c->n_terms > GJ_PROGRAM_MAX_TERMS) return gj_fail(GJ_ERR_INVALID, "1..4 terms per constraint");
    if (c->level < 0 || c->level >= g->levels) return gj_fail(GJ_ERR_INVALID, "score level out of range");
    if (c->name[0] == 0 || !std::memchr(c->name, 0, sizeof(c->name))) return gj_fail(GJ_ERR_INVALID, "constraint name missing / not terminated");
    const GjProblemDev& P = g->p->dev;
    for (int t = 0; t < c->n_terms; ++t) {
        const gj_term& T = c->terms[t];
        if (T.op < GJ_OP_DISTINCT_DEFICIT || T.op > GJ_OP_MAXPLUS_LATENESS) return gj_fail(GJ_ERR_INVALID, "unknown primitive");
        if (T.value_stride < 1 || T.value_offset < 0 || T.value_offset >= T.value_stride)
            return gj_fail(GJ_ERR_INVALID, "column selector: 0 <= offset < stride");
        const bool seg = T.seg_stride != 0;
        if (seg) {
            // segmented terms share the problem's bucketing: (vehicle, customer) column pairs only
            if (P.kind < GJ_VRP || T.value_offset != 1 || T.value_stride != 2 || T.seg_offset != 0 || T.seg_stride != 2)
                return gj_fail(GJ_ERR_UNSUPPORTED, "segmented terms need the VRP layout: segment column 0/2, value column 1/2");
        } else {
            if (T.op == GJ_OP_SEGMENT_OVER_CAP || T.op == GJ_OP_MAXPLUS_LATENESS)
                return gj_fail(GJ_ERR_INVALID, "this primitive needs a segment column");
            if (P.kind >= GJ_VRP) return gj_fail(GJ_ERR_UNSUPPORTED, "un-segmented terms on a VRP problem");
            if (T.op == GJ_OP_GATHER_FOLD && !P.D) return gj_fail(GJ_ERR_INVALID, "GATHER_FOLD needs the problem's distance matrix");
        }
        if (T.op == GJ_OP_DISTINCT_DEFICIT && seg && (T.key_value_coef != 1 || T.key_index_coef != 0))
            return gj_fail(GJ_ERR_UNSUPPORTED, "segmented programs count distinct values only");
    }
    const int at = find_constraint(g, c->name);
    if (at >= 0) {
        g->constraints[at] = *c;
    } else {
        if ((int)g->constraints.size() >= kProgMaxConstraints) return gj_fail(GJ_ERR_UNSUPPORTED, "too many constraints");
        g->constraints.push_back(*c);
        g->weights.push_back(1.0);
    }
    return GJ_OK;
}

extern "C" gj_status gj_program_remove_constraint(gj_program* g, const char* name) {
    if (!g || !name) return gj_fail(GJ_ERR_INVALID, "null argument");
    const int at = find_constraint(g, name);
    if (at < 0) return GJ_OK;                       // HashMap::remove of a missing key is a no-op
    g->constraints.erase(g->constraints.begin() + at);
    g->weights.erase(g->weights.begin() + at);
    return GJ_OK;
}

// set_constraint_weights (:40-42) replaces the whole map: every registered constraint must be named
// (the reference panics in get_score on a missing key, :83)
extern "C" gj_status gj_program_set_constraint_weights(gj_program* g, const char* const* names, const double* weights, int32_t n) {
    if (!g || (n > 0 && (!names || !weights))) return gj_fail(GJ_ERR_INVALID, "null argument");
    std::vector<double> w(g->constraints.size(), 0.0);
    std::vector<char> seen(g->constraints.size(), 0);
    for (int i = 0; i < n; ++i) {
        const int at = find_constraint(g, names[i]);
        if (at >= 0) { w[at] = weights[i]; seen[at] = 1; }
    }
    for (size_t i = 0; i < seen.size(); ++i)
        if (!seen[i]) return gj_fail(GJ_ERR_INVALID, std::string("no weight for constraint ") + g->constraints[i].name);
    g->weights = w;
    return GJ_OK;
}

extern "C" int32_t gj_program_n_constraints(const gj_program* g) { return g ? (int32_t)g->constraints.size() : 0; }

// PlainScoreCalculator::get_score behind request_score_plain: samples host [S][n_vars], scores host [S][levels]
extern "C" gj_status gj_program_get_score(gj_program* g, const double* samples, int64_t S, double* scores) {
    if (!g || !samples || !scores || S < 0) return gj_fail(GJ_ERR_INVALID, "bad argument");
    if (S == 0) return GJ_OK;
    if (g->constraints.empty()) return gj_fail(GJ_ERR_INVALID, "no constraints registered");     // scores_vec[0] panics (:77)
    gj_problem* p = g->p;
    const GjProblemDev& P = p->dev;
    GJ_CUDA_TRY(cudaSetDevice(p->device));
    // ---- lower the registry to the device program ------------------------------------------------------
    GjProgramDev D{};
    D.n_constraints = (int)g->constraints.size();
    D.levels = g->levels;
    bool segmented = false;
    int tw_mode = GJ_TW_PSC;
    int bm_words = 1;
    std::vector<int32_t> colid;
    for (int ci = 0; ci < D.n_constraints; ++ci) {
        const gj_constraint& c = g->constraints[ci];
        GjConstraintDev& C = D.c[ci];
        C.level = c.level; C.n_terms = c.n_terms; C.weight = g->weights[ci];
        for (int t = 0; t < c.n_terms; ++t) {
            const gj_term& T = c.terms[t];
            GjTermDev& X = C.terms[t];
            X.op = T.op; X.value_offset = T.value_offset; X.value_stride = T.value_stride;
            X.n_rows = (P.n_vars - T.value_offset + T.value_stride - 1) / T.value_stride;
            X.key_value_coef = T.key_value_coef; X.key_index_coef = T.key_index_coef;
            X.variant = T.variant; X.scale = T.scale;
            if (T.seg_stride != 0) {
                segmented = true;
                if (T.op == GJ_OP_MAXPLUS_LATENESS) tw_mode = T.variant;
            } else if (T.op == GJ_OP_DISTINCT_DEFICIT) {
                // key range from the variables' bounds and the row indices
                long long lo = LLONG_MAX, hi = LLONG_MIN;
                if (P.column_id && colid.empty()) {
                    colid.resize(P.n_vars);
                    GJ_CUDA_TRY(cudaMemcpy(colid.data(), P.column_id, (size_t)P.n_vars * 4, cudaMemcpyDeviceToHost));
                }
                for (int r = 0; r < X.n_rows; ++r) {
                    const int var = T.value_offset + r * T.value_stride;
                    const long long idx = P.column_id ? colid[var] : r;
                    double a = p->lb[var], b = p->ub[var];
                    if (p->frozen[var]) a = b = p->initial[var];
                    const long long k1 = (long long)T.key_value_coef * (long long)std::llround(a) + (long long)T.key_index_coef * idx;
                    const long long k2 = (long long)T.key_value_coef * (long long)std::llround(b) + (long long)T.key_index_coef * idx;
                    lo = std::min(lo, std::min(k1, k2)); hi = std::max(hi, std::max(k1, k2));
                }
                if (hi - lo > (1ll << 26)) return gj_fail(GJ_ERR_UNSUPPORTED, "distinct-count key range too wide");
                X.key_lo = (int)lo;
                X.key_words = (int)((hi - lo + 1 + 31) / 32);
                bm_words = std::max(bm_words, X.key_words);
            }
        }
    }
    D.bm_words = bm_words;
    for (int ci = 0; ci < D.n_constraints; ++ci)
        for (int t = 0; t < D.c[ci].n_terms; ++t)
            if ((g->constraints[ci].terms[t].seg_stride != 0) != segmented)
                return gj_fail(GJ_ERR_UNSUPPORTED, "a program is either segmented (VRP layout) or not");
    // ---- run ----------------------------------------------------------------------------------------------
    const size_t in_bytes = (size_t)S * (size_t)P.n_vars * 8, out_bytes = (size_t)S * (size_t)g->levels * 8;
    gj_status rc;
    if ((rc = p->d_samples.reserve(in_bytes))) return rc;
    if ((rc = p->d_scores.reserve(out_bytes))) return rc;
    cudaStream_t st = p->stream;
    GJ_CUDA_TRY(cudaMemcpyAsync(p->d_samples.ptr, samples, in_bytes, cudaMemcpyHostToDevice, st));
    if (segmented) {
        const size_t smem = gj_vrp_smem_bytes(P.n_entities, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
        if (smem > 48 * 1024)
            GJ_CUDA_TRY(cudaFuncSetAttribute(k_program_vrp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_program_vrp<<<(unsigned)S, kVrpWarps * 32, smem, st>>>(P, D, (const double*)p->d_samples.ptr, S, tw_mode,
                                                                 (double*)p->d_scores.ptr);
    } else {
        const size_t smem = (size_t)kProgWarps * (size_t)bm_words * 4;
        if (smem > 48 * 1024)
            GJ_CUDA_TRY(cudaFuncSetAttribute(k_program_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_program_warp<<<(unsigned)((S + kProgWarps - 1) / kProgWarps), kProgWarps * 32, smem, st>>>(
            P, D, (const double*)p->d_samples.ptr, S, (double*)p->d_scores.ptr);
    }
    GJ_LAUNCH_CHECK();
    GJ_CUDA_TRY(cudaMemcpyAsync(scores, p->d_scores.ptr, out_bytes, cudaMemcpyDeviceToHost, st));
    GJ_CUDA_TRY(cudaStreamSynchronize(st));
    return GJ_OK;
}
