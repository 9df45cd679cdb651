/*
 * gj_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See gj_oracle.h for scope, parity status ("parity unpinned" at the scorer
 * boundary) and build flags.  Reference paths are relative to the reference root.
 */
#define _POSIX_C_SOURCE 200809L
#include "gj_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------- */
/* value helpers                                                             */
/* ------------------------------------------------------------------------- */

/* f64::total_cmp (Rust std): sign-magnitude bits mapped to two's complement. */
static int total_cmp(double a, double b) {
    int64_t l, r;
    memcpy(&l, &a, 8);
    memcpy(&r, &b, 8);
    l ^= (int64_t)(((uint64_t)(l >> 63)) >> 1);
    r ^= (int64_t)(((uint64_t)(r >> 63)) >> 1);
    return (l < r) ? -1 : (l > r) ? 1 : 0;
}

/* greyjack/src/utils/math_utils.rs:6-8 -- ties go to ceil. */
double gjo_rint(double x) {
    double f = floor(x), c = ceil(x);
    if (fabs(x - f) < fabs(c - x)) return f;
    return c;
}

/* greyjack/src/utils/math_utils.rs:10-13 -- truncation toward -inf at `precision`
   decimals (not a rounding). */
double gjo_round(double value, uint64_t precision) {
    double multiplier = pow(10.0, (double)precision);
    double fl = floor(value);
    return fl + floor((value - fl) * multiplier) / multiplier;
}

/* greyjack/src/variables/gj_integer.rs:114-138 (min/max by total_cmp). */
static double gj_min(double a, double b) {
    int c = total_cmp(a, b);
    return (c < 0) ? a : (c > 0) ? b : a;
}
static double gj_max(double a, double b) {
    int c = total_cmp(a, b);
    return (c < 0) ? b : (c > 0) ? a : b;
}

/* greyjack/src/variables/gj_integer.rs:70-83 */
double gjo_fix_integer(double value, double lb, double ub, int frozen, double initial) {
    if (frozen) return initial;
    double fixed = gj_min(gj_max(value, lb), ub);
    return gjo_rint(fixed);
}

/* greyjack/src/variables/gj_integer.rs:66-68 (`as i64` saturates, NaN -> 0). */
int64_t gjo_inverse_transform_integer(double value, double lb, double ub, int frozen,
                                      double initial) {
    double f = gjo_fix_integer(value, lb, ub, frozen, initial);
    if (f != f) return 0;
    if (f >= 9223372036854775807.0) return INT64_MAX;
    if (f <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)f;
}

/* greyjack/src/variables/gj_float.rs:64-76 */
double gjo_fix_float(double value, double lb, double ub, int frozen, double initial) {
    if (frozen) return initial;
    return gj_min(gj_max(value, lb), ub);
}

int gjo_levels(int kind) {
    switch (kind) {
        case GJO_NQUEENS: return 1;      /* SimpleScore          */
        case GJO_TSP: return 2;          /* HardSoftScore        */
        default: return 3;               /* HardMediumSoftScore  */
    }
}

/* scores/simple_score.rs:70-76, hard_soft_score.rs:84-96,
   hard_medium_soft_score.rs:93-113 */
int gjo_score_cmp(const double* a, const double* b, int levels) {
    for (int l = 0; l < levels; ++l) {
        int c = total_cmp(a[l], b[l]);
        if (c != 0) return c;
    }
    return 0;
}

/* #[derive(PartialOrd)] on the score structs: lexicographic partial_cmp; `<=`
   is false as soon as an unordered (NaN) level is met. */
int gjo_score_le(const double* a, const double* b, int levels) {
    for (int l = 0; l < levels; ++l) {
        if (a[l] != a[l] || b[l] != b[l]) return 0;
        if (a[l] < b[l]) return 1;
        if (a[l] > b[l]) return 0;
    }
    return 1;
}

/* ScoreTrait::round (hard_soft_score.rs:76-79); precision[l] < 0 = leave. */
void gjo_score_round(double* s, const int64_t* precision, int levels) {
    for (int l = 0; l < levels; ++l)
        if (precision[l] >= 0) s[l] = gjo_round(s[l], (uint64_t)precision[l]);
}

/* get_fitness_value: simple_score.rs, hard_soft_score.rs:38-44,
   hard_medium_soft_score.rs:43-50 */
double gjo_fitness(const double* s, int levels) {
    if (levels == 1) return 1.0 - (1.0 / (s[0] + 1.0));
    if (levels == 2) {
        double h = 1.0 - (1.0 / (s[0] + 1.0));
        double f = 1.0 - (1.0 / (s[1] + 1.0));
        return 0.5 * h + 0.5 * f;
    }
    double h = 1.0 - (1.0 / (s[0] + 1.0));
    double m = 1.0 - (1.0 / (s[1] + 1.0));
    double f = 1.0 - (1.0 / (s[2] + 1.0));
    return 0.34 * h + 0.33 * m + 0.33 * f;
}

static int g_sort_levels;
static int sort_cmp(const void* a, const void* b) {
    return gjo_score_cmp((const double*)a, (const double*)b, g_sort_levels);
}
void gjo_sort_scores(double* scores, int n, int levels) {
    g_sort_levels = levels;
    qsort(scores, (size_t)n, sizeof(double) * (size_t)levels, sort_cmp);
}

/* ------------------------------------------------------------------------- */
/* scratch                                                                   */
/* ------------------------------------------------------------------------- */

typedef struct {
    int64_t* vals;        /* decoded candidate (clone of planning ids)       */
    int64_t* vals2;       /* second planning column (VRP customer ids)       */
    uint32_t* stamp;      /* HashSet stand-in: stamp[value - lo] == gen      */
    int64_t stamp_lo, stamp_len;
    uint32_t gen;
    uint64_t* trip_demand;
    int64_t* route_buf;   /* Vec<Vec<usize>> vehicle_stops, flattened        */
    int64_t* route_start;
    int64_t* route_fill;
    int n_vars;
} scratch_t;

static int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
static int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }

static int scratch_init(scratch_t* s, const gjo_problem* p) {
    memset(s, 0, sizeof(*s));
    s->n_vars = p->n_vars;
    int64_t lo = INT64_MAX, hi = INT64_MIN;
    for (int i = 0; i < p->n_vars; ++i) {
        lo = imin64(lo, (int64_t)floor(p->lower_bounds[i]));
        hi = imax64(hi, (int64_t)ceil(p->upper_bounds[i]));
        if (p->frozen && p->frozen[i] && p->initial) {
            lo = imin64(lo, (int64_t)floor(p->initial[i]));
            hi = imax64(hi, (int64_t)ceil(p->initial[i]));
        }
    }
    if (p->n_vars == 0) { lo = 0; hi = 0; }
    /* N-Queens diagonals span [lo - maxcol, hi + maxcol] */
    int64_t pad = (p->kind == GJO_NQUEENS) ? (int64_t)p->n_vars + 1 : 0;
    if (p->kind == GJO_NQUEENS && p->column_id) {
        for (int i = 0; i < p->n_vars; ++i) pad = imax64(pad, p->column_id[i] + 1);
    }
    s->stamp_lo = lo - pad;
    s->stamp_len = (hi + pad) - s->stamp_lo + 1;
    s->vals = (int64_t*)malloc(sizeof(int64_t) * (size_t)(p->n_vars + 1));
    s->vals2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(p->n_vars + 1));
    s->stamp = (uint32_t*)calloc((size_t)s->stamp_len, sizeof(uint32_t));
    int nv = p->n_vehicles > 0 ? p->n_vehicles : 1;
    s->trip_demand = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)nv);
    s->route_buf = (int64_t*)malloc(sizeof(int64_t) * (size_t)(p->n_vars + 1));
    s->route_start = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nv + 1));
    s->route_fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nv + 1));
    if (!s->vals || !s->vals2 || !s->stamp || !s->trip_demand || !s->route_buf ||
        !s->route_start || !s->route_fill)
        return -1;
    return 0;
}

static void scratch_free(scratch_t* s) {
    free(s->vals); free(s->vals2); free(s->stamp); free(s->trip_demand);
    free(s->route_buf); free(s->route_start); free(s->route_fill);
}

/* HashSet<..>::len() of a value stream */
static void set_begin(scratch_t* s) {
    if (++s->gen == 0) { memset(s->stamp, 0, sizeof(uint32_t) * (size_t)s->stamp_len); s->gen = 1; }
}
static int set_insert(scratch_t* s, int64_t v) {
    int64_t k = v - s->stamp_lo;
    if (k < 0 || k >= s->stamp_len) return 1; /* out of modelled range: treated distinct */
    if (s->stamp[k] == s->gen) return 0;
    s->stamp[k] = s->gen;
    return 1;
}

static int64_t decode_var(const gjo_problem* p, int i, double x) {
    return gjo_inverse_transform_integer(
        x, p->lower_bounds[i], p->upper_bounds[i], p->frozen ? p->frozen[i] : 0,
        p->initial ? p->initial[i] : 0.0);
}

/* ------------------------------------------------------------------------- */
/* constraint arithmetic on a decoded candidate                              */
/* ------------------------------------------------------------------------- */

/* examples/nqueens/src/score/incremental_score_calculator.rs:44-56 (ISC) and
   plain_score_calculator.rs:37-59 (PSC: len - n_unique per sample; same counts) */
static double nqueens_conflicts(const gjo_problem* p, scratch_t* s, const int64_t* rows) {
    int64_t n = p->n_vars;
    int64_t uniq_rows = 0, uniq_desc = 0, uniq_asc = 0;
    set_begin(s);
    for (int64_t i = 0; i < n; ++i) uniq_rows += set_insert(s, rows[i]);
    set_begin(s);
    for (int64_t i = 0; i < n; ++i) {
        int64_t col = p->column_id ? p->column_id[i] : i;
        uniq_desc += set_insert(s, col + rows[i]);
    }
    set_begin(s);
    for (int64_t i = 0; i < n; ++i) {
        int64_t col = p->column_id ? p->column_id[i] : i;
        uniq_asc += set_insert(s, col - rows[i]);
    }
    double a = (double)(n - uniq_rows);
    double b = (double)(n - uniq_desc);
    double c = (double)(n - uniq_asc);
    return a + b + c;
}

/* examples/tsp/src/score/incremental_score_calculator.rs:71-80 and
   plain_score_calculator.rs:34-43, 70-84 */
static void tsp_terms(const gjo_problem* p, scratch_t* s, const int64_t* stops,
                      double* dup, double* dist) {
    int64_t n = p->n_vars;
    const double* D = p->distance_matrix;
    int64_t L = p->n_locations;
    int64_t uniq = 0;
    set_begin(s);
    for (int64_t i = 0; i < n; ++i) uniq += set_insert(s, stops[i]);
    *dup = (double)(n - uniq);
    double d = 0.0;
    int64_t last = n - 1;
    d += D[0 * L + stops[0]];
    d += D[stops[last] * L + 0];
    double fold = 0.0;
    for (int64_t i = 1; i < n; ++i) fold = fold + D[stops[i - 1] * L + stops[i]];
    d += fold;
    *dist = d;
}

enum { TW_ISC_FILE = 0, TW_ISC_SERVICE = 1, TW_PSC = 2 };

/* examples/vrp/src/score/incremental_score_calculator.rs:58-137 (ISC, file variant),
   examples/vrp_service/src/score/incremental_score_calculator.rs:58-138 (service),
   examples/vrp/src/score/plain_score_calculator.rs:51-233 (PSC; identical file in
   vrp_service).  Returns the four constraint terms separately. */
static void vrp_terms(const gjo_problem* p, scratch_t* s, const int64_t* veh,
                      const int64_t* cust, int tw_mode, double* dup1000, double* cap,
                      double* dist, double* late) {
    int64_t n = p->n_vars / 2;
    int64_t K = p->n_vehicles;
    int64_t L = p->n_locations;
    const double* D = p->distance_matrix;

    /* no_duplicating_stops_constraint */
    int64_t uniq = 0;
    set_begin(s);
    for (int64_t i = 0; i < n; ++i) uniq += set_insert(s, cust[i]);
    *dup1000 = 1000.0 * (double)(n - uniq);

    /* capacity_constraint */
    for (int64_t v = 0; v < K; ++v) s->trip_demand[v] = 0;
    for (int64_t i = 0; i < n; ++i) s->trip_demand[veh[i]] += p->demand[cust[i]];
    int64_t capacity_penalty = 0;
    for (int64_t v = 0; v < K; ++v) {
        int64_t diff = (int64_t)p->vehicle_capacity[v] - (int64_t)s->trip_demand[v];
        if (diff < 0) capacity_penalty += -diff;
    }
    *cap = (double)capacity_penalty;

    /* vehicle_stops: Vec<Vec<usize>>, push in stop order */
    for (int64_t v = 0; v <= K; ++v) s->route_start[v] = 0;
    for (int64_t i = 0; i < n; ++i) s->route_start[veh[i] + 1]++;
    for (int64_t v = 0; v < K; ++v) s->route_start[v + 1] += s->route_start[v];
    for (int64_t v = 0; v < K; ++v) s->route_fill[v] = s->route_start[v];
    for (int64_t i = 0; i < n; ++i) s->route_buf[s->route_fill[veh[i]]++] = cust[i];

    double sum_distance = 0.0, sum_time_penalty = 0.0;
    for (int64_t v = 0; v < K; ++v) {
        int64_t len = s->route_start[v + 1] - s->route_start[v];
        double current_distance = 0.0, current_time_penalty = 0.0;
        if (len != 0) {
            const int64_t* st = s->route_buf + s->route_start[v];
            int64_t depot = p->vehicle_depot[v];
            int64_t last_id = len - 1;
            current_distance += D[depot * L + st[0]];
            current_distance += D[st[last_id] * L + depot];
            double fold = 0.0;
            for (int64_t i = 1; i <= last_id; ++i) fold = fold + D[st[i - 1] * L + st[i]];
            current_distance += fold;

            if (p->time_windowed) {
                uint64_t arrival = p->work_day_start[v];
                uint64_t day_end = p->work_day_end[v];
                /* PSC iterates 0..len-1, skipping the last stop
                   (plain_score_calculator.rs:205-216) */
                int64_t upto = (tw_mode == TW_PSC) ? len - 1 : len;
                for (int64_t i = 0; i < upto; ++i) {
                    uint64_t ws = p->tw_start[st[i]];
                    uint64_t we = p->tw_end[st[i]];
                    uint64_t sv = p->service_time[st[i]];
                    if (arrival < ws) arrival = ws;
                    if (tw_mode == TW_ISC_FILE) {
                        /* vrp ISC :120-121 */
                        if (arrival + sv > we)
                            current_time_penalty += (double)((arrival + sv) - we);
                    } else {
                        /* vrp_service ISC :121-122 and PSC :208-210 */
                        if (arrival > we + sv)
                            current_time_penalty += (double)(arrival - (we + sv));
                    }
                    arrival += sv;
                }
                if (arrival > day_end) current_time_penalty += (double)(arrival - day_end);
            }
        }
        sum_distance += current_distance;
        sum_time_penalty += current_time_penalty;
    }
    *dist = sum_distance;
    *late = sum_time_penalty;
}

/* Weighted sum of constraint scores: score_calculators/plain_score_calculator.rs:79-90
   (and incremental_score_calculator.rs:84-95): sum = null; sum += w_i * s_i.      */
static void combine_psc(const gjo_problem* p, scratch_t* s, double* out) {
    const double* w = p->weights;
    if (p->kind == GJO_NQUEENS) {
        double v = nqueens_conflicts(p, s, s->vals);
        double acc = 0.0; acc += w[0] * v;
        out[0] = acc;
    } else if (p->kind == GJO_TSP) {
        double dup, dist;
        tsp_terms(p, s, s->vals, &dup, &dist);
        double h = 0.0, f = 0.0;
        h += w[0] * dup;  f += w[0] * 0.0;
        h += w[1] * 0.0;  f += w[1] * dist;
        out[0] = h; out[1] = f;
    } else {
        double dup, cap, dist, late;
        vrp_terms(p, s, s->vals, s->vals2, TW_PSC, &dup, &cap, &dist, &late);
        double h = 0.0, m = 0.0, f = 0.0;
        h += w[0] * dup;  m += w[0] * 0.0;  f += w[0] * 0.0;
        h += w[1] * cap;  m += w[1] * 0.0;  f += w[1] * 0.0;
        h += w[2] * 0.0;  m += w[2] * 0.0;  f += w[2] * dist;
        if (p->time_windowed) { /* constraint removed otherwise: cotwin_builder.rs:296-298 */
            h += w[3] * 0.0;  m += w[3] * late;  f += w[3] * 0.0;
        }
        out[0] = h; out[1] = m; out[2] = f;
    }
}

static void combine_isc(const gjo_problem* p, scratch_t* s, double* out) {
    double w = p->weights[0];
    if (p->kind == GJO_NQUEENS) {
        double v = nqueens_conflicts(p, s, s->vals);
        double acc = 0.0; acc += w * v;
        out[0] = acc;
    } else if (p->kind == GJO_TSP) {
        double dup, dist;
        tsp_terms(p, s, s->vals, &dup, &dist);
        double h = 0.0, f = 0.0;
        h += w * dup; f += w * dist;
        out[0] = h; out[1] = f;
    } else {
        double dup, cap, dist, late;
        int mode = (p->kind == GJO_VRP_SERVICE) ? TW_ISC_SERVICE : TW_ISC_FILE;
        vrp_terms(p, s, s->vals, s->vals2, mode, &dup, &cap, &dist, &late);
        double hard = dup + cap; /* HardMediumSoftScore::new(unique_stops_penalty + capacity_penalty as f64, ..) */
        double h = 0.0, m = 0.0, f = 0.0;
        h += w * hard; m += w * late; f += w * dist;
        out[0] = h; out[1] = m; out[2] = f;
    }
}

/* split an interleaved decoded VRP vector [v0,c0,v1,c1,..] into two columns */
static void vrp_split(const gjo_problem* p, scratch_t* s, const int64_t* interleaved) {
    int64_t n = p->n_vars / 2;
    for (int64_t i = 0; i < n; ++i) {
        s->vals[i] = interleaved[2 * i];
        s->vals2[i] = interleaved[2 * i + 1];
    }
}

/* ------------------------------------------------------------------------- */
/* request_score_plain / request_score_incremental                           */
/* ------------------------------------------------------------------------- */

/* oop_score_requester.rs:336-355 -> cotwin.rs:45-57 -> plain_score_calculator.rs:60-94 */
int gjo_score_plain(const gjo_problem* p, const double* samples, int64_t S, double* out) {
    scratch_t s;
    if (scratch_init(&s, p) != 0) return -1;
    int levels = gjo_levels(p->kind);
    int64_t* dec = (int64_t*)malloc(sizeof(int64_t) * (size_t)(p->n_vars + 1));
    for (int64_t j = 0; j < S; ++j) {
        const double* x = samples + j * (int64_t)p->n_vars;
        /* inverse_transform_variables: variables_manager.rs:136-152 */
        for (int i = 0; i < p->n_vars; ++i) dec[i] = decode_var(p, i, x[i]);
        if (p->kind >= GJO_VRP) vrp_split(p, &s, dec);
        else memcpy(s.vals, dec, sizeof(int64_t) * (size_t)p->n_vars);
        combine_psc(p, &s, out + j * levels);
    }
    free(dec);
    scratch_free(&s);
    return 0;
}

/* Scores one pseudo-incremental sample given the decoded base. */
static void score_one_incremental(const gjo_problem* p, scratch_t* s, const int64_t* base_dec,
                                  int64_t* work, const uint64_t* var_ids, const double* values,
                                  int64_t k, double* out) {
    /* clone of the planning ids (tsp ISC :64, vrp ISC :63-64, nqueens ISC :42) */
    memcpy(work, base_dec, sizeof(int64_t) * (size_t)p->n_vars);
    /* inverse_transform_deltas (variables_manager.rs:154-176), then var-wise
       application in emission order.  The reference sorts the delta frame by
       (sample_id, row) with an unstable sort (oop_score_requester.rs:435), so
       the order among repeated ids is unspecified there; emission order is the
       order the stored individual is updated in (tabu_search_base.rs:175-178). */
    for (int64_t d = 0; d < k; ++d) {
        int v = (int)var_ids[d];
        work[v] = decode_var(p, v, values[d]);
    }
    if (p->kind >= GJO_VRP) vrp_split(p, s, work);
    else memcpy(s->vals, work, sizeof(int64_t) * (size_t)p->n_vars);
    combine_isc(p, s, out);
}

/* oop_score_requester.rs:443-463 -> incremental_score_calculator.rs:60-99 */
int gjo_score_incremental(const gjo_problem* p, const double* base, const uint64_t* offsets,
                          const uint64_t* var_ids, const double* values, int64_t S,
                          double* out) {
    scratch_t s;
    if (scratch_init(&s, p) != 0) return -1;
    int levels = gjo_levels(p->kind);
    int64_t* base_dec = (int64_t*)malloc(sizeof(int64_t) * (size_t)(p->n_vars + 1));
    int64_t* work = (int64_t*)malloc(sizeof(int64_t) * (size_t)(p->n_vars + 1));
    for (int i = 0; i < p->n_vars; ++i) base_dec[i] = decode_var(p, i, base[i]);
    for (int64_t j = 0; j < S; ++j) {
        uint64_t b = offsets[j], e = offsets[j + 1];
        score_one_incremental(p, &s, base_dec, work, var_ids + b, values + b,
                              (int64_t)(e - b), out + j * levels);
    }
    free(base_dec); free(work);
    scratch_free(&s);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* instance helpers                                                          */
/* ------------------------------------------------------------------------- */

/* examples/tsp/src/domain/location.rs:38-50 (+ domain_builder.rs:42-46: the
   second round() is idempotent).  powf(x, 2.0) == x*x. */
void gjo_distance_matrix(const double* xy, int n, double* D) {
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            double dlat = xy[2 * j] - xy[2 * i];
            double dlon = xy[2 * j + 1] - xy[2 * i + 1];
            double a = dlat * dlat;
            double b = dlon * dlon;
            double d = sqrt(a + b);
            d = gjo_round(d, 3);
            D[(size_t)i * (size_t)n + (size_t)j] = gjo_round(d, 3);
        }
    }
}

/* examples/tsp/src/persistence/cotwin_builder.rs:87-117 */
void gjo_tsp_greedy_init(const double* D, int n_locations, double* out_vars) {
    int n_stops = n_locations - 1;
    unsigned char* used = (unsigned char*)calloc((size_t)n_locations, 1);
    int prev = 0;
    for (int i = 0; i < n_stops; ++i) {
        double best = 1.7976931348623157e308;
        int best_id = -1;
        for (int c = 1; c < n_locations; ++c) {
            if (used[c]) continue;
            double d = D[(size_t)prev * (size_t)n_locations + (size_t)c];
            if (d < best) { best = d; best_id = c; }
        }
        used[best_id] = 1;
        out_vars[i] = (double)best_id;
        prev = best_id;
    }
    free(used);
}

/* examples/vrp/src/persistence/cotwin_builder.rs:153-255 */
void gjo_vrp_greedy_init(const gjo_problem* p, int n_depots, double* out_vars) {
    int L = p->n_locations;
    int n_stops = L - n_depots;
    unsigned char* used = (unsigned char*)calloc((size_t)L, 1);
    int remaining = n_stops, filled = 0;
    for (int k = 0; k < p->n_vehicles; ++k) {
        if (remaining <= 0) break;
        int depot = (int)p->vehicle_depot[k];
        uint64_t capacity = p->vehicle_capacity[k], collected = 0;
        int prev = depot;
        while (collected < capacity && remaining > 0) {
            double best = 1.7976931348623157e308;
            int best_id = -1;
            for (int c = n_depots; c < L; ++c) {
                if (used[c]) continue;
                double d = p->distance_matrix[(size_t)prev * (size_t)L + (size_t)c];
                if (d < best) { best = d; best_id = c; }
            }
            uint64_t dem = p->demand[best_id];
            if (collected + dem <= capacity) {
                collected += dem;
                used[best_id] = 1;
                --remaining;
                out_vars[2 * filled] = (double)k;
                out_vars[2 * filled + 1] = (double)best_id;
                ++filled;
                prev = best_id;
            } else {
                break;
            }
        }
    }
    for (; filled < n_stops; ++filled) { /* None -> left for random sampling */
        out_vars[2 * filled] = -1.0;
        out_vars[2 * filled + 1] = -1.0;
    }
    free(used);
}

/* ------------------------------------------------------------------------- */
/* mover (greyjack/src/agents/metaheuristic_bases/mover.rs)                  */
/* ------------------------------------------------------------------------- */

static void swapd(double* a, double* b) { double t = *a; *a = *b; *b = t; }

/* mover.rs:145-178 */
int gjo_move_change(const double* cand, int n_vars, const int32_t* group_ids, int group_len,
                    const int32_t* chosen, int k, const double* new_values, int incremental,
                    int32_t* out_cols, double* out_vals, double* out_candidate) {
    if (k < 1) k = 1;
    if (group_len < k) return -1;
    for (int i = 0; i < k; ++i) out_cols[i] = group_ids[chosen[i]];
    if (incremental) {
        for (int i = 0; i < k; ++i) out_vals[i] = new_values[i];
    } else {
        memcpy(out_candidate, cand, sizeof(double) * (size_t)n_vars);
        for (int i = 0; i < k; ++i) out_candidate[out_cols[i]] = new_values[i];
    }
    return k;
}

/* mover.rs:180-219 */
int gjo_move_swap(const double* cand, int n_vars, const int32_t* group_ids, int group_len,
                  const int32_t* chosen, int k, int incremental, int32_t* out_cols,
                  double* out_vals, double* out_candidate) {
    if (k < 2) k = 2;
    if (group_len < k) return -1;
    for (int i = 0; i < k; ++i) out_cols[i] = group_ids[chosen[i]];
    if (incremental) {
        for (int i = 0; i < k; ++i) out_vals[i] = cand[out_cols[i]];
        for (int i = 1; i < k; ++i) swapd(&out_vals[i - 1], &out_vals[i]);
    } else {
        memcpy(out_candidate, cand, sizeof(double) * (size_t)n_vars);
        for (int i = 1; i < k; ++i) swapd(&out_candidate[out_cols[i - 1]], &out_candidate[out_cols[i]]);
    }
    return k;
}

/* mover.rs:221-277 */
int gjo_move_swap_edges(const double* cand, int n_vars, const int32_t* group_ids, int group_len,
                        const int32_t* chosen, int k, int incremental, int32_t* out_cols,
                        double* out_vals, double* out_candidate) {
    if (group_len == 0) return -1;
    if (k < 2) k = 2;
    if (k > group_len - 1) k = group_len - 1;
    if (k <= 0) return -1; /* reference would panic in choice(); nothing to do */
    int32_t* e0 = (int32_t*)malloc(sizeof(int32_t) * (size_t)k * 2);
    int32_t* e1 = e0 + k;
    for (int i = 0; i < k; ++i) {
        e0[i] = group_ids[chosen[i]];
        e1[i] = group_ids[chosen[i] + 1];
        out_cols[2 * i] = e0[i];
        out_cols[2 * i + 1] = e1[i];
    }
    /* edges.rotate_left(1) */
    int32_t f0 = e0[0], f1 = e1[0];
    for (int i = 0; i + 1 < k; ++i) { e0[i] = e0[i + 1]; e1[i] = e1[i + 1]; }
    e0[k - 1] = f0; e1[k - 1] = f1;
    if (incremental) {
        for (int i = 0; i < k; ++i) {
            out_vals[2 * i] = cand[e0[i]];
            out_vals[2 * i + 1] = cand[e1[i]];
        }
        for (int i = 1; i < k; ++i) {
            swapd(&out_vals[2 * (i - 1)], &out_vals[2 * i]);
            swapd(&out_vals[2 * (i - 1) + 1], &out_vals[2 * i + 1]);
        }
    } else {
        memcpy(out_candidate, cand, sizeof(double) * (size_t)n_vars);
        for (int i = 1; i < k; ++i) {
            swapd(&out_candidate[e0[i - 1]], &out_candidate[e0[i]]);
            swapd(&out_candidate[e1[i - 1]], &out_candidate[e1[i]]);
        }
    }
    free(e0);
    return 2 * k;
}

/* mover.rs:279-317.  perm = the shuffle applied to native_columns. */
int gjo_move_scramble(const double* cand, int n_vars, const int32_t* group_ids, int group_len,
                      int start, int count, const int32_t* perm, int incremental,
                      int32_t* out_cols, double* out_vals, double* out_candidate) {
    if (group_len < count - 1) return -1;
    if (start + count > group_len) return -1;
    int32_t native[8], scrambled[8];
    for (int i = 0; i < count; ++i) native[i] = group_ids[start + i];
    for (int i = 0; i < count; ++i) scrambled[i] = native[perm[i]];
    if (incremental) {
        for (int i = 0; i < count; ++i) {
            out_cols[i] = scrambled[i];
            out_vals[i] = cand[scrambled[i]];
        }
    } else {
        memcpy(out_candidate, cand, sizeof(double) * (size_t)n_vars);
        for (int i = 0; i < count; ++i) {
            out_cols[i] = native[i];
            swapd(&out_candidate[native[i]], &out_candidate[scrambled[i]]);
        }
    }
    return count;
}

/* mover.rs:319-376 */
int gjo_move_insertion(const double* cand, int n_vars, const int32_t* group_ids, int group_len,
                       int get_out, int put_in, int incremental, int32_t* out_cols,
                       double* out_vals, double* out_candidate) {
    if (group_len <= 1) return -1;
    if (get_out == put_in) return -1;
    int lo = get_out < put_in ? get_out : put_in;
    int hi = get_out < put_in ? put_in : get_out;
    int left_rotate = get_out < put_in;
    int m = hi - lo + 1;
    for (int i = 0; i < m; ++i) out_cols[i] = group_ids[lo + i];
    if (incremental) {
        for (int i = 0; i < m; ++i) {
            int src = left_rotate ? (i + 1) % m : (i + m - 1) % m;
            out_vals[i] = cand[out_cols[src]];
        }
    } else {
        memcpy(out_candidate, cand, sizeof(double) * (size_t)n_vars);
        for (int i = 0; i < m; ++i) {
            int si = left_rotate ? (i + 1) % m : (i + m - 1) % m; /* shifted_ids[i] */
            swapd(&out_candidate[out_cols[i]], &out_candidate[out_cols[si]]);
        }
    }
    return m;
}

/* mover.rs:378-420 */
int gjo_move_inverse(const double* cand, int n_vars, const int32_t* group_ids, int group_len,
                     int a, int b, int incremental, int32_t* out_cols, double* out_vals,
                     double* out_candidate) {
    if (group_len <= 1) return -1;
    if (b < a) { int t = a; a = b; b = t; }
    int m = b - a + 1;
    for (int i = 0; i < m; ++i) out_cols[i] = group_ids[a + i];
    if (incremental) {
        for (int i = 0; i < m; ++i) out_vals[i] = cand[out_cols[m - 1 - i]];
    } else {
        memcpy(out_candidate, cand, sizeof(double) * (size_t)n_vars);
        for (int i = 0; i < m; ++i) out_candidate[out_cols[i]] = cand[out_cols[m - 1 - i]];
    }
    return m;
}

/* variables_manager.rs:203-220 */
void gjo_fix_deltas(const gjo_problem* p, const int32_t* cols, double* vals, int k) {
    for (int d = 0; d < k; ++d) {
        int v = cols[d];
        vals[d] = gjo_fix_integer(vals[d], p->lower_bounds[v], p->upper_bounds[v],
                                  p->frozen ? p->frozen[v] : 0, p->initial ? p->initial[v] : 0.0);
    }
}

/* variables_manager.rs:187-201 */
void gjo_fix_variables(const gjo_problem* p, double* candidate, const int32_t* cols, int k) {
    for (int d = 0; d < k; ++d) {
        int v = cols[d];
        candidate[v] = gjo_fix_integer(candidate[v], p->lower_bounds[v], p->upper_bounds[v],
                                       p->frozen ? p->frozen[v] : 0,
                                       p->initial ? p->initial[v] : 0.0);
    }
}

/* ------------------------------------------------------------------------- */
/* selection rules                                                           */
/* ------------------------------------------------------------------------- */

/* tabu_search_base.rs:157-188: min_by(cmp) keeps the FIRST minimum. */
int64_t gjo_ts_select(const double* scores, int64_t S, int levels, const double* current,
                      int* accept) {
    int64_t best = 0;
    for (int64_t j = 1; j < S; ++j)
        if (gjo_score_cmp(scores + j * levels, scores + best * levels, levels) < 0) best = j;
    *accept = gjo_score_le(scores + best * levels, current, levels);
    return best;
}

/* late_acceptance_base.rs:188-241.  late[0..late_len) with late[0] the front. */
int gjo_la_accept(const double* cand, const double* current, double* late, int* late_len,
                  int late_size, int levels) {
    const double* late_native = (*late_len == 0) ? current : late + (size_t)(*late_len - 1) * levels;
    if (gjo_score_le(cand, late_native, levels) || gjo_score_le(cand, current, levels)) {
        /* push_front */
        memmove(late + levels, late, sizeof(double) * (size_t)levels * (size_t)(*late_len));
        memcpy(late, cand, sizeof(double) * (size_t)levels);
        ++*late_len;
        if (*late_len > late_size) --*late_len; /* pop_back */
        return 1;
    }
    return 0;
}

/* genetic_algorithm_base.rs:198-213 */
int gjo_sa_accept(const double* cand, const double* current, int levels, double* temperature,
                  int has_cooling_rate, double cooling_rate, double inverted_accomplish_rate,
                  double u, double* proba_out) {
    /* simulated_annealing_base.rs:206-216 */
    for (int l = 0; l < levels; ++l) {
        if (has_cooling_rate) {
            double t = temperature[l] * cooling_rate;
            if (t < 0.000001) t = 0.0000001;
            temperature[l] = t;
        } else {
            temperature[l] = inverted_accomplish_rate;
        }
    }
    /* :218-225: self.exp.powf(-((can_e - cur_e) / T)), folded from 1.0 */
    double proba = 1.0;
    for (int l = 0; l < levels; ++l)
        proba = proba * pow(2.7182818284590452, -((cand[l] - current[l]) / temperature[l]));
    if (proba_out) *proba_out = proba;
    return u < proba ? 1 : 0;                       /* :229 */
}

void gjo_ga_replace(const double* cand_scores, const double* pop_scores, const int64_t* worst_ids,
                    int64_t pop, int levels, int64_t* out_src) {
    for (int64_t i = 0; i < pop; ++i) {
        const double* c = cand_scores + i * levels;
        const double* w = pop_scores + worst_ids[i] * levels;
        out_src[i] = gjo_score_le(c, w, levels) ? i : -(worst_ids[i] + 1);
    }
}

/* genetic_algorithm_base.rs:83-103: select_p_best / select_p_worst with the two random draws made
   explicit.  p_best_proba ~ U(1e-6, p_best_rate); last_top_id = ceil(p * pop); chosen_id =
   U[0, last_top_id) (best) | U[pop - last_top_id, pop) (worst) -- `id_draw` is the offset inside that
   range.  Returns the chosen index into the SORTED population, or -1 when the draw is out of range. */
int64_t gjo_ga_select(double p_best_proba, int64_t id_draw, int64_t pop, int worst, int64_t* last_top_out) {
    int64_t last_top_id = (int64_t)ceil(p_best_proba * (double)pop);
    if (last_top_out) *last_top_out = last_top_id;
    if (id_draw < 0 || id_draw >= last_top_id || last_top_id > pop) return -1;
    return worst ? (pop - last_top_id + id_draw) : id_draw;
}

/* genetic_algorithm_base.rs:105-134: cross with ONE weight for every gene (`vec![sample(); n]` draws
   once), rint-ed on discrete columns; children are convex sums of the parents.  discrete: [n] 0/1
   or NULL (= all discrete, every GJInteger problem). */
void gjo_ga_cross(const double* c1, const double* c2, int n, double weight, const uint8_t* discrete,
                  double* out1, double* out2) {
    for (int i = 0; i < n; ++i) {
        double w = weight;
        if (!discrete || discrete[i]) w = gjo_rint(w);
        out1[i] = c1[i] * w + c2[i] * (1.0 - w);
        out2[i] = c2[i] * w + c1[i] * (1.0 - w);
    }
}

/* ------------------------------------------------------------------------- */
/* CPU baseline drivers                                                      */
/* ------------------------------------------------------------------------- */

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double rnd01(uint64_t* s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }
static int rnd_below(uint64_t* s, int n) { return (int)(splitmix64(s) % (uint64_t)n); }

typedef struct {
    const gjo_problem* p;
    const double* base;
    int n_moves, n_steps, tid;
    uint64_t seed;
    const double* move_probas;
    const int64_t* precision;
    int64_t scored;
    double best[3];
} ts_job;

/* One island: Agent::step_incremental (agent_base.rs:300-320) in a loop with
   TabuSearchBase (tabu_search_base.rs:107-188), tabu_entity_rate = 0 and
   mutation_rate_multiplier = None (minimal move sizes).  The reference's O(N)
   RNG draws per move (mover.rs:138) and Polars marshalling are NOT reproduced:
   this baseline is faster than the real reference (CPU-favouring).            */
static void* ts_worker(void* arg) {
    ts_job* job = (ts_job*)arg;
    const gjo_problem* p = job->p;
    int n = p->n_vars, levels = gjo_levels(p->kind);
    scratch_t s;
    scratch_init(&s, p);
    uint64_t rng = job->seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(job->tid + 1));
    double* cur = (double*)malloc(sizeof(double) * (size_t)n);
    memcpy(cur, job->base, sizeof(double) * (size_t)n);
    int64_t* base_dec = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int64_t* work = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int32_t* group = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; ++i) group[i] = i;
    /* per-move delta storage: CSR */
    size_t cap = (size_t)job->n_moves * 8 + (size_t)n * 4;
    uint64_t* offs = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(job->n_moves + 1));
    uint64_t* ids = (uint64_t*)malloc(sizeof(uint64_t) * cap);
    double* vals = (double*)malloc(sizeof(double) * cap);
    int32_t* cols32 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 8));
    double* dv = (double*)malloc(sizeof(double) * (size_t)(n + 8));
    double* scores = (double*)malloc(sizeof(double) * (size_t)job->n_moves * (size_t)levels);
    double cur_score[3] = {0, 0, 0};
    int64_t prec[3] = {-1, -1, -1}; /* Solver::solve score_precision; <0 = None */
    if (job->precision) for (int l = 0; l < levels; ++l) prec[l] = job->precision[l];
    {
        for (int i = 0; i < n; ++i) base_dec[i] = decode_var(p, i, cur[i]);
        score_one_incremental(p, &s, base_dec, work, NULL, NULL, 0, cur_score);
    }
    double thr[6], acc = 0.0;
    for (int m = 0; m < 6; ++m) { acc += job->move_probas[m]; thr[m] = acc; }
    int64_t scored = 0;
    for (int step = 0; step < job->n_steps; ++step) {
        size_t fill = 0;
        for (int j = 0; j < job->n_moves; ++j) {
            offs[j] = fill;
            if (fill + (size_t)n + 8 > cap) {
                cap = cap * 2 + (size_t)n;
                ids = (uint64_t*)realloc(ids, sizeof(uint64_t) * cap);
                vals = (double*)realloc(vals, sizeof(double) * cap);
            }
            double u = rnd01(&rng);
            int mv = 5;
            for (int m = 0; m < 6; ++m) if (u <= thr[m]) { mv = m; break; }
            int k = -1;
            int a = rnd_below(&rng, n), b = rnd_below(&rng, n - 1);
            if (b >= a) ++b;
            int32_t chosen[2] = {a, b};
            switch (mv) {
                case GJO_MOVE_CHANGE: {
                    double nv = p->lower_bounds[group[a]] +
                                rnd01(&rng) * (p->upper_bounds[group[a]] - p->lower_bounds[group[a]]);
                    k = gjo_move_change(cur, n, group, n, chosen, 1, &nv, 1, cols32, dv, NULL);
                } break;
                case GJO_MOVE_SWAP:
                    k = gjo_move_swap(cur, n, group, n, chosen, 2, 1, cols32, dv, NULL); break;
                case GJO_MOVE_SWAP_EDGES: {
                    int32_t ce[2] = {rnd_below(&rng, n - 1), 0};
                    ce[1] = rnd_below(&rng, n - 2); if (ce[1] >= ce[0]) ++ce[1];
                    k = gjo_move_swap_edges(cur, n, group, n, ce, 2, 1, cols32, dv, NULL);
                } break;
                case GJO_MOVE_SCRAMBLE: {
                    int count = 3 + rnd_below(&rng, 4);
                    int32_t perm[6] = {0, 1, 2, 3, 4, 5};
                    for (int i = count - 1; i > 0; --i) {
                        int r = rnd_below(&rng, i + 1);
                        int32_t t = perm[i]; perm[i] = perm[r]; perm[r] = t;
                    }
                    int start = rnd_below(&rng, n - count);
                    k = gjo_move_scramble(cur, n, group, n, start, count, perm, 1, cols32, dv, NULL);
                } break;
                case GJO_MOVE_INSERTION:
                    k = gjo_move_insertion(cur, n, group, n, a, b, 1, cols32, dv, NULL); break;
                default:
                    k = gjo_move_inverse(cur, n, group, n, a, b, 1, cols32, dv, NULL); break;
            }
            if (k < 0) k = 0;
            gjo_fix_deltas(p, cols32, dv, k);
            for (int d = 0; d < k; ++d) { ids[fill + (size_t)d] = (uint64_t)cols32[d]; vals[fill + (size_t)d] = dv[d]; }
            fill += (size_t)k;
        }
        offs[job->n_moves] = fill;
        /* request_score_incremental */
        for (int i = 0; i < n; ++i) base_dec[i] = decode_var(p, i, cur[i]);
        for (int j = 0; j < job->n_moves; ++j) {
            score_one_incremental(p, &s, base_dec, work, ids + offs[j], vals + offs[j],
                                  (int64_t)(offs[j + 1] - offs[j]), scores + (size_t)j * levels);
            gjo_score_round(scores + (size_t)j * levels, prec, levels); /* agent_base.rs:311-314 */
        }
        scored += job->n_moves;
        int accept = 0;
        int64_t best = gjo_ts_select(scores, job->n_moves, levels, cur_score, &accept);
        if (accept) {
            for (uint64_t d = offs[best]; d < offs[best + 1]; ++d) cur[ids[d]] = vals[d];
            memcpy(cur_score, scores + (size_t)best * levels, sizeof(double) * (size_t)levels);
        }
    }
    job->scored = scored;
    memcpy(job->best, cur_score, sizeof(double) * 3);
    free(cur); free(base_dec); free(work); free(group); free(offs); free(ids); free(vals);
    free(cols32); free(dv); free(scores);
    scratch_free(&s);
    return NULL;
}

int64_t gjo_bench_ts(const gjo_problem* p, const double* base, int n_moves, int n_steps,
                     int n_threads, uint64_t seed, const double* move_probas,
                     const int64_t* precision, double* seconds, double* best_out) {
    if (n_threads < 1) n_threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    ts_job* jobs = (ts_job*)calloc((size_t)n_threads, sizeof(ts_job));
    double t0 = now_s();
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].p = p; jobs[t].base = base; jobs[t].n_moves = n_moves; jobs[t].n_steps = n_steps;
        jobs[t].tid = t; jobs[t].seed = seed; jobs[t].move_probas = move_probas;
        jobs[t].precision = precision;
        pthread_create(&th[t], NULL, ts_worker, &jobs[t]);
    }
    int64_t total = 0;
    int levels = gjo_levels(p->kind);
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        total += jobs[t].scored;
        if (best_out && (t == 0 || gjo_score_cmp(jobs[t].best, best_out, levels) < 0))
            memcpy(best_out, jobs[t].best, sizeof(double) * (size_t)levels);
    }
    *seconds = now_s() - t0;
    free(th); free(jobs);
    return total;
}

typedef struct {
    const gjo_problem* p;
    const double* samples;
    int64_t lo, hi;
    int repeats;
    double* out;
} plain_job;

static void* plain_worker(void* arg) {
    plain_job* job = (plain_job*)arg;
    int levels = gjo_levels(job->p->kind);
    for (int r = 0; r < job->repeats; ++r)
        gjo_score_plain(job->p, job->samples + job->lo * (int64_t)job->p->n_vars,
                        job->hi - job->lo, job->out + job->lo * levels);
    return NULL;
}

int64_t gjo_bench_plain(const gjo_problem* p, const double* samples, int64_t S, int n_threads,
                        int repeats, double* seconds, double* out) {
    if (n_threads < 1) n_threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    plain_job* jobs = (plain_job*)calloc((size_t)n_threads, sizeof(plain_job));
    double t0 = now_s();
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].p = p; jobs[t].samples = samples; jobs[t].repeats = repeats; jobs[t].out = out;
        jobs[t].lo = S * t / n_threads; jobs[t].hi = S * (t + 1) / n_threads;
        pthread_create(&th[t], NULL, plain_worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    *seconds = now_s() - t0;
    free(th); free(jobs);
    return S * repeats;
}

/* ------------------------------------------------------------------------- */
/* LateAcceptance and GeneticAlgorithm baseline drivers (BASELINE configs 1, 3, 4) */
/* ------------------------------------------------------------------------- */

typedef struct {
    const gjo_problem* p;
    const double* base;
    const int64_t* group_offsets; const int32_t* group_ids; int n_groups;
    int late_size, n_steps, tid;
    uint64_t seed;
    const double* move_probas;
    const int64_t* precision;
    int64_t scored;
    double best[3];
} la_job;

/* One LateAcceptance agent: Agent::step_incremental (agent_base.rs:300-320) with
   LateAcceptanceBase (late_acceptance_base.rs:116-241): ONE neighbour per step -- a move of minimal
   size on a random semantic group -- scored by the pseudo-incremental ISC (full re-evaluation of
   base + deltas), accepted against the late list.  tabu_entity_rate = 0 and no Polars
   marshalling: faster than the real reference (CPU-favouring).  Only change / swap / insertion /
   inverse are drawn (what the examples' LateAcceptance configurations use).                      */
static void* la_worker(void* arg) {
    la_job* job = (la_job*)arg;
    const gjo_problem* p = job->p;
    const int n = p->n_vars, levels = gjo_levels(p->kind);
    scratch_t s;
    scratch_init(&s, p);
    uint64_t rng = job->seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(job->tid + 1));
    double* cur = (double*)malloc(sizeof(double) * (size_t)n);
    memcpy(cur, job->base, sizeof(double) * (size_t)n);
    int64_t* base_dec = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int64_t* work = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int32_t* cols32 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 8));
    double* dv = (double*)malloc(sizeof(double) * (size_t)(n + 8));
    uint64_t* ids = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n + 8));
    double* late = (double*)calloc((size_t)(job->late_size + 2) * 3, sizeof(double));
    int late_len = 0;
    double cur_score[3] = {0, 0, 0}, top[3], sc[3];
    int64_t prec[3] = {-1, -1, -1};
    if (job->precision) for (int l = 0; l < levels; ++l) prec[l] = job->precision[l];
    for (int i = 0; i < n; ++i) base_dec[i] = decode_var(p, i, cur[i]);
    score_one_incremental(p, &s, base_dec, work, NULL, NULL, 0, cur_score);
    memcpy(top, cur_score, sizeof(top));
    double thr[6], acc = 0.0;
    for (int m = 0; m < 6; ++m) { acc += job->move_probas[m]; thr[m] = acc; }
    for (int step = 0; step < job->n_steps; ++step) {
        const int gi = rnd_below(&rng, job->n_groups);
        const int32_t* group = job->group_ids + job->group_offsets[gi];
        const int glen = (int)(job->group_offsets[gi + 1] - job->group_offsets[gi]);
        const double u = rnd01(&rng);
        int mv = 5;
        for (int m = 0; m < 6; ++m) if (u <= thr[m]) { mv = m; break; }
        int a = rnd_below(&rng, glen), b = rnd_below(&rng, glen - 1);
        if (b >= a) ++b;
        int32_t chosen[2] = {a, b};
        int k;
        if (mv == GJO_MOVE_CHANGE) {
            double nv = p->lower_bounds[group[a]] + rnd01(&rng) * (p->upper_bounds[group[a]] - p->lower_bounds[group[a]]);
            k = gjo_move_change(cur, n, group, glen, chosen, 1, &nv, 1, cols32, dv, NULL);
        } else if (mv == GJO_MOVE_INSERTION) {
            k = gjo_move_insertion(cur, n, group, glen, a, b, 1, cols32, dv, NULL);
        } else if (mv == GJO_MOVE_INVERSE) {
            k = gjo_move_inverse(cur, n, group, glen, a, b, 1, cols32, dv, NULL);
        } else {
            k = gjo_move_swap(cur, n, group, glen, chosen, 2, 1, cols32, dv, NULL);
        }
        if (k < 0) k = 0;
        gjo_fix_deltas(p, cols32, dv, k);
        for (int d = 0; d < k; ++d) ids[d] = (uint64_t)cols32[d];
        score_one_incremental(p, &s, base_dec, work, ids, dv, k, sc);
        gjo_score_round(sc, prec, levels);                       /* agent_base.rs:311-314 */
        if (gjo_la_accept(sc, cur_score, late, &late_len, job->late_size, levels)) {
            for (int d = 0; d < k; ++d) { cur[cols32[d]] = dv[d]; base_dec[cols32[d]] = decode_var(p, cols32[d], dv[d]); }
            memcpy(cur_score, sc, sizeof(double) * (size_t)levels);
            if (gjo_score_le(cur_score, top, levels)) memcpy(top, cur_score, sizeof(top));   /* agent_base.rs:220-224 */
        }
    }
    job->scored = job->n_steps;
    memcpy(job->best, top, sizeof(top));
    free(cur); free(base_dec); free(work); free(cols32); free(dv); free(ids); free(late);
    scratch_free(&s);
    return NULL;
}

int64_t gjo_bench_la(const gjo_problem* p, const double* base, const int64_t* group_offsets,
                     const int32_t* group_ids, int n_groups, int late_size, int n_steps, int n_threads,
                     uint64_t seed, const double* move_probas, const int64_t* precision,
                     double* seconds, double* best_out) {
    if (n_threads < 1) n_threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    la_job* jobs = (la_job*)calloc((size_t)n_threads, sizeof(la_job));
    double t0 = now_s();
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].p = p; jobs[t].base = base; jobs[t].group_offsets = group_offsets; jobs[t].group_ids = group_ids;
        jobs[t].n_groups = n_groups; jobs[t].late_size = late_size; jobs[t].n_steps = n_steps; jobs[t].tid = t;
        jobs[t].seed = seed; jobs[t].move_probas = move_probas; jobs[t].precision = precision;
        pthread_create(&th[t], NULL, la_worker, &jobs[t]);
    }
    int64_t total = 0;
    const int levels = gjo_levels(p->kind);
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        total += jobs[t].scored;
        if (best_out && (t == 0 || gjo_score_cmp(jobs[t].best, best_out, levels) < 0))
            memcpy(best_out, jobs[t].best, sizeof(double) * (size_t)levels);
    }
    *seconds = now_s() - t0;
    free(th); free(jobs);
    return total;
}

typedef struct {
    const gjo_problem* p;
    const int64_t* group_offsets; const int32_t* group_ids; int n_groups;
    int pop, n_generations, tid;
    double crossover_probability, p_best_rate;
    uint64_t seed;
    const double* move_probas;
    const int64_t* precision;
    int64_t scored;
    double best[3];
} ga_job;

static int g_ga_levels;         /* qsort comparator context (one value for every worker of a run) */
static const double* g_ga_dummy;
typedef struct { double sc[3]; int idx; } ga_key;
static int ga_key_cmp(const void* a, const void* b) {
    const ga_key* x = (const ga_key*)a; const ga_key* y = (const ga_key*)b;
    int c = gjo_score_cmp(x->sc, y->sc, g_ga_levels);
    if (c) return c;
    return (x->idx > y->idx) - (x->idx < y->idx);      /* Vec::sort is stable */
}

/* One GeneticAlgorithm agent: Agent::step_plain (agent_base.rs:273-298) with GeneticAlgorithmBase
   (genetic_algorithm_base.rs:141-213): population.sort(); per pair two p-best parents, crossover
   (one rint-ed weight: the parents swap or stay), one plain-form move each, fix_variables;
   request_score_plain on the offspring (PSC), round; slot i = candidate i if <= a random p-worst
   native.  Initial population = uniform samples (GJInteger::sample, initial None).               */
static void* ga_worker(void* arg) {
    ga_job* job = (ga_job*)arg;
    const gjo_problem* p = job->p;
    const int n = p->n_vars, levels = gjo_levels(p->kind), pop = job->pop;
    const int half = (pop + 1) / 2, nc = 2 * half;
    uint64_t rng = job->seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(job->tid + 1));
    double* rows = (double*)malloc(sizeof(double) * (size_t)pop * (size_t)n);
    double* next = (double*)malloc(sizeof(double) * (size_t)pop * (size_t)n);
    double* cand = (double*)malloc(sizeof(double) * (size_t)nc * (size_t)n);
    double* sc = (double*)malloc(sizeof(double) * (size_t)pop * 3);
    double* sc_next = (double*)malloc(sizeof(double) * (size_t)pop * 3);
    double* csc = (double*)malloc(sizeof(double) * (size_t)nc * 3);
    double* tmp = (double*)malloc(sizeof(double) * (size_t)n);
    ga_key* keys = (ga_key*)malloc(sizeof(ga_key) * (size_t)pop);
    int32_t* cols32 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(2 * n + 16));
    double* dv = (double*)malloc(sizeof(double) * (size_t)(2 * n + 16));
    int64_t prec[3] = {-1, -1, -1};
    if (job->precision) for (int l = 0; l < levels; ++l) prec[l] = job->precision[l];
    for (int k = 0; k < pop; ++k)
        for (int i = 0; i < n; ++i) {
            const int64_t lo = (int64_t)p->lower_bounds[i], hi = (int64_t)p->upper_bounds[i];
            rows[(size_t)k * n + i] = (double)(lo + (int64_t)(splitmix64(&rng) % (uint64_t)(hi - lo + 1)));
        }
    {
        double* flat = (double*)malloc(sizeof(double) * (size_t)pop * (size_t)levels);
        gjo_score_plain(p, rows, pop, flat);
        for (int k = 0; k < pop; ++k) for (int l = 0; l < levels; ++l) sc[(size_t)k * 3 + l] = flat[(size_t)k * levels + l];
        free(flat);
    }
    double top[3] = {1.7976931348623157e308, 0, 0};
    double thr[6], acc = 0.0;
    for (int m = 0; m < 6; ++m) { acc += job->move_probas[m]; thr[m] = acc; }
    double* flat = (double*)malloc(sizeof(double) * (size_t)nc * (size_t)levels);
    int64_t scored = 0;
    for (int gen = 0; gen < job->n_generations; ++gen) {
        for (int k = 0; k < pop; ++k) { memcpy(keys[k].sc, sc + (size_t)k * 3, sizeof(double) * 3); keys[k].idx = k; }
        qsort(keys, (size_t)pop, sizeof(ga_key), ga_key_cmp);              /* population.sort() */
        if (gjo_score_le(keys[0].sc, top, levels)) memcpy(top, keys[0].sc, sizeof(top));
        for (int q = 0; q < half; ++q) {
            int64_t r[2];
            for (int h = 0; h < 2; ++h) {
                const double pb = 0.000001 + rnd01(&rng) * (job->p_best_rate - 0.000001);
                int64_t last_top = (int64_t)ceil(pb * (double)pop);
                if (last_top < 1) last_top = 1;
                r[h] = gjo_ga_select(pb, rnd_below(&rng, (int)last_top), pop, 0, NULL);
            }
            const double* p1 = rows + (size_t)keys[r[0]].idx * n;
            const double* p2 = rows + (size_t)keys[r[1]].idx * n;
            double* c1 = cand + (size_t)(2 * q) * n;
            double* c2 = cand + (size_t)(2 * q + 1) * n;
            if (rnd01(&rng) <= job->crossover_probability) gjo_ga_cross(p1, p2, n, rnd01(&rng), NULL, c1, c2);
            else { memcpy(c1, p1, sizeof(double) * (size_t)n); memcpy(c2, p2, sizeof(double) * (size_t)n); }
            for (int h = 0; h < 2; ++h) {
                double* c = h ? c2 : c1;
                const int gi = rnd_below(&rng, job->n_groups);
                const int32_t* group = job->group_ids + job->group_offsets[gi];
                const int glen = (int)(job->group_offsets[gi + 1] - job->group_offsets[gi]);
                const double u = rnd01(&rng);
                int mv = 5;
                for (int m = 0; m < 6; ++m) if (u <= thr[m]) { mv = m; break; }
                int a = rnd_below(&rng, glen), b = rnd_below(&rng, glen - 1);
                if (b >= a) ++b;
                int32_t chosen[2] = {a, b};
                int k = -1;
                switch (mv) {
                    case GJO_MOVE_CHANGE: {
                        double nv = p->lower_bounds[group[a]] + rnd01(&rng) * (p->upper_bounds[group[a]] - p->lower_bounds[group[a]]);
                        k = gjo_move_change(c, n, group, glen, chosen, 1, &nv, 0, cols32, dv, tmp);
                    } break;
                    case GJO_MOVE_SWAP: k = gjo_move_swap(c, n, group, glen, chosen, 2, 0, cols32, dv, tmp); break;
                    case GJO_MOVE_SWAP_EDGES: {
                        int32_t ce[2] = {rnd_below(&rng, glen - 1), 0};
                        ce[1] = rnd_below(&rng, glen - 2); if (ce[1] >= ce[0]) ++ce[1];
                        k = gjo_move_swap_edges(c, n, group, glen, ce, 2, 0, cols32, dv, tmp);
                    } break;
                    case GJO_MOVE_SCRAMBLE: {
                        int count = 3 + rnd_below(&rng, 4);
                        int32_t perm[6] = {0, 1, 2, 3, 4, 5};
                        for (int i = count - 1; i > 0; --i) { int rr = rnd_below(&rng, i + 1); int32_t t = perm[i]; perm[i] = perm[rr]; perm[rr] = t; }
                        k = gjo_move_scramble(c, n, group, glen, rnd_below(&rng, glen - count), count, perm, 0, cols32, dv, tmp);
                    } break;
                    case GJO_MOVE_INSERTION: k = gjo_move_insertion(c, n, group, glen, a, b, 0, cols32, dv, tmp); break;
                    default: k = gjo_move_inverse(c, n, group, glen, a, b, 0, cols32, dv, tmp); break;
                }
                if (k >= 0) {
                    memcpy(c, tmp, sizeof(double) * (size_t)n);
                    gjo_fix_variables(p, c, cols32, k);
                }
            }
        }
        gjo_score_plain(p, cand, nc, flat);                                     /* request_score_plain */
        for (int k = 0; k < nc; ++k) {
            gjo_score_round(flat + (size_t)k * levels, prec, levels);           /* agent_base.rs:284-287 */
            for (int l = 0; l < 3; ++l) csc[(size_t)k * 3 + l] = l < levels ? flat[(size_t)k * levels + l] : 0.0;
        }
        scored += nc;
        for (int i = 0; i < pop; ++i) {                                         /* :198-213 */
            const double pb = 0.000001 + rnd01(&rng) * (job->p_best_rate - 0.000001);
            int64_t last_top = (int64_t)ceil(pb * (double)pop);
            if (last_top < 1) last_top = 1;
            const int64_t rk = gjo_ga_select(pb, rnd_below(&rng, (int)last_top), pop, 1, NULL);
            const int native = keys[rk].idx;
            if (gjo_score_le(csc + (size_t)i * 3, sc + (size_t)native * 3, levels)) {
                memcpy(next + (size_t)i * n, cand + (size_t)i * n, sizeof(double) * (size_t)n);
                memcpy(sc_next + (size_t)i * 3, csc + (size_t)i * 3, sizeof(double) * 3);
            } else {
                memcpy(next + (size_t)i * n, rows + (size_t)native * n, sizeof(double) * (size_t)n);
                memcpy(sc_next + (size_t)i * 3, sc + (size_t)native * 3, sizeof(double) * 3);
            }
        }
        { double* t = rows; rows = next; next = t; t = sc; sc = sc_next; sc_next = t; }
    }
    for (int k = 0; k < pop; ++k) if (gjo_score_le(sc + (size_t)k * 3, top, levels)) memcpy(top, sc + (size_t)k * 3, sizeof(top));
    job->scored = scored;
    memcpy(job->best, top, sizeof(top));
    free(rows); free(next); free(cand); free(sc); free(sc_next); free(csc); free(tmp); free(keys);
    free(cols32); free(dv); free(flat);
    return NULL;
}

int64_t gjo_bench_ga(const gjo_problem* p, const int64_t* group_offsets, const int32_t* group_ids,
                     int n_groups, int pop, double crossover_probability, double p_best_rate,
                     int n_generations, int n_threads, uint64_t seed, const double* move_probas,
                     const int64_t* precision, double* seconds, double* best_out) {
    if (n_threads < 1) n_threads = 1;
    g_ga_levels = gjo_levels(p->kind);
    (void)g_ga_dummy;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    ga_job* jobs = (ga_job*)calloc((size_t)n_threads, sizeof(ga_job));
    double t0 = now_s();
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].p = p; jobs[t].group_offsets = group_offsets; jobs[t].group_ids = group_ids; jobs[t].n_groups = n_groups;
        jobs[t].pop = pop; jobs[t].n_generations = n_generations; jobs[t].tid = t;
        jobs[t].crossover_probability = crossover_probability; jobs[t].p_best_rate = p_best_rate;
        jobs[t].seed = seed; jobs[t].move_probas = move_probas; jobs[t].precision = precision;
        pthread_create(&th[t], NULL, ga_worker, &jobs[t]);
    }
    int64_t total = 0;
    const int levels = gjo_levels(p->kind);
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        total += jobs[t].scored;
        if (best_out && (t == 0 || gjo_score_cmp(jobs[t].best, best_out, levels) < 0))
            memcpy(best_out, jobs[t].best, sizeof(double) * (size_t)levels);
    }
    *seconds = now_s() - t0;
    free(th); free(jobs);
    return total;
}
