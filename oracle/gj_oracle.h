/*
 * gj_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's candidate-scoring hot path
 * (CameleoGrey/greyjack-solver-rust, crate greyjack 0.4.13 + examples).  Every
 * function cites the reference file:line it follows (paths relative to the
 * reference root).  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may link or call this library; the CUDA
 * product path never does.
 *
 * PARITY STATUS: "parity unpinned" at the scorer boundary -- the reference ships
 * no test that pins any constraint score.  What the
 * reference's own unit tests DO pin (GJInteger clamp / inverse_transform, score
 * ordering, fitness values) is reproduced in tests/test_oracle_pinned.py.  The
 * reference cannot be compiled here (no cargo/rustc, polars 0.46.0 not vendored).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fPIC -shared -pthread (see Makefile).
 * -ffp-contract=off matters: Rust never fuses a*b+c, so neither may we.
 */
#ifndef GJ_ORACLE_H
#define GJ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { GJO_NQUEENS = 0, GJO_TSP = 1, GJO_VRP = 2, GJO_VRP_SERVICE = 3 };

/* Move ids in the order of Mover::do_move thresholds (mover.rs:105-121). */
enum { GJO_MOVE_CHANGE = 0, GJO_MOVE_SWAP = 1, GJO_MOVE_SWAP_EDGES = 2,
       GJO_MOVE_SCRAMBLE = 3, GJO_MOVE_INSERTION = 4, GJO_MOVE_INVERSE = 5 };

typedef struct {
    int32_t kind;
    int32_t n_vars;              /* planning variables in reference enumeration order
                                    (oop_score_requester.rs:93-123): entity by entity,
                                    field by field; VRP: [v0,c0,v1,c1,...]            */
    const double*  lower_bounds; /* [n_vars] GJInteger.lower_bound                     */
    const double*  upper_bounds; /* [n_vars]                                           */
    const uint8_t* frozen;       /* [n_vars] or NULL                                   */
    const double*  initial;      /* [n_vars] or NULL; used only for frozen variables   */

    /* N-Queens: column_id fact per queen (NULL -> column_id[i] = i,
       examples/nqueens/src/persistence/cotwin_builder.rs:58-75) */
    const int64_t* column_id;

    /* TSP / VRP utility objects */
    int32_t n_locations;
    const double* distance_matrix;   /* row-major [n_locations][n_locations] */

    int32_t n_vehicles;
    const int64_t*  vehicle_depot;   /* [n_vehicles] depot_vec_id   */
    const uint64_t* vehicle_capacity;
    const uint64_t* work_day_start;
    const uint64_t* work_day_end;
    const uint64_t* demand;          /* [n_locations] customers_info[c].demand */
    const uint64_t* tw_start;
    const uint64_t* tw_end;
    const uint64_t* service_time;
    int32_t time_windowed;

    /* constraint weights (score_calculators/plain_score_calculator.rs:79-90).
       PSC order: nqueens [all_different]; tsp [no_dup, distance];
       vrp [no_dup, capacity, distance, late_arrival].  ISC uses weights[0] for
       the single all_in_one constraint.                                        */
    double weights[4];
} gjo_problem;

/* ---- value helpers -------------------------------------------------------- */
double  gjo_rint(double x);                                  /* utils/math_utils.rs:6-8   */
double  gjo_round(double value, uint64_t precision);         /* utils/math_utils.rs:10-13 */
double  gjo_fix_integer(double value, double lb, double ub,
                        int frozen, double initial);         /* variables/gj_integer.rs:70-83 */
int64_t gjo_inverse_transform_integer(double value, double lb, double ub,
                                      int frozen, double initial); /* gj_integer.rs:66-68 */
double  gjo_fix_float(double value, double lb, double ub,
                      int frozen, double initial);           /* variables/gj_float.rs:64-76 */
int     gjo_levels(int kind);                                /* ScoreTrait::precision_len */

/* Ord::cmp (total_cmp, lexicographic; lower is better): scores/{simple,hard_soft,hard_medium_soft}_score.rs */
int     gjo_score_cmp(const double* a, const double* b, int levels);
/* derived PartialOrd `a <= b` used by every acceptance test                  */
int     gjo_score_le(const double* a, const double* b, int levels);
void    gjo_score_round(double* s, const int64_t* precision, int levels);
double  gjo_fitness(const double* s, int levels);            /* get_fitness_value */
void    gjo_sort_scores(double* scores, int n, int levels);  /* Vec<Score>::sort() */

/* ---- scorers --------------------------------------------------------------- */
/* request_score_plain -> PlainScoreCalculator::get_score (PSC semantics).
   samples: [S][n_vars] f64 row-major.  out: [S][levels].                      */
int gjo_score_plain(const gjo_problem* p, const double* samples, int64_t S, double* out);

/* request_score_incremental -> IncrementalScoreCalculator::get_score (ISC
   semantics, pseudo-incremental).  Deltas as CSR of Vec<Vec<(usize,f64)>>.    */
int gjo_score_incremental(const gjo_problem* p, const double* base,
                          const uint64_t* offsets, const uint64_t* var_ids,
                          const double* values, int64_t S, double* out);

/* Distance matrix of the examples: round(sqrt(dx^2+dy^2), 3)
   (examples/tsp/src/domain/location.rs:38-50).  xy: [n][2].                   */
void gjo_distance_matrix(const double* xy, int n, double* D);

/* Greedy nearest-neighbour TSP init (tsp/.../cotwin_builder.rs:87-117); the
   reference iterates a HashSet (arbitrary order) -- ties broken by lowest id. */
void gjo_tsp_greedy_init(const double* D, int n_locations, double* out_vars);
/* Greedy VRP init (vrp/.../cotwin_builder.rs:153-255); unassigned tail -> -1. */
void gjo_vrp_greedy_init(const gjo_problem* p, int n_depots, double* out_vars);

/* ---- mover (explicit random choices; reference RNG is entropy-seeded) ------ */
/* Each returns the number of (column,value) pairs written (incremental form) or
   n_changed columns (plain form; out_candidate gets the whole changed vector),
   or -1 when the reference returns (None,None,None).                           */
int gjo_move_change(const double* cand, int n_vars, const int32_t* group_ids, int group_len,
                    const int32_t* chosen, int k, const double* new_values, int incremental,
                    int32_t* out_cols, double* out_vals, double* out_candidate);
int gjo_move_swap(const double* cand, int n_vars, const int32_t* group_ids, int group_len,
                  const int32_t* chosen, int k, int incremental,
                  int32_t* out_cols, double* out_vals, double* out_candidate);
int gjo_move_swap_edges(const double* cand, int n_vars, const int32_t* group_ids, int group_len,
                        const int32_t* chosen, int k, int incremental,
                        int32_t* out_cols, double* out_vals, double* out_candidate);
int gjo_move_scramble(const double* cand, int n_vars, const int32_t* group_ids, int group_len,
                      int start, int count, const int32_t* perm, int incremental,
                      int32_t* out_cols, double* out_vals, double* out_candidate);
int gjo_move_insertion(const double* cand, int n_vars, const int32_t* group_ids, int group_len,
                       int get_out, int put_in, int incremental,
                       int32_t* out_cols, double* out_vals, double* out_candidate);
int gjo_move_inverse(const double* cand, int n_vars, const int32_t* group_ids, int group_len,
                     int a, int b, int incremental,
                     int32_t* out_cols, double* out_vals, double* out_candidate);
/* VariablesManager::fix_deltas / fix_variables (variables_manager.rs:187-220) */
void gjo_fix_deltas(const gjo_problem* p, const int32_t* cols, double* vals, int k);
void gjo_fix_variables(const gjo_problem* p, double* candidate, const int32_t* cols, int k);

/* ---- selection rules -------------------------------------------------------- */
/* TabuSearchBase::build_updated_population_incremental (tabu_search_base.rs:157-188):
   returns index of first minimum; *accept = best <= current.                   */
int64_t gjo_ts_select(const double* scores, int64_t S, int levels,
                      const double* current, int* accept);
/* LateAcceptanceBase (late_acceptance_base.rs:188-241): late = deque of scores,
   front at late[0].  Returns 1 on accept and updates the deque in place.       */
int gjo_la_accept(const double* cand, const double* current, double* late,
                  int* late_len, int late_size, int levels);
/* SimulatedAnnealingBase::build_updated_population_incremental
   (metaheuristic_bases/simulated_annealing_base.rs:198-233) with an explicit uniform `u`:
   temperature update per level (cooling: T *= rate, floored at 1e-7 once below 1e-6; none:
   T = inverted_accomplish_rate), accept iff u < prod_l e^(-(cand_l - cur_l) / T_l).
   temperature: [levels] in/out.  Returns 1 on accept; *proba_out = the product.            */
int gjo_sa_accept(const double* cand, const double* current, int levels, double* temperature,
                  int has_cooling_rate, double cooling_rate, double inverted_accomplish_rate,
                  double u, double* proba_out);
/* GeneticAlgorithmBase::build_updated_population (genetic_algorithm_base.rs:198-213)
   with explicit p-worst ids: winner[i] = cand[i] <= pop[worst_id[i]] ? cand : native;
   out_src[i] = i (candidate) or -(worst_id+1) (native).                         */
void gjo_ga_replace(const double* cand_scores, const double* pop_scores,
                    const int64_t* worst_ids, int64_t pop, int levels, int64_t* out_src);

/* ---- CPU baseline driver (one island per thread, solver.rs:94) -------------- */
/* Runs `n_steps` TabuSearch-style steps per thread: each step generates n_moves
   swap/2-opt moves from the thread's base, scores them pseudo-incrementally
   (clone + apply + distinct + fold, exactly the ISC arithmetic), selects per
   tabu_search_base.rs:157-188.  Returns candidates scored; *seconds = wall.    */
int64_t gjo_bench_ts(const gjo_problem* p, const double* base, int n_moves, int n_steps,
                     int n_threads, uint64_t seed, const double* move_probas,
                     const int64_t* precision, double* seconds, double* best_out);
/* Scores S plain candidates (PSC arithmetic) split over n_threads; wall time.  */
int64_t gjo_bench_plain(const gjo_problem* p, const double* samples, int64_t S,
                        int n_threads, int repeats, double* seconds, double* out);

/* GeneticAlgorithm decisions with explicit draws (genetic_algorithm_base.rs:83-134) */
int64_t gjo_ga_select(double p_best_proba, int64_t id_draw, int64_t pop, int worst, int64_t* last_top_out);
void gjo_ga_cross(const double* c1, const double* c2, int n, double weight, const uint8_t* discrete,
                  double* out1, double* out2);

/* CPU baselines for BASELINE configs 1 / 4 (LateAcceptance) and 3 (GeneticAlgorithm): one agent per
   host thread like the reference (solver.rs:94); returns candidates scored. */
int64_t gjo_bench_la(const gjo_problem* p, const double* base, const int64_t* group_offsets,
                     const int32_t* group_ids, int n_groups, int late_size, int n_steps, int n_threads,
                     uint64_t seed, const double* move_probas, const int64_t* precision,
                     double* seconds, double* best_out);
int64_t gjo_bench_ga(const gjo_problem* p, const int64_t* group_offsets, const int32_t* group_ids,
                     int n_groups, int pop, double crossover_probability, double p_best_rate,
                     int n_generations, int n_threads, uint64_t seed, const double* move_probas,
                     const int64_t* precision, double* seconds, double* best_out);

#ifdef __cplusplus
}
#endif
#endif
