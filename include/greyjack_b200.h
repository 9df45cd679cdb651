/*
 * greyjack_b200.h -- C ABI of the B200-native candidate-scoring engine for GreyJack.
 *
 * This is the drop-in boundary for ONE hot path of CameleoGrey/greyjack-solver-rust:
 * scoring populations and move neighbourhoods against the constraint model, plus the
 * move generation / selection / migration that sit directly either side of it.  The
 * reference has no FFI (it is all Rust trait objects); each entry point below names
 * the reference interface it replaces (paths relative to the reference root).  A Rust
 * maintainer binds these with an `extern "C"` block -- see INTEGRATION.md.
 *
 * Conventions: every call returns gj_status (0 = ok); on failure gj_last_error()
 * (thread-local) describes it -- nothing panics or throws across the ABI (the
 * reference itself panics: cotwin.rs:55, agent_base.rs:142).  All buffers are owned by
 * the caller.  A handle is single-threaded like the reference's `&mut self` scorer
 * (oop_score_requester.rs:336,443); distinct handles are independent and each is bound
 * to one CUDA device.  There is NO CPU fallback: without a CUDA device every compute
 * entry point fails with GJ_ERR_CUDA.
 *
 * Variable order: the reference's enumeration (oop_score_requester.rs:93-123): entity
 * by entity, field by field.  N-Queens: row_id per queen.  TSP: location_vec_id per
 * stop.  VRP: [vehicle_id_0, customer_id_0, vehicle_id_1, customer_id_1, ...].
 * Candidates travel as f64 even for integer variables, exactly like the reference's
 * Vec<f64>; decoding (frozen -> initial, clamp by total_cmp, rint ties-to-ceil) follows
 * variables/gj_integer.rs:66-83 and utils/math_utils.rs:6-8 on the device.
 */
#ifndef GREYJACK_B200_H
#define GREYJACK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define GJ_API
#else
#define GJ_API __attribute__((visibility("default")))
#endif

typedef int32_t gj_status;
enum {
    GJ_OK = 0,
    GJ_ERR_INVALID = 1,   /* bad argument / malformed description           */
    GJ_ERR_CUDA = 2,      /* CUDA runtime failure or no device              */
    GJ_ERR_UNSUPPORTED = 3,
    GJ_ERR_OOM = 4
};

/* Constraint models = the reference's example score calculators. */
enum {
    GJ_NQUEENS = 0,      /* examples/nqueens/src/score/{plain,incremental}_score_calculator.rs   */
    GJ_TSP = 1,          /* examples/tsp/src/score/{plain,incremental}_score_calculator.rs       */
    GJ_VRP = 2,          /* examples/vrp/src/score/* (CVRP, and VRPTW when time_windowed)        */
    GJ_VRP_SERVICE = 3   /* examples/vrp_service/src/score/* (lateness rule of :121-122)         */
};

/* Move ids, in the order of Mover::do_move's thresholds (mover.rs:105-121). */
enum {
    GJ_MOVE_CHANGE = 0, GJ_MOVE_SWAP = 1, GJ_MOVE_SWAP_EDGES = 2,
    GJ_MOVE_SCRAMBLE = 3, GJ_MOVE_INSERTION = 4, GJ_MOVE_INVERSE = 5
};

/* How TabuSearch / LateAcceptance islands score base + move.
   FULL : every candidate is materialised and fully re-scored -- the reference's own
          pseudo-incremental ISC semantics, bit-exact incl. the float level
          when exact sums are on.
   DELTA: the island keeps per-solution state on the device and a candidate is scored from
          the constraint terms its move touches.  N-Queens / TSP: value counts + exact base
          score; integer levels bit-exact with FULL, the float level agrees to 1e-12 relative
          before ScoreTrait::round (one 10^-precision quantum after).  VRP models: the routes
          a move touches are re-walked in the reference's order -- every level bit-exact with
          FULL (agents that score >= 8 neighbours per step; LateAcceptance / SimulatedAnnealing
          chains when move_probas has no insertion / inverse).  Moves the delta evaluator
          does not cover (listed in DESIGN.md) are re-scored by the FULL kernel inside the
          same step.                                                                        */
enum { GJ_SCORING_FULL = 0, GJ_SCORING_DELTA = 1,
       GJ_SCORING_DELTA_UNFUSED = 2,  /* DELTA as separate kernels (generate+score | select |
                                         refresh); what DELTA falls back to when an island does
                                         not fit in shared memory.  Exposed for tests / ablation. */
       GJ_SCORING_DELTA_F64 = 3       /* DELTA without the fixed-point TabuSearch step DELTA picks for
                                         TSP on a milli-unit matrix (neighbours ordered by integer
                                         tour-length changes): the same fused kernel in f64.  Exposed
                                         for tests / ablation.                                      */ };

/* Agents = AgentBuildersVariants (agents/agent_builders_variants.rs:9-18). */
enum { GJ_AGENT_TABU_SEARCH = 0, GJ_AGENT_LATE_ACCEPTANCE = 1, GJ_AGENT_GENETIC_ALGORITHM = 2,
       GJ_AGENT_SIMULATED_ANNEALING = 3 };

typedef struct gj_problem gj_problem;   /* Cotwin + OOPScoreRequester + VariablesManager */
typedef struct gj_islands gj_islands;   /* a group of Agents resident on one GPU          */

/*
 * Problem description == what CotwinBuilderTrait::build_cotwin hands to the score
 * requester (cotwin/cotwin_builder_trait.rs:7-11): planning variables (GJInteger
 * bounds / frozen / initial / semantic groups, variables/gj_integer.rs:21-64), problem
 * facts and the scorer's utility objects.
 */
typedef struct gj_problem_desc {
    int32_t kind;                  /* GJ_NQUEENS .. GJ_VRP_SERVICE                       */
    int32_t n_vars;
    const double*  lower_bounds;   /* [n_vars]                                           */
    const double*  upper_bounds;   /* [n_vars]                                           */
    const uint8_t* frozen;         /* [n_vars] or NULL                                   */
    const double*  initial;        /* [n_vars] or NULL; NaN = None (sampled uniformly)   */

    /* semantic groups (variables_manager.rs:76-106), CSR over variable ids; frozen
       variables are dropped from the groups by the library like the reference does. */
    int32_t n_groups;
    const int64_t* group_offsets;  /* [n_groups + 1]                                     */
    const int32_t* group_var_ids;

    const int64_t* column_id;      /* N-Queens: [n_vars] or NULL (= i)                   */

    /* utility objects (examples' UtilityObjectVariants) */
    int32_t n_locations;
    const double* distance_matrix; /* row-major [n_locations]^2, or NULL if coords given */
    const double* coords;          /* [n_locations][2] (lat, lon) or NULL: the matrix is
                                      then built on the device with the examples' formula
                                      round(sqrt(dlat^2+dlon^2), 3) applied twice
                                      (tsp/src/domain/location.rs:38-50,
                                       tsp/src/persistence/domain_builder.rs:42-46)      */
    int32_t n_vehicles;
    const int64_t*  vehicle_depot;     /* [n_vehicles] Vehicle.depot_vec_id              */
    const uint64_t* vehicle_capacity;
    const uint64_t* work_day_start;
    const uint64_t* work_day_end;
    const uint64_t* demand;            /* [n_locations] Customer.demand                  */
    const uint64_t* tw_start;
    const uint64_t* tw_end;
    const uint64_t* service_time;
    int32_t time_windowed;             /* VehicleRoutingPlan.time_windowed               */

    /* set_constraint_weights (plain_score_calculator.rs:40-42); PSC constraint order:
       nqueens [all_different]; tsp [no_duplicating_stops, minimize_distance];
       vrp [no_duplicating_stops, capacity, minimize_distance, late_arrival_penalty].
       The ISC has a single all_in_one constraint and uses weights[0].                   */
    double weights[4];

    /* Solver::solve arg `score_precision: Option<Vec<u64>>` (solver.rs:29); -1 = None.  */
    int64_t score_precision[3];
} gj_problem_desc;

GJ_API const char* gj_last_error(void);
GJ_API int32_t     gj_abi_version(void);
GJ_API int32_t     gj_device_count(void);
/* sizeof(gj_problem_desc) / sizeof(gj_agent_params) as compiled: lets a foreign binding
   (Rust #[repr(C)], ctypes) verify its struct layout at start-up.                       */
GJ_API size_t      gj_sizeof_problem_desc(void);
GJ_API size_t      gj_sizeof_agent_params(void);
/* Number of CUDA kernels this library has launched in this process (all handles). */
GJ_API int64_t     gj_launch_count(void);

/* OOPScoreRequester::new(cotwin) (oop_score_requester.rs:47-83): uploads everything. */
GJ_API gj_status gj_problem_create(const gj_problem_desc* desc, int32_t device, gj_problem** out);
GJ_API void      gj_problem_destroy(gj_problem* p);
GJ_API int32_t   gj_problem_levels(const gj_problem* p);   /* ScoreTrait::precision_len  */
GJ_API int32_t   gj_problem_n_vars(const gj_problem* p);
/* {Plain,Incremental}ScoreCalculator::set_constraint_weights */
GJ_API gj_status gj_problem_set_constraint_weights(gj_problem* p, const double* weights, int32_t n);
/* Float-level summation order of the TSP distance fold.  on != 0 (default): the reference's
   own sequential order, bit-exact with tsp/src/score/incremental_score_calculator.rs:76-80;
   0: per-lane partial sums + tree reduction (about 2x faster on the full-evaluation
   kernels; 1e-12 relative before rounding, one 1e-3 truncation quantum after).  The VRP
   scorers always walk each route in the reference's order.                              */
GJ_API gj_status gj_problem_set_exact_sums(gj_problem* p, int32_t on);
/* Copies the device distance matrix back (checks / warm starts).  out: [L*L]. */
GJ_API gj_status gj_problem_get_distance_matrix(gj_problem* p, double* out);

/* Pinned host memory for candidate / delta / score buffers (optional; any host
   pointer works, pinned ones make the copies asynchronous DMA).                        */
GJ_API gj_status gj_host_alloc(size_t bytes, void** out);
GJ_API void      gj_host_free(void* ptr);

/*
 * OOPScoreRequester::request_score_plain(&Vec<Vec<f64>>) -> Vec<Score>
 * (oop_score_requester.rs:336-355 -> PlainScoreCalculator::get_score,
 * plain_score_calculator.rs:60-94).  samples: host [S][n_vars] row-major;
 * scores: host [S][levels], index = sample_id.  PSC semantics.
 */
GJ_API gj_status gj_score_plain(gj_problem* p, const double* samples, int64_t S, double* scores);

/*
 * OOPScoreRequester::request_score_incremental(&Vec<f64>, &Vec<Vec<(usize,f64)>>)
 * (oop_score_requester.rs:443-463 -> IncrementalScoreCalculator::get_score,
 * incremental_score_calculator.rs:60-99).  The delta lists arrive as CSR:
 * sample j owns (var_ids[k], values[k]) for k in [offsets[j], offsets[j+1]).
 * ISC (pseudo-incremental) semantics: every candidate = base with its deltas applied
 * var-wise in emission order, then fully re-scored.
 */
GJ_API gj_status gj_score_incremental(gj_problem* p, const double* base,
                                      const uint64_t* offsets, const uint64_t* var_ids,
                                      const double* values, int64_t S, double* scores);

/* request_score_incremental with a packed wire format: u32 column ids and i32 values that are
   ALREADY inverse-transformed (what VariablesManager::inverse_transform_deltas yields for
   GJInteger variables, variables_manager.rs:165-185) -- 8 bytes per delta instead of 16.  The
   binding's flatten loop narrows the pairs while it builds the CSR arrays; the call is bound by
   the bytes that cross PCIe.  Integer variables only (all the reference's examples).            */
GJ_API gj_status gj_score_incremental_packed(gj_problem* p, const double* base,
                                             const uint64_t* offsets, const uint32_t* var_ids,
                                             const int32_t* values, int64_t S, double* scores);

/* Same two calls with every buffer already resident in device memory (HBM) and an
   explicit cudaStream_t (NULL = default stream); asynchronous.                          */
GJ_API gj_status gj_score_plain_device(gj_problem* p, const double* d_samples, int64_t S,
                                       double* d_scores, void* stream);
GJ_API gj_status gj_score_plain_i32_device(gj_problem* p, const int32_t* d_samples,
                                           int64_t row_stride, int64_t S, double* d_scores,
                                           void* stream);
GJ_API gj_status gj_score_incremental_device(gj_problem* p, const double* d_base,
                                             const uint64_t* d_offsets, const uint64_t* d_var_ids,
                                             const double* d_values, int64_t S, double* d_scores,
                                             void* stream);

/*
 * Constraint programs: the constraint REGISTRY of the reference's score calculators
 * (PlainScoreCalculator::{new, add_constraint, remove_constraint, set_constraint_weights, get_score},
 * score_calculators/plain_score_calculator.rs:20-94) for constraints that are written as data instead
 * of Polars closures.  A constraint = name + score level + up to four terms; a term = one relational
 * primitive over a planning column (variables value_offset + row * value_stride):
 *   GJ_OP_DISTINCT_DEFICIT   rows - n_unique(key_value_coef * value + key_index_coef * index(row))
 *                            (index(row) = column_id of the variable when the problem has one, else row)
 *   GJ_OP_GATHER_FOLD        fold of distance_matrix[prev][cur] along the column in row order, from and
 *                            back to the depot (location 0; the segment's depot when segmented)
 *   GJ_OP_SEGMENT_OVER_CAP   per segment: sum of demand[value] over the segment - capacity, if positive
 *   GJ_OP_MAXPLUS_LATENESS   per segment: arrival = max(arrival, tw_start) + service along the route;
 *                            lateness against tw_end by `variant` (0 file ISC, 1 service ISC, 2 PSC rule)
 * Segmented terms (seg_stride != 0) need the VRP column layout of the problem: segment (vehicle) column
 * seg_offset 0 / seg_stride 2, value (customer) column value_offset 1 / value_stride 2.
 * get_score = request_score_plain through the program: for every sample and level
 * sum over constraints of weight * (sum over terms of scale * term), constraints in insertion order.
 */
enum { GJ_OP_DISTINCT_DEFICIT = 0, GJ_OP_GATHER_FOLD = 1, GJ_OP_SEGMENT_OVER_CAP = 2, GJ_OP_MAXPLUS_LATENESS = 3 };
#define GJ_PROGRAM_MAX_TERMS 4
typedef struct gj_term {
    int32_t op;
    int32_t value_offset, value_stride;
    int32_t seg_offset, seg_stride;            /* seg_stride == 0: one segment */
    int32_t key_value_coef, key_index_coef;    /* GJ_OP_DISTINCT_DEFICIT */
    int32_t variant;                           /* GJ_OP_MAXPLUS_LATENESS */
    double  scale;                             /* constant factor of the term (1000 for the VRP duplicate penalty) */
} gj_term;
typedef struct gj_constraint {
    char    name[48];
    int32_t level;
    int32_t n_terms;
    gj_term terms[GJ_PROGRAM_MAX_TERMS];
} gj_constraint;
typedef struct gj_program gj_program;
GJ_API gj_status gj_program_create(gj_problem* p, int32_t levels, gj_program** out);          /* ::new            */
GJ_API void      gj_program_destroy(gj_program* g);
GJ_API gj_status gj_program_add_constraint(gj_program* g, const gj_constraint* c);            /* add_constraint: a new name
                                                                                                 gets weight 1.0   */
GJ_API gj_status gj_program_remove_constraint(gj_program* g, const char* name);               /* remove_constraint */
GJ_API gj_status gj_program_set_constraint_weights(gj_program* g, const char* const* names,
                                                   const double* weights, int32_t n);          /* replaces the map  */
GJ_API int32_t   gj_program_n_constraints(const gj_program* g);
GJ_API gj_status gj_program_get_score(gj_program* g, const double* samples, int64_t S, double* scores);

/*
 * Agent builders.  Field names and meaning mirror the Rust `::new` argument lists:
 *   TabuSearch::new(neighbours_count, tabu_entity_rate, compare_to_global,
 *                   mutation_rate_multiplier, move_probas, migration_frequency, termination)
 *                                                            (agents/tabu_search.rs:32-40)
 *   LateAcceptance::new(late_acceptance_size, tabu_entity_rate, mutation_rate_multiplier,
 *                   move_probas, migration_frequency, termination) (late_acceptance.rs:31-38)
 *   GeneticAlgorithm::new(population_size, crossover_probability, p_best_rate,
 *                   tabu_entity_rate, mutation_rate_multiplier, move_probas, migration_rate,
 *                   migration_frequency, termination)          (genetic_algorithm.rs:34-44)
 *   SimulatedAnnealing::new(initial_temperature, cooling_rate, tabu_entity_rate,
 *                   mutation_rate_multiplier, move_probas, migration_frequency, termination)
 *                                                            (agents/simulated_annealing.rs:31-39)
 * Termination strategies stay on the host (they are trivial control logic).
 */
typedef struct gj_agent_params {
    int32_t agent;                   /* GJ_AGENT_*                                       */
    int32_t n_islands;               /* agents of this kind resident on this GPU
                                        (Solver::solve n_jobs share, solver.rs:58-64)    */
    uint64_t seed;                   /* Philox key; the reference is entropy-seeded      */
    int64_t neighbours_count;        /* TS                                               */
    int64_t late_acceptance_size;    /* LA                                               */
    int64_t population_size;         /* GA                                               */
    double  crossover_probability;   /* GA                                               */
    double  p_best_rate;             /* GA                                               */
    double  migration_rate;          /* GA                                               */
    double  tabu_entity_rate;
    int32_t compare_to_global;       /* TS                                               */
    int32_t has_mutation_rate_multiplier;  /* Option<f64>: 0 = None (-> 0.0)             */
    double  mutation_rate_multiplier;
    int32_t has_move_probas;         /* Option<Vec<f64>>: 0 = None (uniform, mover.rs:38-48) */
    double  move_probas[6];
    int64_t migration_frequency;
    int32_t reference_noop_moves;    /* 1 = reproduce the reference's incremental-form
                                        no-op scramble / swap_edges(k=2);
                                        0 = apply the plain-form permutation             */
    int32_t scoring_mode;            /* GJ_SCORING_*: how TS / LA islands score a neighbour     */
    int32_t chain_steps_per_launch;  /* LateAcceptance + GJ_SCORING_DELTA: steps a chain runs between
                                        two refreshes of the global top (one kernel launch); 0 = 8. 
                                        The reference refreshes after every step of every agent
                                        thread (agent_base.rs:185); 1 reproduces that cadence.    */
    int32_t has_cooling_rate;        /* SA: Option<f64>: 0 = None (temperature = 1 - accomplish rate,
                                        agent_base.rs:537-552, see gj_islands_set_accomplish_rate)  */
    double  cooling_rate;            /* SA                                                       */
    double  initial_temperature[3];  /* SA: one per score level                                  */
} gj_agent_params;

/* <Agent>::build_agent + Agent::init_population (agent_base.rs:190-218).  `initial`:
   host [n_islands][n_vars] start vectors (InitialSolutionVariants) or NULL to use the
   problem's initial values (None entries sampled uniformly in [lb, ub]).               */
GJ_API gj_status gj_islands_create(gj_problem* p, const gj_agent_params* params,
                                   const double* initial, gj_islands** out);
GJ_API void      gj_islands_destroy(gj_islands* g);

/* Agent::solve loop body x n_steps for every island, entirely on the device
   (agent_base.rs:135-186: step_plain | step_incremental, sort, update_top_individual;
   migration every migration_frequency steps inside the group, ring i -> i+1,
   solver.rs:85-92; global best shared as in update_global_top :446-490).              */
GJ_API gj_status gj_islands_step(gj_islands* g, int64_t n_steps, void* stream);

/* Agent::set_agent_step_dependent_params (agent_base.rs:537-552): SimulatedAnnealing without a
   cooling rate uses temperature = 1 - termination_strategy.get_accomplish_rate(); the host owns
   the termination strategy and passes its accomplish rate before stepping.                     */
GJ_API gj_status gj_islands_set_accomplish_rate(gj_islands* g, double accomplish_rate);

/* Counters: candidates scored so far, steps done, accepted moves. */
GJ_API gj_status gj_islands_stats(gj_islands* g, int64_t* candidates, int64_t* steps,
                                  int64_t* accepted);

/* Optional CUDA-event timing of the dominant (scoring) kernel of every step, recorded on the
   launching stream; profile_read synchronises, returns the summed duration and resets.     */
GJ_API gj_status gj_islands_set_profiling(gj_islands* g, int32_t on);
GJ_API gj_status gj_islands_profile_read(gj_islands* g, double* total_ms, int64_t* launches);

/* agent_top_individual of one island, or (island < 0) the group's global_top_individual.
   vars: host [n_vars] f64; score: host [levels] (rounded by score_precision).          */
GJ_API gj_status gj_islands_best(gj_islands* g, int32_t island, double* vars, double* score);
/* population[0] of one island (current, not best). */
GJ_API gj_status gj_islands_current(gj_islands* g, int32_t island, double* vars, double* score);

/* Cross-GPU migration plumbing (AgentToAgentUpdate, agent_to_agent_update.rs): packs
   the migrants of the LAST island of this group into a device buffer / accepts migrants
   into the FIRST island with the reference's acceptance rule (agent_base.rs:414-440).
   The transport between ranks (NCCL send/recv over NVLink) is the caller's.
   Layout: n_migrants x (n_vars int32 + levels f64), see gj_islands_migrant_bytes.
   The two calls are the two halves of ONE migration and keep the reference's order inside the
   group (agent_base.rs:161-183, even agents send then receive, odd agents receive then send):
   export = even islands send, odd islands receive and then send; import = even islands
   receive (island 0 from d_buffer).  An import must follow the export of the same exchange.  */
GJ_API int64_t   gj_islands_migrant_bytes(const gj_islands* g);
/* on != 0: gj_islands_step stops doing the wrap-around (last island -> first island) and the
   caller moves the migrants between GPUs with export/import every migration_frequency steps;
   island_base = global index of this group's island 0 (RNG key / ring position).          */
GJ_API gj_status gj_islands_set_external_ring(gj_islands* g, int32_t on, int32_t island_base);
GJ_API gj_status gj_islands_export_migrants(gj_islands* g, void* d_buffer, void* stream);
GJ_API gj_status gj_islands_import_migrants(gj_islands* g, const void* d_buffer, void* stream);

/* The group's global_top_individual as a device record [n_vars int32 (row stride) | levels f64], for a
   caller-owned transport between ranks (NCCL all-gather, gloo in tests): export packs it,
   import takes `count` records (gj_islands_global_top_bytes apart) and adopts the best one when it is
   STRICTLY better than the group's own (agent_base.rs:451); islands then adopt it by their usual rule
   (agent_base.rs:465-489).  Replaces the Arc<Mutex<global_top_individual>> across processes.        */
GJ_API int64_t   gj_islands_global_top_bytes(const gj_islands* g);
GJ_API gj_status gj_islands_export_global_top(gj_islands* g, void* d_buffer, void* stream);
GJ_API gj_status gj_islands_import_global_top(gj_islands* g, const void* d_records, int32_t count, void* stream);

/*
 * The ring and the shared global top across GPUs over CUDA peer memory (NVLink), one process per GPU,
 * no collective library on the data path (csrc/gj_ring.cu).  Set-up: every rank creates its ring,
 * publishes gj_ring_handle (an IPC handle of its inbox) to the others by any host channel, and
 * connects with the table of all handles.  gj_ring_exchange, called by every rank every
 * migration_frequency steps on its stream: migrants of the rank's last island -> first island of
 * rank + 1 (solver.rs:85-92, acceptance rule agent_base.rs:414-440), every rank's global top -> every
 * rank (agent_base.rs:446-490).  Asynchronous: nothing waits on the host.  Requires
 * gj_islands_set_external_ring(g, 1, rank * islands_per_rank).
 */
typedef struct gj_ring gj_ring;
typedef struct gj_peer_handle { unsigned char bytes[64]; } gj_peer_handle;
GJ_API gj_status gj_ring_create(gj_islands* g, int32_t rank, int32_t world, gj_ring** out);
GJ_API void      gj_ring_destroy(gj_ring* r);
GJ_API gj_status gj_ring_handle(gj_ring* r, gj_peer_handle* out);
GJ_API gj_status gj_ring_connect(gj_ring* r, const gj_peer_handle* handles /*[world]*/);
GJ_API gj_status gj_ring_exchange(gj_ring* r, void* stream);
/* exchanges done; flags that did not arrive within the time-out (a missed exchange, never a hang) */
GJ_API gj_status gj_ring_stats(gj_ring* r, int64_t* exchanges, int64_t* missed);

/*
 * Test / inspection hook: runs ONE TabuSearch/LateAcceptance step of one island and
 * returns what happened -- the generated moves as delta lists (CSR, capacities given
 * by the caller), each candidate's rounded score, the selected index and whether it was
 * accepted -- so the path can be checked move by move against the reference mover and
 * scorer (mover.rs:145-421, tabu_search_base.rs:157-188).
 */
GJ_API gj_status gj_islands_trace_step(gj_islands* g, int32_t island,
                                       uint64_t* offsets /*[K+1]*/, uint64_t* var_ids,
                                       double* values, int64_t delta_capacity,
                                       int32_t* move_kinds /*[K]*/,
                                       int32_t* move_desc /*[K][20]: kind, group, k, 0,
                                                            chosen positions[8], values[8]*/,
                                       double* scores /*[K][levels]*/,
                                       int64_t* selected, int32_t* accepted);

/* Test hook next to gj_islands_trace_step: what the acceptance rule of the LAST traced step saw.
   out[0] = the uniform random value, out[1] = accept probability (SimulatedAnnealing),
   out[2..4] = temperatures per level after the step's update.                              */
GJ_API gj_status gj_islands_trace_aux(gj_islands* g, int32_t island, double* out /*[5]*/);

/* Test / inspection hooks for GeneticAlgorithm islands: one generation of every island with the
   decisions of ONE island exposed, so that select_p_best / cross / Mover::do_move(plain) /
   build_updated_population (genetic_algorithm_base.rs:83-134, 141-213) can be replayed decision by
   decision through the reference's rules.  Every buffer is the caller's; any may be NULL.          */
typedef struct gj_ga_trace {
    int32_t* order_before;  /* [pop] rank -> individual index: population.sort() before sampling      */
    double*  pairs;         /* [ceil(pop/2)][8]: select_p_best draws of the pair {p1, last_top1, id1,
                               p2, last_top2, id2}, the crossover coin u, the crossover weight w
                               before rint (-1: no crossover)                                        */
    int32_t* move_desc;     /* [n_cand][20] as gj_islands_trace_step                                  */
    double*  cand_rows;     /* [n_cand][n_vars] offspring after cross + move + fix_variables          */
    double*  cand_scores;   /* [n_cand][levels] their PSC scores, rounded (agent_base.rs:284-287)     */
    double*  replace;       /* [pop][3] select_p_worst draws of slot i: {p, last_top, id}             */
    int32_t* src;           /* [pop] slot i of the new population: i = candidate i, -(rank+1) = the
                               p-worst native of that rank                                          */
} gj_ga_trace;
GJ_API gj_status gj_islands_ga_trace_generation(gj_islands* g, int32_t island, gj_ga_trace* out);
/* population of one GA island as stored (unsorted) + its rank table; rows [pop][n_vars],
   scores [pop][levels], order [pop].                                                              */
GJ_API gj_status gj_islands_ga_population(gj_islands* g, int32_t island, double* rows, double* scores,
                                          int32_t* order);

/* Test / inspection hook: the tabu deque of one island and semantic group
   (Mover::tabu_ids_vecdeque_map, mover.rs:75-96), newest id first.  ids: caller buffer of `capacity`
   ints (group positions); *fill = ids currently held, *size = the deque's capacity
   max(ceil(tabu_entity_rate * group_len), 1) (0 when tabu_entity_rate == 0).                      */
GJ_API gj_status gj_islands_trace_tabu(gj_islands* g, int32_t island, int32_t group, int32_t* ids,
                                       int32_t capacity, int32_t* fill, int32_t* size);

/* Name of the kernel path a step of this group takes, chosen at creation: "fused", "fused_fixed", "fused_lean",
   "chain", "vrp_chain", "delta", "vrp_delta", "full" or "ga" (static string).                     */
GJ_API const char* gj_islands_step_path(const gj_islands* g);

#ifdef __cplusplus
}
#endif
#endif /* GREYJACK_B200_H */
