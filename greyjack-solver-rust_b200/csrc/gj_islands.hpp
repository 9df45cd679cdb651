// gj_islands.hpp -- state of a group of device-resident agents (opaque gj_islands handle).
#pragma once

#include <vector>

#include "gj_internal.hpp"
#include "gj_moves.cuh"
#include "gj_delta.cuh"
#include "gj_vrp_delta.cuh"

struct gj_islands {
    gj_problem* p = nullptr;
    int device = 0;              // p->device (kept here: the handle may outlive a destroyed problem in a failing caller)
    gj_agent_params prm{};
    int I = 0;                   // islands in the group
    int K = 1;                   // candidates per island per step (TS: neighbours; LA: 1)
    int levels = 1, n_vars = 0, stride = 0;
    int noop = 0;
    int late_size = 0;
    bool external_ring = false;  // migration transport handled by the caller (multi-GPU)
    int island_base = 0;         // global id of island 0 (RNG key, ring position)
    uint64_t step = 0;
    int64_t steps_to_send = 1;
    int64_t migrants = 1;        // individuals per exchange (GA: ceil(rate * pop))

    GjGroups groups{};
    GjMoverParams mover{};
    std::vector<void*> allocs;

    // local-search agents (TS / LA): population[0] per island
    int32_t* cur = nullptr;  double* cur_score = nullptr;
    int32_t* best = nullptr; double* best_score = nullptr;      // agent_top_individual
    int32_t* gbest = nullptr; double* gbest_score = nullptr;    // global_top_individual
    int* dirty = nullptr;
    int* gver = nullptr;         // version of the published global top
    int* gseen = nullptr;        // [I] last version every island has looked at (fused islands)
    GjMove* moves = nullptr;
    double* cand_scores = nullptr;
    unsigned char* mailbox = nullptr;
    double* late = nullptr; int* late_head = nullptr; int* late_len = nullptr;
    long long* selected = nullptr; int* accepted = nullptr;

    // tabu deques: rank-indexed (slot 0 = newest id), double-buffered by step parity
    uint32_t* tabu_bits = nullptr; int tabu_words = 0; const int32_t* tabu_word_off = nullptr;
    int32_t* tabu_ring[2] = {nullptr, nullptr}; int tabu_ring_len = 0; const int32_t* tabu_ring_off = nullptr;
    const int32_t* tabu_size = nullptr; int* tabu_fill = nullptr;

    // delta scoring (GJ_SCORING_DELTA, gj_delta.cuh)
    int scoring_mode = 0;
    bool delta_may_fallback = true;      // some generated moves may need the full evaluator
    GjDeltaState ds{};
    GjVrpState vs{};                     // VRP route-level base state (gj_vrp_delta.cuh)
    int* worklist = nullptr;             // [I*K] neighbours queued for the full evaluator
    int* work_count = nullptr;
    // fused single-kernel step (gj_islands_fused.cuh)
    bool fused = false;
    bool fused_lean = false;             // long solutions: only the solution is staged in shared memory
    // lean layout: the edge lengths / unrounded terms of the published global top, for adopters to copy
    int* top_is_cur = nullptr;           // [I] the agent's top row IS its current row (set at the end of a step)
    double* gedge = nullptr; double* graw = nullptr; int* gedge_ver = nullptr;
    int fused_threads = 0, fused_clones = 0, fused_mb = 0;
    size_t fused_smem = 0;
    long long* phase_clocks = nullptr;   // GJ_PHASE_TIMING=1 development aid
    // fixed-point TabuSearch step for TSP (gj_islands_tsfast.cuh): the fused step with integer arithmetic
    bool ts_fast = false;
    double* ts_edge = nullptr;           // [I][ts_edge_stride] persistent f64 edge lengths
    int ts_edge_stride = 0;
    unsigned long long* ts_pub = nullptr;   // {best key offered, key of the published global top, lock}
    unsigned int* done_counter = nullptr;   // islands finished in the current launch (the last one publishes the global top)
    // LateAcceptance chains: many steps per launch, one warp per island (gj_islands_chain.cuh)
    bool chain = false;
    bool vrp_chain = false;              // ... on a VRP model: route index in HBM (gj_islands_vrp_chain.cuh)
    GjVrpChainState vcs{};
    size_t chain_bytes = 0;              // shared memory per chain
    uint32_t* ctabu = nullptr; int ctabu_words = 0; const int32_t* ctabu_off = nullptr;
    GjMove* trace_moves = nullptr; double* trace_scores = nullptr; int* trace_accept = nullptr;
    // SimulatedAnnealing: temperatures [I][GJ_MAX_LEVELS] on the device, schedule, trace of the rule
    double* sa_temp = nullptr; GjSaParams sa{}; double* trace_aux = nullptr;

    unsigned long long* counters = nullptr;

    // optional CUDA-event timing of the dominant (scoring) kernel, on the launching stream
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;

    // genetic algorithm (gj_islands_ga.cu)
    int pop = 0, half = 0, n_cand = 0;
    int32_t* pop_rows = nullptr;   // [I][pop][stride]
    int32_t* pop_next = nullptr;
    double* pop_scores = nullptr;  // [I][pop][3]
    double* pop_scores_next = nullptr;
    int32_t* cand_rows = nullptr;  // [I][n_cand][stride]
    int* order = nullptr;          // [I][pop] rank -> row index after the sort
    int* ga_src = nullptr;         // [I][pop] replacement source (trace)
    int32_t* ga_pairs = nullptr;   // [I][n_cand][2 + 2 * GJ_MOVE_MAXPAIRS] planned small moves as (column, value) pairs
    int* ga_take = nullptr;        // [I][pop] source slot of every new individual, bit 31 = offspring of that slot
    cudaStream_t ga_side = nullptr; cudaEvent_t ga_ev[4] = {nullptr, nullptr, nullptr, nullptr}; int ga_side_state = 0;
    GjMove* ga_moves_next = nullptr; int64_t ga_moves_step = -1;   // moves of generation ga_moves_step, generated ahead
    int* ga_parent = nullptr;      // [I][n_cand] population slot every planned offspring descends from
    int* ga_rank = nullptr;        // [I][pop] rank scratch of the counting sort (kept zeroed)
    double* ga_trace_sel = nullptr;  // [I][half][8] trace of the parent draws (gj_islands_ga_trace_generation)
    double* ga_trace_rep = nullptr;  // [I][pop][3] trace of the p-worst draws

    ~gj_islands();
};

int gj_tabu_deque_size(double rate, int group_len);
gj_status gj_islands_common_init(gj_islands* g, gj_problem* p, const gj_agent_params* prm);
void gj_islands_start_vector(const gj_problem* p, const double* given, uint64_t& rng, std::vector<int32_t>& row);

gj_status gj_ga_create(gj_problem* p, const gj_agent_params* prm, const double* initial, gj_islands** out);
gj_status gj_ga_step(gj_islands* g, int64_t n_steps, cudaStream_t st);
gj_status gj_ga_global_top(gj_islands* g, cudaStream_t st);
gj_status gj_ga_current(gj_islands* g, int32_t island, double* vars, double* score);
gj_status gj_ga_export(gj_islands* g, void* d_buffer, cudaStream_t st);
gj_status gj_ga_pack_outgoing(gj_islands* g, cudaStream_t st, const unsigned char** d_slot);
gj_status gj_islands_pack_outgoing(gj_islands* g, cudaStream_t st, const unsigned char** d_slot);
gj_status gj_ga_import(gj_islands* g, const void* d_buffer, cudaStream_t st);
gj_status gj_ls_global_top(gj_islands* g, cudaStream_t st);
gj_status gj_prof_begin(gj_islands* g, cudaStream_t st);
gj_status gj_prof_end(gj_islands* g, cudaStream_t st);
