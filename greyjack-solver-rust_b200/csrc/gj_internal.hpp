// gj_internal.hpp -- host-side state behind the opaque C-ABI handles.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/greyjack_b200.h"
#include "gj_device.cuh"

void gj_set_error(const std::string& msg);
gj_status gj_fail(gj_status code, const std::string& msg);

#define GJ_CUDA_TRY(expr)                                                            \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess)                                                       \
            return gj_fail(GJ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// Placed after every kernel launch: counts it (gj_launch_count, bench.py's gpu_launches) and
// surfaces launch-configuration errors.
extern std::atomic<long long> gj_launch_counter;
#define GJ_LAUNCH_CHECK()                                        \
    do {                                                         \
        gj_launch_counter.fetch_add(1, std::memory_order_relaxed); \
        GJ_CUDA_TRY(cudaGetLastError());                         \
    } while (0)

// Grow-only device / pinned-host scratch buffer.
struct GjBuffer {
    void* ptr = nullptr;
    size_t bytes = 0;
    gj_status reserve(size_t need);
    void release();
};

struct gj_problem {
    int device = 0;
    GjProblemDev dev{};
    std::vector<void*> allocs;             // device allocations owned by the handle
    cudaStream_t stream = nullptr;         // stream of the host-buffer entry points

    // host mirrors (VariablesManager state)
    std::vector<double> lb, ub, initial;
    std::vector<uint8_t> frozen;
    std::vector<std::vector<int32_t>> groups;   // semantic groups, frozen ids dropped
    int64_t precision[3] = {-1, -1, -1};
    bool symmetric_D = false;
    int d32_state = 0;                     // milli-unit matrix (dev.D32): 0 not built, 1 valid, -1 unavailable
    int n_warps_vrp = 4;

    // scratch for host-buffer calls
    GjBuffer d_samples, d_scores, d_base, d_base_i32, d_offsets, d_ids, d_vals;

    ~gj_problem();
};

gj_status gj_problem_ensure_d32(gj_problem* p);

// launchers shared by the ABI entry points and the island code (gj_score.cu)
gj_status gj_launch_score_plain_f64(gj_problem* p, const double* d_samples, int64_t S,
                                    double* d_scores, bool isc, cudaStream_t st);
gj_status gj_launch_score_plain_i32(gj_problem* p, const int32_t* d_samples, int64_t stride,
                                    int64_t S, double* d_scores, bool isc, cudaStream_t st);
gj_status gj_launch_score_incremental(gj_problem* p, const double* d_base, int32_t* d_base_i32,
                                      const uint64_t* d_offsets, const uint64_t* d_ids,
                                      const double* d_vals, int64_t S, double* d_scores,
                                      cudaStream_t st);
