// gj_score.cu -- the two scoring entry points of the reference's score requester,
// rebuilt as kernels:
//   request_score_plain        oop_score_requester.rs:336-355  -> k_plain_*
//   request_score_incremental  oop_score_requester.rs:443-463  -> k_incr_*
// The Polars marshalling of the reference (build_group_data_map / build_delta_dfs,
// :261-334, :384-441) has no device counterpart: candidates stay flat arrays and the
// "DataFrame" is the kernel's shared memory.
#include <algorithm>

#include "gj_eval.cuh"
#include "gj_internal.hpp"

static constexpr int kWarpsPerCta = 4;

// ---- plain: one warp per candidate (N-Queens, TSP) -------------------------------------
template <int KIND, class RowT>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_plain_warp(GjProblemDev P, const RowT* __restrict__ samples, int64_t stride, int64_t S,
             int isc, double* __restrict__ scores) {
    extern __shared__ uint32_t smem_u32[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int words = P.bm_words + P.desc_words + P.asc_words;
    uint32_t* bm = smem_u32 + warp * words;
    const int64_t j = (int64_t)blockIdx.x * kWarpsPerCta + warp;
    if (j >= S) return;
    const RowT* row = samples + j * stride;
    double out[GJ_MAX_LEVELS];
    if constexpr (sizeof(RowT) == 8) {
        GjSrcF64 src{(const double*)row, &P};
        if constexpr (KIND == GJ_NQUEENS) {
            gj_combine_nqueens(P, gj_nqueens_eval_warp(P, src, bm, lane), out);
        } else {
            double dup, dist;
            gj_tsp_eval_warp(P, src, bm, lane, dup, dist);
            gj_combine_tsp(P, isc != 0, dup, dist, out);
        }
    } else {
        GjSrcI32 src{(const int32_t*)row};
        if constexpr (KIND == GJ_NQUEENS) {
            gj_combine_nqueens(P, gj_nqueens_eval_warp(P, src, bm, lane), out);
        } else {
            double dup, dist;
            gj_tsp_eval_warp(P, src, bm, lane, dup, dist);
            gj_combine_tsp(P, isc != 0, dup, dist, out);
        }
    }
    if (lane == 0)
        for (int l = 0; l < P.levels; ++l) scores[j * P.levels + l] = out[l];
}

// ---- plain: one CTA per candidate (VRP) -------------------------------------------------
template <class RowT>
__global__ void __launch_bounds__(kVrpWarps * 32)
k_plain_vrp(GjProblemDev P, const RowT* __restrict__ samples, int64_t stride, int64_t S,
            int isc, double* __restrict__ scores) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = P.n_entities;
    GjVrpSmem s = gj_vrp_carve(smem_raw, n, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
    const int64_t j = blockIdx.x;
    const RowT* row = samples + j * stride;
    if constexpr (sizeof(RowT) == 8) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double2 pr = *reinterpret_cast<const double2*>((const double*)row + 2 * i);
            s.veh[i] = (uint16_t)gj_decode(P, 2 * i, pr.x);
            s.cust[i] = gj_decode(P, 2 * i + 1, pr.y);
        }
    } else {
        // a candidate row is read exactly once: streaming (evict-first) loads keep the gather tables
        // -- the distance matrix above all -- resident in L2 while 131 MB of offspring go by.  Eight
        // loads per thread are in flight before the first store: one DRAM round trip per 2 048 stops
        // instead of one per loop iteration.
        constexpr int U = 8;
        for (int i0 = threadIdx.x; i0 < n; i0 += U * blockDim.x) {
            int2 pr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * blockDim.x;
                pr[u] = make_int2(0, 0);
                if (i < n) pr[u] = __ldcs(reinterpret_cast<const int2*>((const int32_t*)row + 2 * i));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i < n) { s.veh[i] = (uint16_t)pr[u].x; s.cust[i] = pr[u].y; }
            }
        }
    }
    __syncthreads();
    const int tw_mode = isc ? (P.kind == GJ_VRP_SERVICE ? GJ_TW_ISC_SERVICE : GJ_TW_ISC_FILE)
                            : GJ_TW_PSC;
    double dup1000 = 0, cap = 0, dist = 0, late = 0;
    gj_vrp_eval_cta(P, s, tw_mode, dup1000, cap, dist, late);
    if (threadIdx.x == 0) {
        double out[GJ_MAX_LEVELS];
        gj_combine_vrp(P, isc != 0, dup1000, cap, dist, late, out);
        for (int l = 0; l < 3; ++l) scores[j * 3 + l] = out[l];
    }
}

// ---- incremental ------------------------------------------------------------------------
// inverse_transform_variables(base) (variables_manager.rs:136-152), once per call.
__global__ void k_decode_base(GjProblemDev P, const double* __restrict__ base,
                              int32_t* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n_vars; i += gridDim.x * blockDim.x)
        out[i] = gj_decode(P, i, base[i]);
}

// Applies one sample's delta list to the shared-memory clone of the base, 32 deltas at a
// time, in emission order: when several deltas of one chunk hit the same variable the
// last one wins (the stored individual is updated the same way,
// tabu_search_base.rs:175-178).  Must be called by a full warp.
// A delta value on the wire: the reference's f64 (decoded like any variable) or, in the packed
// form, an int32 that is already an inverse-transformed value (clamped into the bounds again).
__device__ __forceinline__ int gj_decode_delta(const GjProblemDev& P, int id, double v) { return gj_decode(P, id, v); }
__device__ __forceinline__ int gj_decode_delta(const GjProblemDev& P, int id, int32_t v) {
    if (P.frozen[id]) return (int)P.initial[id];
    return min(max(v, P.lbi[id]), P.ubi[id]);
}

template <class IdT, class ValT, class StoreFn>
__device__ __forceinline__ void gj_apply_deltas_warp(const GjProblemDev& P, int lane,
                                                     const IdT* __restrict__ ids,
                                                     const ValT* __restrict__ vals,
                                                     uint64_t b, uint64_t e, StoreFn store) {
    for (uint64_t k0 = b; k0 < e; k0 += 32) {
        const uint64_t k = k0 + lane;
        const bool on = k < e;
        uint64_t id64 = on ? (uint64_t)ids[k] : 0;
        const bool valid = on && id64 < (uint64_t)P.n_vars;
        const int id = valid ? (int)id64 : -1 - lane;
        const unsigned grp = __match_any_sync(GJ_FULL_MASK, id);
        if (valid && lane == 31 - __clz(grp)) store(id, gj_decode_delta(P, id, vals[k]));
        __syncwarp();
    }
}

template <int KIND, class IdT, class ValT>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_incr_warp(GjProblemDev P, const int32_t* __restrict__ base, const uint64_t* __restrict__ offsets,
            const IdT* __restrict__ ids, const ValT* __restrict__ vals, int64_t S,
            double* __restrict__ scores) {
    extern __shared__ uint32_t smem_u32[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int words = P.bm_words + P.desc_words + P.asc_words;
    const int per_warp = words + P.n_vars;
    uint32_t* bm = smem_u32 + warp * per_warp;
    int32_t* cand = (int32_t*)(bm + words);
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;      // 1..kWarpsPerCta warps per CTA
    if (j >= S) return;
    // clone of the planning ids (tsp ISC :64, nqueens ISC :42)
    for (int i = lane; i < P.n_vars; i += 32) cand[i] = base[i];
    __syncwarp();
    gj_apply_deltas_warp(P, lane, ids, vals, offsets[j], offsets[j + 1],
                         [&](int id, int v) { cand[id] = v; });
    GjSrcI32 src{cand};
    double out[GJ_MAX_LEVELS];
    if constexpr (KIND == GJ_NQUEENS) {
        gj_combine_nqueens(P, gj_nqueens_eval_warp(P, src, bm, lane), out);
    } else {
        double dup, dist;
        gj_tsp_eval_warp(P, src, bm, lane, dup, dist);
        gj_combine_tsp(P, true, dup, dist, out);
    }
    if (lane == 0)
        for (int l = 0; l < P.levels; ++l) scores[j * P.levels + l] = out[l];
}

template <class IdT, class ValT>
__global__ void __launch_bounds__(kVrpWarps * 32)
k_incr_vrp(GjProblemDev P, const int32_t* __restrict__ base, const uint64_t* __restrict__ offsets,
           const IdT* __restrict__ ids, const ValT* __restrict__ vals, int64_t S,
           double* __restrict__ scores) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = P.n_entities;
    GjVrpSmem s = gj_vrp_carve(smem_raw, n, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
    const int64_t j = blockIdx.x;
    // clones of candidate_vehicle_ids / candidate_customer_ids (vrp ISC :63-64)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int2 pr = *reinterpret_cast<const int2*>(base + 2 * i);
        s.veh[i] = (uint16_t)pr.x;
        s.cust[i] = pr.y;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // var-wise application (the same-row double delta of the reference's
        // delta frame is a declared deviation; this is the PSC / stored-individual semantics)
        gj_apply_deltas_warp(P, threadIdx.x, ids, vals, offsets[j], offsets[j + 1],
                             [&](int id, int v) {
                                 if (id & 1) s.cust[id >> 1] = v; else s.veh[id >> 1] = (uint16_t)v;
                             });
    }
    __syncthreads();
    const int tw_mode = (P.kind == GJ_VRP_SERVICE) ? GJ_TW_ISC_SERVICE : GJ_TW_ISC_FILE;
    double dup1000 = 0, cap = 0, dist = 0, late = 0;
    gj_vrp_eval_cta(P, s, tw_mode, dup1000, cap, dist, late);
    if (threadIdx.x == 0) {
        double out[GJ_MAX_LEVELS];
        gj_combine_vrp(P, true, dup1000, cap, dist, late, out);
        for (int l = 0; l < 3; ++l) scores[j * 3 + l] = out[l];
    }
}

// ---- launchers ----------------------------------------------------------------------------

template <class K>
static gj_status set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024)
        GJ_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return GJ_OK;
}

template <class RowT>
static gj_status launch_plain(gj_problem* p, const RowT* d_samples, int64_t stride, int64_t S,
                              double* d_scores, bool isc, cudaStream_t st) {
    if (S <= 0) return GJ_OK;
    const GjProblemDev& P = p->dev;
    gj_status rc;
    if (P.kind >= GJ_VRP) {
        size_t smem = gj_vrp_smem_bytes(P.n_entities, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
        if ((rc = set_smem(k_plain_vrp<RowT>, smem))) return rc;
        k_plain_vrp<RowT><<<(unsigned)S, kVrpWarps * 32, smem, st>>>(P, d_samples, stride, S, isc, d_scores);
    } else {
        size_t smem = (size_t)kWarpsPerCta * (size_t)(P.bm_words + P.desc_words + P.asc_words) * 4;
        unsigned grid = (unsigned)((S + kWarpsPerCta - 1) / kWarpsPerCta);
        if (P.kind == GJ_NQUEENS) {
            if ((rc = set_smem(k_plain_warp<GJ_NQUEENS, RowT>, smem))) return rc;
            k_plain_warp<GJ_NQUEENS, RowT><<<grid, kWarpsPerCta * 32, smem, st>>>(P, d_samples, stride, S, isc, d_scores);
        } else {
            if ((rc = set_smem(k_plain_warp<GJ_TSP, RowT>, smem))) return rc;
            k_plain_warp<GJ_TSP, RowT><<<grid, kWarpsPerCta * 32, smem, st>>>(P, d_samples, stride, S, isc, d_scores);
        }
    }
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

gj_status gj_launch_score_plain_f64(gj_problem* p, const double* d_samples, int64_t S,
                                    double* d_scores, bool isc, cudaStream_t st) {
    return launch_plain<double>(p, d_samples, p->dev.n_vars, S, d_scores, isc, st);
}

gj_status gj_launch_score_plain_i32(gj_problem* p, const int32_t* d_samples, int64_t stride,
                                    int64_t S, double* d_scores, bool isc, cudaStream_t st) {
    if (p->dev.kind >= GJ_VRP && (stride % 2) != 0)
        return gj_fail(GJ_ERR_INVALID, "VRP int32 rows need an even stride");
    return launch_plain<int32_t>(p, d_samples, stride, S, d_scores, isc, st);
}

template <class IdT, class ValT>
static gj_status launch_incremental(gj_problem* p, const double* d_base, int32_t* d_base_i32,
                                    const uint64_t* d_offsets, const IdT* d_ids, const ValT* d_vals,
                                    int64_t S, double* d_scores, cudaStream_t st) {
    if (S <= 0) return GJ_OK;
    const GjProblemDev& P = p->dev;
    gj_status rc;
    if (d_base) {
        k_decode_base<<<std::min(148, (P.n_vars + 255) / 256), 256, 0, st>>>(P, d_base, d_base_i32);
        GJ_LAUNCH_CHECK();
    }
    if (P.kind >= GJ_VRP) {
        size_t smem = gj_vrp_smem_bytes(P.n_entities, P.n_vehicles, P.bm_words, kVrpWarps, !P.time_windowed);
        if ((rc = set_smem(k_incr_vrp<IdT, ValT>, smem))) return rc;
        k_incr_vrp<IdT, ValT><<<(unsigned)S, kVrpWarps * 32, smem, st>>>(P, d_base_i32, d_offsets, d_ids, d_vals, S, d_scores);
    } else {
        // one shared-memory clone per warp: fewer warps per CTA for large instances
        size_t per_warp = (size_t)(P.bm_words + P.desc_words + P.asc_words + P.n_vars) * 4;
        const int warps = (int)std::min<size_t>(kWarpsPerCta, (220 * 1024) / per_warp);
        if (warps < 1) return gj_fail(GJ_ERR_UNSUPPORTED, "instance too large for the shared-memory candidate clone");
        size_t smem = per_warp * warps;
        unsigned grid = (unsigned)((S + warps - 1) / warps);
        if (P.kind == GJ_NQUEENS) {
            if ((rc = set_smem(k_incr_warp<GJ_NQUEENS, IdT, ValT>, smem))) return rc;
            k_incr_warp<GJ_NQUEENS, IdT, ValT><<<grid, warps * 32, smem, st>>>(P, d_base_i32, d_offsets, d_ids, d_vals, S, d_scores);
        } else {
            if ((rc = set_smem(k_incr_warp<GJ_TSP, IdT, ValT>, smem))) return rc;
            k_incr_warp<GJ_TSP, IdT, ValT><<<grid, warps * 32, smem, st>>>(P, d_base_i32, d_offsets, d_ids, d_vals, S, d_scores);
        }
    }
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

gj_status gj_launch_score_incremental(gj_problem* p, const double* d_base, int32_t* d_base_i32,
                                      const uint64_t* d_offsets, const uint64_t* d_ids,
                                      const double* d_vals, int64_t S, double* d_scores,
                                      cudaStream_t st) {
    return launch_incremental<uint64_t, double>(p, d_base, d_base_i32, d_offsets, d_ids, d_vals, S, d_scores, st);
}

// ---- C ABI ------------------------------------------------------------------------------------

extern "C" gj_status gj_score_plain_device(gj_problem* p, const double* d_samples, int64_t S,
                                           double* d_scores, void* stream) {
    if (!p || !d_samples || !d_scores || S < 0) return gj_fail(GJ_ERR_INVALID, "bad argument");
    GJ_CUDA_TRY(cudaSetDevice(p->device));
    return gj_launch_score_plain_f64(p, d_samples, S, d_scores, false, (cudaStream_t)stream);
}

extern "C" gj_status gj_score_plain_i32_device(gj_problem* p, const int32_t* d_samples,
                                               int64_t row_stride, int64_t S, double* d_scores,
                                               void* stream) {
    if (!p || !d_samples || !d_scores || S < 0 || row_stride < p->dev.n_vars)
        return gj_fail(GJ_ERR_INVALID, "bad argument");
    GJ_CUDA_TRY(cudaSetDevice(p->device));
    return gj_launch_score_plain_i32(p, d_samples, row_stride, S, d_scores, false, (cudaStream_t)stream);
}

extern "C" gj_status gj_score_incremental_device(gj_problem* p, const double* d_base,
                                                 const uint64_t* d_offsets, const uint64_t* d_var_ids,
                                                 const double* d_values, int64_t S, double* d_scores,
                                                 void* stream) {
    if (!p || !d_base || !d_offsets || !d_scores || S < 0) return gj_fail(GJ_ERR_INVALID, "bad argument");
    GJ_CUDA_TRY(cudaSetDevice(p->device));
    gj_status rc;
    if ((rc = p->d_base_i32.reserve((size_t)p->dev.n_vars * 4))) return rc;
    return gj_launch_score_incremental(p, d_base, (int32_t*)p->d_base_i32.ptr, d_offsets, d_var_ids,
                                       d_values, S, d_scores, (cudaStream_t)stream);
}

extern "C" gj_status gj_score_plain(gj_problem* p, const double* samples, int64_t S, double* scores) {
    if (!p || !samples || !scores || S < 0) return gj_fail(GJ_ERR_INVALID, "bad argument");
    if (S == 0) return GJ_OK;
    GJ_CUDA_TRY(cudaSetDevice(p->device));
    const size_t in_bytes = (size_t)S * (size_t)p->dev.n_vars * 8;
    const size_t out_bytes = (size_t)S * (size_t)p->dev.levels * 8;
    gj_status rc;
    if ((rc = p->d_samples.reserve(in_bytes))) return rc;
    if ((rc = p->d_scores.reserve(out_bytes))) return rc;
    GJ_CUDA_TRY(cudaMemcpyAsync(p->d_samples.ptr, samples, in_bytes, cudaMemcpyHostToDevice, p->stream));
    if ((rc = gj_launch_score_plain_f64(p, (const double*)p->d_samples.ptr, S, (double*)p->d_scores.ptr, false, p->stream))) return rc;
    GJ_CUDA_TRY(cudaMemcpyAsync(scores, p->d_scores.ptr, out_bytes, cudaMemcpyDeviceToHost, p->stream));
    GJ_CUDA_TRY(cudaStreamSynchronize(p->stream));
    return GJ_OK;
}

extern "C" gj_status gj_score_incremental(gj_problem* p, const double* base, const uint64_t* offsets,
                                          const uint64_t* var_ids, const double* values, int64_t S,
                                          double* scores) {
    if (!p || !base || !offsets || !scores || S < 0) return gj_fail(GJ_ERR_INVALID, "bad argument");
    if (S == 0) return GJ_OK;
    const uint64_t total = offsets[S];
    if (total > 0 && (!var_ids || !values)) return gj_fail(GJ_ERR_INVALID, "delta arrays missing");
    GJ_CUDA_TRY(cudaSetDevice(p->device));
    const size_t out_bytes = (size_t)S * (size_t)p->dev.levels * 8;
    gj_status rc;
    if ((rc = p->d_base.reserve((size_t)p->dev.n_vars * 8))) return rc;
    if ((rc = p->d_base_i32.reserve((size_t)p->dev.n_vars * 4))) return rc;
    if ((rc = p->d_offsets.reserve((size_t)(S + 1) * 8))) return rc;
    if ((rc = p->d_ids.reserve((size_t)(total + 1) * 8))) return rc;
    if ((rc = p->d_vals.reserve((size_t)(total + 1) * 8))) return rc;
    if ((rc = p->d_scores.reserve(out_bytes))) return rc;
    cudaStream_t st = p->stream;
    GJ_CUDA_TRY(cudaMemcpyAsync(p->d_base.ptr, base, (size_t)p->dev.n_vars * 8, cudaMemcpyHostToDevice, st));
    GJ_CUDA_TRY(cudaMemcpyAsync(p->d_offsets.ptr, offsets, (size_t)(S + 1) * 8, cudaMemcpyHostToDevice, st));
    if (total) {
        GJ_CUDA_TRY(cudaMemcpyAsync(p->d_ids.ptr, var_ids, (size_t)total * 8, cudaMemcpyHostToDevice, st));
        GJ_CUDA_TRY(cudaMemcpyAsync(p->d_vals.ptr, values, (size_t)total * 8, cudaMemcpyHostToDevice, st));
    }
    if ((rc = gj_launch_score_incremental(p, (const double*)p->d_base.ptr, (int32_t*)p->d_base_i32.ptr,
                                          (const uint64_t*)p->d_offsets.ptr, (const uint64_t*)p->d_ids.ptr,
                                          (const double*)p->d_vals.ptr, S, (double*)p->d_scores.ptr, st))) return rc;
    GJ_CUDA_TRY(cudaMemcpyAsync(scores, p->d_scores.ptr, out_bytes, cudaMemcpyDeviceToHost, st));
    GJ_CUDA_TRY(cudaStreamSynchronize(st));
    return GJ_OK;
}

// Packed wire format of the same call: 8 bytes per delta instead of 16.  The Rust shim that flattens
// Vec<Vec<(usize, f64)>> into CSR walks every pair anyway; narrowing the column id to u32 and the
// (already inverse-transformed, integer) value to i32 in that loop halves the bytes that cross PCIe,
// which is what bounds this call.
extern "C" gj_status gj_score_incremental_packed(gj_problem* p, const double* base, const uint64_t* offsets,
                                                 const uint32_t* var_ids, const int32_t* values, int64_t S,
                                                 double* scores) {
    if (!p || !base || !offsets || !scores || S < 0) return gj_fail(GJ_ERR_INVALID, "bad argument");
    if (S == 0) return GJ_OK;
    const uint64_t total = offsets[S];
    if (total > 0 && (!var_ids || !values)) return gj_fail(GJ_ERR_INVALID, "delta arrays missing");
    GJ_CUDA_TRY(cudaSetDevice(p->device));
    const size_t out_bytes = (size_t)S * (size_t)p->dev.levels * 8;
    gj_status rc;
    if ((rc = p->d_base.reserve((size_t)p->dev.n_vars * 8))) return rc;
    if ((rc = p->d_base_i32.reserve((size_t)p->dev.n_vars * 4))) return rc;
    if ((rc = p->d_offsets.reserve((size_t)(S + 1) * 8))) return rc;
    if ((rc = p->d_ids.reserve((size_t)(total + 1) * 4))) return rc;
    if ((rc = p->d_vals.reserve((size_t)(total + 1) * 4))) return rc;
    if ((rc = p->d_scores.reserve(out_bytes))) return rc;
    cudaStream_t st = p->stream;
    GJ_CUDA_TRY(cudaMemcpyAsync(p->d_base.ptr, base, (size_t)p->dev.n_vars * 8, cudaMemcpyHostToDevice, st));
    GJ_CUDA_TRY(cudaMemcpyAsync(p->d_offsets.ptr, offsets, (size_t)(S + 1) * 8, cudaMemcpyHostToDevice, st));
    if (total) {
        GJ_CUDA_TRY(cudaMemcpyAsync(p->d_ids.ptr, var_ids, (size_t)total * 4, cudaMemcpyHostToDevice, st));
        GJ_CUDA_TRY(cudaMemcpyAsync(p->d_vals.ptr, values, (size_t)total * 4, cudaMemcpyHostToDevice, st));
    }
    if ((rc = launch_incremental<uint32_t, int32_t>(p, (const double*)p->d_base.ptr, (int32_t*)p->d_base_i32.ptr,
                                                    (const uint64_t*)p->d_offsets.ptr, (const uint32_t*)p->d_ids.ptr,
                                                    (const int32_t*)p->d_vals.ptr, S, (double*)p->d_scores.ptr, st))) return rc;
    GJ_CUDA_TRY(cudaMemcpyAsync(scores, p->d_scores.ptr, out_bytes, cudaMemcpyDeviceToHost, st));
    GJ_CUDA_TRY(cudaStreamSynchronize(st));
    return GJ_OK;
}
