// gj_problem.cu -- problem upload: the device-resident form of the reference's
// Cotwin + VariablesManager + scorer utility objects (cotwin/cotwin.rs:12-18,
// score_requesters/variables_manager.rs:10-74, examples' UtilityObjectVariants).
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <algorithm>

#include "gj_eval.cuh"
#include "gj_internal.hpp"

static thread_local std::string g_last_error;

void gj_set_error(const std::string& msg) { g_last_error = msg; }
gj_status gj_fail(gj_status code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

extern "C" const char* gj_last_error(void) { return g_last_error.c_str(); }
extern "C" int32_t gj_abi_version(void) { return 1; }
std::atomic<long long> gj_launch_counter{0};
extern "C" int64_t gj_launch_count(void) { return (int64_t)gj_launch_counter.load(); }
extern "C" size_t gj_sizeof_problem_desc(void) { return sizeof(gj_problem_desc); }
extern "C" size_t gj_sizeof_agent_params(void) { return sizeof(gj_agent_params); }
extern "C" int32_t gj_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

gj_status GjBuffer::reserve(size_t need) {
    if (need <= bytes) return GJ_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; bytes = 0;
    size_t cap = need + need / 4 + 256;
    cudaError_t e = cudaMalloc(&ptr, cap);
    if (e != cudaSuccess) return gj_fail(GJ_ERR_OOM, std::string("cudaMalloc scratch: ") + cudaGetErrorString(e));
    bytes = cap;
    return GJ_OK;
}
void GjBuffer::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr; bytes = 0;
}

gj_problem::~gj_problem() {
    cudaSetDevice(device);
    for (void* a : allocs) cudaFree(a);
    d_samples.release(); d_scores.release(); d_base.release(); d_base_i32.release();
    d_offsets.release(); d_ids.release(); d_vals.release();
    if (stream) cudaStreamDestroy(stream);
}

template <class T>
static gj_status upload(gj_problem* p, const T* host, size_t n, const T** out) {
    void* d = nullptr;
    size_t bytes = (n ? n : 1) * sizeof(T);
    GJ_CUDA_TRY(cudaMalloc(&d, bytes));
    p->allocs.push_back(d);
    if (n) GJ_CUDA_TRY(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)d;
    return GJ_OK;
}

// examples/tsp/src/domain/location.rs:38-50: round(sqrt(dlat^2 + dlon^2), 3), then
// rounded once more by examples/tsp/src/persistence/domain_builder.rs:42-46.  IEEE
// sqrt / mul / add without contraction (--fmad=false) are bit-identical to the CPU.
__global__ void gj_build_distance_matrix_kernel(const double* __restrict__ xy, int n,
                                                double* __restrict__ D) {
    const size_t total = (size_t)n * (size_t)n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / (size_t)n), j = (int)(idx % (size_t)n);
        const double dlat = xy[2 * j] - xy[2 * i];
        const double dlon = xy[2 * j + 1] - xy[2 * i + 1];
        const double a = __dmul_rn(dlat, dlat);
        const double b = __dmul_rn(dlon, dlon);
        double d = __dsqrt_rn(__dadd_rn(a, b));
        d = gj_round_mult(d, 1000.0);
        d = gj_round_mult(d, 1000.0);
        D[idx] = d;
    }
}

// 1 if D[i][j] == D[j][i] bitwise for all i, j (lets 2-opt deltas skip the interior).
__global__ void gj_check_symmetric_kernel(const double* __restrict__ D, int n, int* flag) {
    const size_t total = (size_t)n * (size_t)n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t i = idx / (size_t)n, j = idx % (size_t)n;
        if (i < j) {
            if (__double_as_longlong(D[idx]) != __double_as_longlong(D[j * (size_t)n + i]))
                atomicExch(flag, 0);
        }
    }
}

// Fixed-point copy of the matrix: D32[i][j] = rint(D[i][j] * 1000).  The examples truncate every
// distance to 3 decimals (location.rs:42-47), so the matrix is a table of milli-units; `flag` drops
// to 0 if an entry is not within 1e-6 of a whole milli-unit or too large for sums of four in int32.
__global__ void gj_build_d32_kernel(const double* __restrict__ D, size_t total, int32_t* __restrict__ D32, int* flag) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const double x = D[idx] * 1000.0;
        const double r = rint(x);
        if (!(fabs(x - r) <= 1e-6) || !(r >= 0.0) || r > 268435456.0) atomicExch(flag, 0);
        D32[idx] = (int32_t)fmin(fmax(r, 0.0), 268435456.0);
    }
}

// Builds (once) the milli-unit matrix; p->d32_state: 0 = not tried, 1 = valid, -1 = the matrix is not
// a table of milli-units (delta scoring then stays in f64).
gj_status gj_problem_ensure_d32(gj_problem* p) {
    if (p->d32_state != 0) return GJ_OK;
    p->d32_state = -1;
    if (!p->dev.D || p->dev.n_locations <= 0) return GJ_OK;
    GJ_CUDA_TRY(cudaSetDevice(p->device));
    const size_t total = (size_t)p->dev.n_locations * (size_t)p->dev.n_locations;
    int32_t* d32 = nullptr;
    int* flag = nullptr;
    if (cudaMalloc(&d32, total * 4) != cudaSuccess) { cudaGetLastError(); return GJ_OK; }   // no room: stay in f64
    p->allocs.push_back(d32);
    GJ_CUDA_TRY(cudaMalloc(&flag, 4));
    int one = 1;
    GJ_CUDA_TRY(cudaMemcpy(flag, &one, 4, cudaMemcpyHostToDevice));
    gj_build_d32_kernel<<<148 * 8, 256, 0, p->stream>>>(p->dev.D, total, d32, flag);
    GJ_LAUNCH_CHECK();
    GJ_CUDA_TRY(cudaStreamSynchronize(p->stream));
    GJ_CUDA_TRY(cudaMemcpy(&one, flag, 4, cudaMemcpyDeviceToHost));
    cudaFree(flag);
    if (one) { p->dev.D32 = d32; p->d32_state = 1; }
    return GJ_OK;
}

static bool fits_i32(double x) { return x >= -2147483647.0 && x <= 2147483647.0 && x == x; }

// host restatement of GJInteger::fix on a bound (only used to pre-round the bounds)
static double host_rint(double x) {
    double f = std::floor(x), c = std::ceil(x);
    return (std::fabs(x - f) < std::fabs(c - x)) ? f : c;
}

extern "C" gj_status gj_problem_create(const gj_problem_desc* desc, int32_t device,
                                       gj_problem** out) {
    if (!desc || !out) return gj_fail(GJ_ERR_INVALID, "null argument");
    *out = nullptr;
    if (desc->kind < GJ_NQUEENS || desc->kind > GJ_VRP_SERVICE)
        return gj_fail(GJ_ERR_INVALID, "unknown problem kind");
    if (desc->n_vars <= 0 || !desc->lower_bounds || !desc->upper_bounds)
        return gj_fail(GJ_ERR_INVALID, "n_vars / bounds missing");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return gj_fail(GJ_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return gj_fail(GJ_ERR_INVALID, "device out of range");
    GJ_CUDA_TRY(cudaSetDevice(device));

    const int n = desc->n_vars;
    const bool is_vrp = desc->kind >= GJ_VRP;
    if (is_vrp && (n % 2) != 0) return gj_fail(GJ_ERR_INVALID, "VRP needs 2 variables per stop");

    std::unique_ptr<gj_problem> p(new gj_problem());
    p->device = device;
    GJ_CUDA_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    p->lb.assign(desc->lower_bounds, desc->lower_bounds + n);
    p->ub.assign(desc->upper_bounds, desc->upper_bounds + n);
    p->frozen.assign(n, 0);
    if (desc->frozen) p->frozen.assign(desc->frozen, desc->frozen + n);
    p->initial.assign(n, std::numeric_limits<double>::quiet_NaN());
    if (desc->initial) p->initial.assign(desc->initial, desc->initial + n);
    for (int l = 0; l < 3; ++l) p->precision[l] = desc->score_precision[l];

    std::vector<int32_t> lbi(n), ubi(n);
    for (int i = 0; i < n; ++i) {
        if (!fits_i32(p->lb[i]) || !fits_i32(p->ub[i]) || p->lb[i] > p->ub[i])
            return gj_fail(GJ_ERR_INVALID, "variable bounds must be ordered and fit int32");
        lbi[i] = (int32_t)host_rint(p->lb[i]);
        ubi[i] = (int32_t)host_rint(p->ub[i]);
        if (p->frozen[i]) {
            // gj_integer.rs:72-75: "Frozen value must be initialized"
            if (!(p->initial[i] == p->initial[i]) || !fits_i32(p->initial[i]))
                return gj_fail(GJ_ERR_INVALID, "Frozen value must be initialized");
        }
    }

    // semantic groups (variables_manager.rs:76-106): frozen variables are skipped
    if (desc->n_groups > 0) {
        if (!desc->group_offsets || !desc->group_var_ids)
            return gj_fail(GJ_ERR_INVALID, "semantic groups missing");
        for (int g = 0; g < desc->n_groups; ++g) {
            std::vector<int32_t> ids;
            for (int64_t k = desc->group_offsets[g]; k < desc->group_offsets[g + 1]; ++k) {
                int32_t v = desc->group_var_ids[k];
                if (v < 0 || v >= n) return gj_fail(GJ_ERR_INVALID, "group variable id out of range");
                if (!p->frozen[v]) ids.push_back(v);
            }
            p->groups.push_back(std::move(ids));
        }
    } else {
        std::vector<int32_t> ids;                     // GJInteger default group "common"
        for (int i = 0; i < n; ++i) if (!p->frozen[i]) ids.push_back(i);
        p->groups.push_back(std::move(ids));
    }

    GjProblemDev& P = p->dev;
    P.kind = desc->kind;
    P.n_vars = n;
    P.levels = (desc->kind == GJ_NQUEENS) ? 1 : (desc->kind == GJ_TSP ? 2 : 3);
    P.n_entities = is_vrp ? n / 2 : n;
    for (int i = 0; i < 4; ++i) P.w[i] = desc->weights[i];
    P.exact_sums = 1;
    for (int l = 0; l < GJ_MAX_LEVELS; ++l) {
        P.round_mult[l] = 0.0;
        if (l < P.levels && desc->score_precision[l] >= 0)
            P.round_mult[l] = std::pow(10.0, (double)desc->score_precision[l]);
    }

    gj_status st;
    if ((st = upload(p.get(), p->lb.data(), n, &P.lb))) return st;
    if ((st = upload(p.get(), p->ub.data(), n, &P.ub))) return st;
    if ((st = upload(p.get(), p->frozen.data(), n, &P.frozen))) return st;
    {
        std::vector<double> init_dev(p->initial);
        for (auto& x : init_dev) if (!(x == x)) x = 0.0;
        if ((st = upload(p.get(), init_dev.data(), n, &P.initial))) return st;
    }
    if ((st = upload(p.get(), lbi.data(), n, &P.lbi))) return st;
    if ((st = upload(p.get(), ubi.data(), n, &P.ubi))) return st;

    // value range of the column whose distinct values are counted
    auto col_range = [&](int first, int stride, int& lo, int& hi) {
        lo = INT32_MAX; hi = INT32_MIN;
        for (int i = first; i < n; i += stride) {
            int a = lbi[i], b = ubi[i];
            if (p->frozen[i]) { a = b = (int32_t)p->initial[i]; }
            lo = std::min(lo, a); hi = std::max(hi, b);
        }
    };
    int vlo, vhi;
    if (is_vrp) col_range(1, 2, vlo, vhi); else col_range(0, 1, vlo, vhi);
    P.val_lo = vlo;
    P.bm_words = (int)(((int64_t)vhi - vlo + 1 + 31) / 32);

    if (desc->kind == GJ_NQUEENS) {
        std::vector<int32_t> col(n);
        int cmin = INT32_MAX, cmax = INT32_MIN;
        for (int i = 0; i < n; ++i) {
            int64_t c = desc->column_id ? desc->column_id[i] : i;
            if (c < -1000000000LL || c > 1000000000LL) return gj_fail(GJ_ERR_INVALID, "column_id out of range");
            col[i] = (int32_t)c;
            cmin = std::min(cmin, col[i]); cmax = std::max(cmax, col[i]);
        }
        if ((st = upload(p.get(), col.data(), n, &P.column_id))) return st;
        P.desc_lo = cmin + vlo;
        P.desc_words = (int)(((int64_t)(cmax + vhi) - P.desc_lo + 1 + 31) / 32);
        P.asc_lo = cmin - vhi;
        P.asc_words = (int)(((int64_t)(cmax - vlo) - P.asc_lo + 1 + 31) / 32);
    }

    if (desc->kind != GJ_NQUEENS) {
        const int L = desc->n_locations;
        if (L <= 0) return gj_fail(GJ_ERR_INVALID, "n_locations missing");
        if (vlo < 0 || vhi >= L) return gj_fail(GJ_ERR_INVALID, "location ids must lie in [0, n_locations)");
        P.n_locations = L;
        double* dD = nullptr;
        const size_t LL = (size_t)L * (size_t)L;
        GJ_CUDA_TRY(cudaMalloc((void**)&dD, LL * sizeof(double)));
        p->allocs.push_back(dD);
        if (desc->distance_matrix) {
            GJ_CUDA_TRY(cudaMemcpy(dD, desc->distance_matrix, LL * sizeof(double), cudaMemcpyHostToDevice));
        } else if (desc->coords) {
            const double* dxy = nullptr;
            if ((st = upload(p.get(), desc->coords, (size_t)L * 2, &dxy))) return st;
            gj_build_distance_matrix_kernel<<<148 * 8, 256>>>(dxy, L, dD);
            GJ_LAUNCH_CHECK();
        } else {
            return gj_fail(GJ_ERR_INVALID, "distance_matrix or coords required");
        }
        P.D = dD;
        int* dflag = nullptr;
        int one = 1;
        GJ_CUDA_TRY(cudaMalloc((void**)&dflag, sizeof(int)));
        GJ_CUDA_TRY(cudaMemcpy(dflag, &one, sizeof(int), cudaMemcpyHostToDevice));
        gj_check_symmetric_kernel<<<148 * 8, 256>>>(dD, L, dflag);
        GJ_LAUNCH_CHECK();
        GJ_CUDA_TRY(cudaMemcpy(&one, dflag, sizeof(int), cudaMemcpyDeviceToHost));
        cudaFree(dflag);
        p->symmetric_D = one != 0;
    }

    if (is_vrp) {
        const int K = desc->n_vehicles, L = desc->n_locations;
        if (K <= 0 || K > 65535) return gj_fail(GJ_ERR_INVALID, "n_vehicles must be in [1, 65535]");
        if (!desc->vehicle_depot || !desc->vehicle_capacity || !desc->demand)
            return gj_fail(GJ_ERR_INVALID, "vehicle / customer facts missing");
        int vehlo, vehhi;
        col_range(0, 2, vehlo, vehhi);
        if (vehlo < 0 || vehhi >= K) return gj_fail(GJ_ERR_INVALID, "vehicle ids must lie in [0, n_vehicles)");
        P.veh_lo = 0;
        P.n_vehicles = K;
        P.time_windowed = desc->time_windowed ? 1 : 0;
        std::vector<int32_t> depot(K);
        std::vector<unsigned long long> cap(K), ds(K, 0), de(K, 0);
        for (int v = 0; v < K; ++v) {
            if (desc->vehicle_depot[v] < 0 || desc->vehicle_depot[v] >= L)
                return gj_fail(GJ_ERR_INVALID, "vehicle depot out of range");
            depot[v] = (int32_t)desc->vehicle_depot[v];
            cap[v] = desc->vehicle_capacity[v];
            if (desc->work_day_start) ds[v] = desc->work_day_start[v];
            if (desc->work_day_end) de[v] = desc->work_day_end[v];
        }
        std::vector<uint4> cust(L);
        for (int c = 0; c < L; ++c) {
            uint64_t d = desc->demand[c];
            uint64_t a = desc->tw_start ? desc->tw_start[c] : 0;
            uint64_t b = desc->tw_end ? desc->tw_end[c] : 0;
            uint64_t s = desc->service_time ? desc->service_time[c] : 0;
            if ((d | a | b | s) >> 32)
                return gj_fail(GJ_ERR_UNSUPPORTED, "customer facts must fit in 32 bits");
            cust[c] = make_uint4((unsigned)d, (unsigned)a, (unsigned)b, (unsigned)s);
        }
        if (P.time_windowed && (!desc->tw_start || !desc->tw_end || !desc->service_time ||
                                !desc->work_day_start || !desc->work_day_end))
            return gj_fail(GJ_ERR_INVALID, "time_windowed needs tw_start/tw_end/service_time/work day");
        if ((st = upload(p.get(), depot.data(), K, &P.veh_depot))) return st;
        if ((st = upload(p.get(), cap.data(), K, &P.veh_capacity))) return st;
        if ((st = upload(p.get(), ds.data(), K, &P.day_start))) return st;
        if ((st = upload(p.get(), de.data(), K, &P.day_end))) return st;
        if ((st = upload(p.get(), cust.data(), L, &P.cust))) return st;
        size_t smem = gj_vrp_smem_bytes(P.n_entities, K, P.bm_words, kVrpWarps, !P.time_windowed);
        if (smem > 220 * 1024)
            return gj_fail(GJ_ERR_UNSUPPORTED, "VRP instance too large for the shared-memory route sort");
        if (L > 65536)
            return gj_fail(GJ_ERR_UNSUPPORTED, "VRP models are limited to 65536 locations (16-bit ids in the route sort)");
    }
    GJ_CUDA_TRY(cudaDeviceSynchronize());
    *out = p.release();
    return GJ_OK;
}

extern "C" void gj_problem_destroy(gj_problem* p) { delete p; }
extern "C" int32_t gj_problem_levels(const gj_problem* p) { return p ? p->dev.levels : 0; }
extern "C" int32_t gj_problem_n_vars(const gj_problem* p) { return p ? p->dev.n_vars : 0; }

extern "C" gj_status gj_problem_set_constraint_weights(gj_problem* p, const double* w, int32_t n) {
    if (!p || !w || n < 1 || n > 4) return gj_fail(GJ_ERR_INVALID, "bad weights");
    for (int i = 0; i < n; ++i) p->dev.w[i] = w[i];
    return GJ_OK;
}

extern "C" gj_status gj_problem_set_exact_sums(gj_problem* p, int32_t on) {
    if (!p) return gj_fail(GJ_ERR_INVALID, "null handle");
    p->dev.exact_sums = on ? 1 : 0;
    return GJ_OK;
}

extern "C" gj_status gj_problem_get_distance_matrix(gj_problem* p, double* out) {
    if (!p || !out || !p->dev.D) return gj_fail(GJ_ERR_INVALID, "no distance matrix");
    GJ_CUDA_TRY(cudaSetDevice(p->device));
    size_t LL = (size_t)p->dev.n_locations * (size_t)p->dev.n_locations;
    GJ_CUDA_TRY(cudaMemcpy(out, p->dev.D, LL * sizeof(double), cudaMemcpyDeviceToHost));
    return GJ_OK;
}

extern "C" gj_status gj_host_alloc(size_t bytes, void** out) {
    if (!out) return gj_fail(GJ_ERR_INVALID, "null out");
    GJ_CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return GJ_OK;
}
extern "C" void gj_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}
