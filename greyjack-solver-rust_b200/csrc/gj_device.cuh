// gj_device.cuh -- device-side building blocks shared by every kernel of the engine:
// variable decoding, score ordering / rounding, warp reductions, Philox4x32-10.
// Compiled for sm_100a only, with --fmad=false so that no a*b+c is contracted
// (the Rust reference never fuses; SURVEY.md section 7 "FMA contraction").
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define GJ_FULL_MASK 0xffffffffu
#define GJ_MAX_LEVELS 3

// Everything a scoring kernel needs to know about the problem; passed by value
// (lives in the kernel parameter / constant bank).  Built by gj_problem.cu.
struct GjProblemDev {
    int kind;
    int n_vars;
    int levels;
    int n_entities;              // queens / stops (VRP: n_vars / 2)

    // planning variables (variables/gj_integer.rs)
    const double* lb;            // [n_vars]
    const double* ub;            // [n_vars]
    const uint8_t* frozen;       // [n_vars]
    const double* initial;       // [n_vars]
    const int32_t* lbi;          // decoded integer bounds (rint'ed), [n_vars]
    const int32_t* ubi;

    // bitmap geometry for distinct counting (values are clamped into [val_lo, val_hi])
    int val_lo;                  // min decoded value of the counted column
    int bm_words;                // 32-bit words covering the counted column's range
    // N-Queens diagonals
    const int32_t* column_id;    // [n_vars]
    int desc_lo, desc_words;     // col + row
    int asc_lo, asc_words;       // col - row

    // utility objects
    int n_locations;
    const double* D;             // [L][L] row-major distance matrix
    const int32_t* D32;          // [L][L] the same in milli-units (rint(D * 1000)) when every entry is a
                                 // whole number of them (the examples truncate to 3 decimals); else nullptr
    int n_vehicles;
    const int32_t* veh_depot;    // [K]
    const unsigned long long* veh_capacity;
    const unsigned long long* day_start;
    const unsigned long long* day_end;
    const uint4* cust;           // [L] {demand, tw_start, tw_end, service_time}
    int time_windowed;
    int veh_lo;                  // min decoded vehicle id (0 in the examples)

    int exact_sums;              // 1: reference summation order (bit-exact float level)
    double w[4];                 // constraint weights
    double round_mult[GJ_MAX_LEVELS];   // 10^precision, or 0 = None
};

// ---- f64 helpers -----------------------------------------------------------------

// Rust f64::total_cmp
__device__ __forceinline__ int gj_total_cmp(double a, double b) {
    long long l = __double_as_longlong(a), r = __double_as_longlong(b);
    l ^= (long long)(((unsigned long long)(l >> 63)) >> 1);
    r ^= (long long)(((unsigned long long)(r >> 63)) >> 1);
    return (l < r) ? -1 : ((l > r) ? 1 : 0);
}

// greyjack/src/utils/math_utils.rs:6-8 (ties -> ceil)
__device__ __forceinline__ double gj_rint(double x) {
    double f = floor(x), c = ceil(x);
    return (fabs(x - f) < fabs(c - x)) ? f : c;
}

// greyjack/src/utils/math_utils.rs:10-13 with multiplier = 10^precision
__device__ __forceinline__ double gj_round_mult(double v, double mult) {
    double fl = floor(v);
    return fl + floor((v - fl) * mult) / mult;
}

// greyjack/src/variables/gj_integer.rs:70-83, 114-138 (unfrozen part)
__device__ __forceinline__ double gj_fix_integer(double v, double lb, double ub) {
    int c = gj_total_cmp(v, lb);
    double m = (c > 0) ? v : lb;                 // max: Less->b, Greater->a, Equal->b
    c = gj_total_cmp(m, ub);
    double r = (c > 0) ? ub : m;                 // min: Less->a, Greater->b, Equal->a
    return gj_rint(r);
}

// GJInteger::inverse_transform (gj_integer.rs:66-68) narrowed to int32 (bounds are
// validated to fit at problem creation).
__device__ __forceinline__ int gj_decode(const GjProblemDev& P, int i, double x) {
    if (P.frozen[i]) return (int)P.initial[i];
    return (int)gj_fix_integer(x, P.lb[i], P.ub[i]);
}

// ---- scores ------------------------------------------------------------------------

struct GjScore {
    double v[GJ_MAX_LEVELS];
};

// Ord::cmp of the score structs (lexicographic total_cmp; lower is better)
__device__ __forceinline__ int gj_score_cmp(const GjScore& a, const GjScore& b, int levels) {
    for (int l = 0; l < levels; ++l) {
        int c = gj_total_cmp(a.v[l], b.v[l]);
        if (c != 0) return c;
    }
    return 0;
}

// derived PartialOrd `a <= b` (every acceptance test of the reference)
__device__ __forceinline__ bool gj_score_le(const GjScore& a, const GjScore& b, int levels) {
    for (int l = 0; l < levels; ++l) {
        if (a.v[l] != a.v[l] || b.v[l] != b.v[l]) return false;
        if (a.v[l] < b.v[l]) return true;
        if (a.v[l] > b.v[l]) return false;
    }
    return true;
}

// ScoreTrait::round applied after scoring (agent_base.rs:284-287, 311-314)
__device__ __forceinline__ void gj_score_round(GjScore& s, const GjProblemDev& P) {
    for (int l = 0; l < P.levels; ++l)
        if (P.round_mult[l] != 0.0) s.v[l] = gj_round_mult(s.v[l], P.round_mult[l]);
}

// ---- SimulatedAnnealing acceptance ------------------------------------------------------------
struct GjSaParams {
    int has_cooling;            // cooling_rate: Option<f64>
    double cooling_rate;
    double inv_rate;            // 1 - termination_strategy.get_accomplish_rate() (agent_base.rs:544)
};

// SimulatedAnnealingBase::build_updated_population_incremental
// (metaheuristic_bases/simulated_annealing_base.rs:198-233): temperature update per level, then
// accept iff u < prod_l e^(-(candidate_l - current_l) / T_l).  temp: [levels], updated in place.
__device__ __forceinline__ bool gj_sa_accept(const GjScore& cand, const GjScore& cur, int levels,
                                             double* temp, const GjSaParams& sa, double u, double* proba_out) {
    double proba = 1.0;
    for (int l = 0; l < levels; ++l) {
        double t = temp[l];
        if (sa.has_cooling) {
            t = t * sa.cooling_rate;
            if (t < 0.000001) t = 0.0000001;
        } else {
            t = sa.inv_rate;
        }
        temp[l] = t;
        proba = proba * pow(2.7182818284590452, -((cand.v[l] - cur.v[l]) / t));
    }
    if (proba_out) *proba_out = proba;
    return u < proba;
}

// ---- warp primitives ---------------------------------------------------------------

__device__ __forceinline__ double gj_warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(GJ_FULL_MASK, x, o);
    return x;
}
__device__ __forceinline__ int gj_warp_sum(int x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(GJ_FULL_MASK, x, o);
    return x;
}
__device__ __forceinline__ long long gj_warp_sum(long long x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(GJ_FULL_MASK, x, o);
    return x;
}

// CTA-wide sums (every thread gets the total).  `scratch` holds 32 entries.
__device__ __forceinline__ int gj_block_sum(int x, int* scratch) {
    x = gj_warp_sum(x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = x;
    __syncthreads();
    int tot = 0;
    for (int w = 0; w < nw; ++w) tot += scratch[w];
    return tot;
}
__device__ __forceinline__ double gj_block_sum(double x, double* scratch) {
    x = gj_warp_sum(x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = x;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < nw; ++w) tot += scratch[w];
    return tot;
}

// CTA-wide inclusive prefix sum (every thread gets the sum of x over threads 0..tid; *total = the CTA's
// sum).  `warp_tot` holds 33 ints of shared memory.  Two barriers, whatever the CTA size.
__device__ __forceinline__ int gj_block_scan_incl(int x, int* warp_tot, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(GJ_FULL_MASK, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();                       // warp_tot may still be read from a previous call
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nw ? warp_tot[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(GJ_FULL_MASK, w, o);
            if (lane >= o) w += y;
        }
        warp_tot[lane] = w;                // inclusive over warps
        if (lane == 31) warp_tot[32] = w;
    }
    __syncthreads();
    if (warp > 0) x += warp_tot[warp - 1];
    if (total) *total = warp_tot[32];
    return x;
}

// ---- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier ------------------------------
// Stages a contiguous table (a solution row, a tabu table, a fact table) from global to shared
// memory with one instruction issued by one thread; consumers wait on the mbarrier's phase.
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ uint32_t gj_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void gj_mbar_init(uint64_t* mbar, int arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gj_smem_u32(mbar)), "r"(arrivals));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void gj_mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gj_smem_u32(mbar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void gj_tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                               uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(gj_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(gj_smem_u32(mbar))
                 : "memory");
}
// Pulls a contiguous global range into L2 without a destination (cp.async.bulk.prefetch.L2): turns the
// first touches of a gather table that another kernel evicted into one streamed read.  16-byte
// aligned address, bytes a multiple of 16.
__device__ __forceinline__ void gj_l2_prefetch_bulk(const void* gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gj_mbar_wait(uint64_t* mbar, uint32_t phase) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n"
                     "}\n"
                     : "=r"(done)
                     : "r"(gj_smem_u32(mbar)), "r"(phase)
                     : "memory");
    }
}

// ---- Philox4x32-10 counter-based RNG -----------------------------------------------
// key = (seed_lo ^ island, seed_hi); counter = (step, candidate, stream, 0): moves are a
// pure function of (seed, island, step, candidate) and never visit the host.

struct GjPhilox {
    uint32_t c[4];
    uint32_t k[2];
    uint32_t out[4];
    int have;
    int compact;      // 1: refill through ONE out-of-line Philox block (latency-bound kernels where the
                      // instruction footprint matters more than the call)
};

__device__ __forceinline__ void gj_philox_round(uint32_t* c, const uint32_t* k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k[0];
    uint32_t n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__device__ __forceinline__ void gj_philox_block(const uint32_t* ctr, const uint32_t* key,
                                                uint32_t* out) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        gj_philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

static __device__ __noinline__ uint4 gj_philox_block_call(uint4 c, uint2 k) {
    const uint32_t cc[4] = {c.x, c.y, c.z, c.w}, kk[2] = {k.x, k.y};
    uint32_t o[4];
    gj_philox_block(cc, kk, o);
    return make_uint4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ void gj_rng_init(GjPhilox& g, uint64_t seed, uint32_t island,
                                            uint32_t step_lo, uint32_t step_hi,
                                            uint32_t candidate) {
    g.k[0] = (uint32_t)seed ^ (island * 0x9E3779B1u);
    g.k[1] = (uint32_t)(seed >> 32) + island;
    g.c[0] = step_lo; g.c[1] = candidate; g.c[2] = step_hi; g.c[3] = 0;
    g.have = 0;
    g.compact = 0;
}

__device__ __forceinline__ uint32_t gj_rng_u32(GjPhilox& g) {
    if (g.have == 0) {
        if (g.compact) {
            const uint4 o = gj_philox_block_call(make_uint4(g.c[0], g.c[1], g.c[2], g.c[3]), make_uint2(g.k[0], g.k[1]));
            g.out[0] = o.x; g.out[1] = o.y; g.out[2] = o.z; g.out[3] = o.w;
        } else {
            gj_philox_block(g.c, g.k, g.out);
        }
        g.c[3] += 1;
        g.have = 4;
    }
    // static indices only: the state stays in registers (out[3] is handed out first)
    g.have -= 1;
    uint32_t r = g.out[0];
    if (g.have == 1) r = g.out[1];
    if (g.have == 2) r = g.out[2];
    if (g.have == 3) r = g.out[3];
    return r;
}

// uniform integer in [0, n) (n > 0); multiply-shift, bias < n / 2^32
__device__ __forceinline__ uint32_t gj_rng_below(GjPhilox& g, uint32_t n) {
    return __umulhi(gj_rng_u32(g), n);
}

// uniform double in [0, 1)
__device__ __forceinline__ double gj_rng_f64(GjPhilox& g) {
    uint64_t hi = gj_rng_u32(g), lo = gj_rng_u32(g);
    uint64_t bits = ((hi << 32) | lo) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
}
