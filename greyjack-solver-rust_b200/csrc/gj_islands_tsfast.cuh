// gj_islands_tsfast.cuh -- the TabuSearch island step for TSP in FIXED POINT (included by
// gj_islands_tsfast.cu).  Same phases and the same observable behaviour as k_ls_step_fused
// (gj_islands_fused.cuh); what changes is the arithmetic of the hot loop:
//
//   * The examples' distance matrix is truncated to 3 decimals (location.rs:42-47), i.e. it is a table
//     of milli-units.  gj_problem keeps an int32 copy D32 = rint(D * 1000) (validated entry by entry),
//     and a neighbour's tour-length change is an exact integer: (sum of <= 4 added edges) - (sum of
//     <= 4 removed edges).  ScoreTrait::round at precision 3 (agent_base.rs:311-314 ->
//     math_utils.rs:10-13) maps a score to its milli-unit index, so ordering neighbours by
//     (duplicate-stop change, tour-length change) IS ordering them by their rounded scores -- with no
//     f64 instruction, no rounding key and half the L2 sectors per gathered edge.  The reference's own
//     f64 fold lands within one quantum of that index (the exact tour length sits ON a truncation
//     boundary; which side it falls is decided by its summation order), the declared float tolerance.
//   * A swap of two stops, a 2-opt reversal and an insertion need: ONE Philox4x32-10 block (move
//     kind, group draw, two ranks), two loads from the island's free-position list, six tour loads
//     and <= 4 removed-edge loads from shared memory, <= 4 added-edge gathers from D32 in L2.
//     Nothing is indexed dynamically: no GjMove descriptor, no local memory.  Every other move
//     (change / swap_edges / scramble) goes to a work list and through the generic generator +
//     delta evaluator of gj_moves.cuh / gj_delta.cuh in an out-of-line slow path.
//   * The island's n + 1 edge lengths persist in HBM (f64, the values the matrix holds) and arrive
//     with the tour and the tabu table as TMA bulk copies; an accepted move patches only the edges it
//     touched.  The stored score of an accepted neighbour is still the FULL evaluation in the
//     reference's summation order (exact sums), or the exact integer sum of the edges (tree mode).
//
// Eligibility (checked by the host, gj_islands.cu): TSP, TabuSearch, one semantic group of consecutive
// columns with uniform bounds, symmetric matrix with a valid D32, weight 1, score_precision[1] == 3,
// mutation_rate_multiplier None.  Everything else keeps k_ls_step_fused.
#pragma once

struct GjTsFastArgs {
    GjSelectArgs A;
    double* edge;               // [I][edge_stride] f64 edge lengths (n + 1 live) of every island's current tour
    int edge_stride;            // n + 1 rounded up to an even count: 16-byte rows for the TMA copies
    int* stale;                 // [I] 1: the tour was replaced (migrant / global top) -> rebuild the edges
    uint32_t kind_thr[5];       // move-kind thresholds on the raw 32-bit draw (see gj_kind_thresholds_u32)
    int first;                  // first column of the group (columns first .. first + glen - 1)
    int glen;
    int cnt_stride;             // 32 * bm_words: value counts of the slow path
    double* scores_out;         // trace only: [I][K][2]
    GjMove* moves_out;          // trace only
    int* worklist;              // [I][K]
    long long* phase_clocks;
    unsigned int* done_counter; // not null: the last island to finish publishes the global top (k_global_top's work)
    // not null: update_global_top's publish half by key tournament (see the end of the kernel):
    // pub[0] = best key offered so far, pub[1] = key of the published global top
    unsigned long long* pub;
    // edge lengths of the published global top (k_ts_publish), valid for version *gedge_ver: an adopting
    // island copies them with the row instead of re-gathering n + 1 matrix entries
    const double* gedge; const int* gedge_ver;
};

__host__ __device__ inline size_t gj_tsfast_smem_bytes(int n_vars, int tabu_words, int cnt_stride) {
    const size_t n_pad = ((size_t)n_vars + 3) & ~(size_t)3;
    size_t b = (n_pad + 8) * 4;                                  // tour with sentinels
    b += (((size_t)n_vars + 1 + 3) & ~(size_t)3) * 4;           // e32
    b += (((size_t)tabu_words + 3) & ~(size_t)3) * 4;           // tabu table
    b += (size_t)cnt_stride * 4;                                 // value counts (slow path)
    b = (b + 15) & ~(size_t)15;
    b += (((size_t)n_vars + 2) & ~(size_t)1) * 8;               // e64
    return b;
}

// neighbour ordering key: (-d_uniq) * 2^40 + delta_milli; lower is better
#define GJ_TSF_HSHIFT 40
__device__ __forceinline__ long long gj_tsf_key(int d_uniq, long long delta_milli) {
    return ((long long)(-d_uniq) << GJ_TSF_HSHIFT) + delta_milli;
}
__device__ __forceinline__ void gj_tsf_unkey(long long key, int& d_uniq, long long& delta_milli) {
    const long long h = (key + (1ll << (GJ_TSF_HSHIFT - 1))) >> GJ_TSF_HSHIFT;
    d_uniq = (int)(-h);
    delta_milli = key - (h << GJ_TSF_HSHIFT);
}

// milli-unit index of a stored score level (the integer ScoreTrait::round, math_utils.rs:10-13,
// is built from)
__device__ __forceinline__ long long gj_tsf_milli_index(double v) {
    // nearest, not floor: a stored score is either already rounded (fl + j / 1000, whose fraction
    // times 1000 may land a hair under j) or an unrounded initial fold within 1e-9 of a whole index
    return llrint(v * 1000.0);
}
// ... and the score value the reference's rounding produces for that index
__device__ __forceinline__ double gj_tsf_from_index(long long k) {
    long long q = k / 1000, r = k % 1000;
    if (r < 0) { r += 1000; q -= 1; }
    return (double)q + (double)r / 1000.0;
}

// an agent top as a 64-bit key whose unsigned order is "better score first, lower island on ties"
__device__ __forceinline__ unsigned long long gj_tsf_top_key(double hard, double soft, int island) {
    return ((unsigned long long)llrint(hard) << 48) | ((unsigned long long)gj_tsf_milli_index(soft) << 12) |
           (unsigned long long)island;
}

// update_global_top, publish half, after a fixed-point step: the launch's best key against the
// published one (strictly better score only, agent_base.rs:451) -> row, score, version.  One CTA.
__global__ void __launch_bounds__(256)
k_ts_publish(unsigned long long* pub, int stride, int n_vars, const int32_t* __restrict__ best,
             const double* __restrict__ best_score, int32_t* gbest, double* gbest_score, int* gver,
             const double* __restrict__ cur_score, const int* __restrict__ stale, const int* __restrict__ dirty,
             const double* __restrict__ edge, int edge_stride, double* gedge, int* gedge_ver) {
    const unsigned long long k = pub[0], pk = pub[1];
    if ((k >> 12) >= (pk >> 12)) return;                       // uniform over the CTA
    const int owner = (int)(k & 0xfffull);
    for (int i = threadIdx.x; i < n_vars; i += blockDim.x) gbest[i] = best[(size_t)owner * stride + i];
    // update_top_individual copies on `<=`, so after a step equal scores <=> the owner's top row IS its
    // current row -- whose edge lengths the step keeps in HBM: published with the row
    bool with_edges = gedge != nullptr && stale[owner] == 0 && dirty[owner] == 0;
    for (int l = 0; l < 2; ++l)
        with_edges = with_edges && best_score[(size_t)owner * GJ_MAX_LEVELS + l] == cur_score[(size_t)owner * GJ_MAX_LEVELS + l];
    if (with_edges)
        for (int i = threadIdx.x; i < edge_stride; i += blockDim.x) gedge[i] = edge[(size_t)owner * edge_stride + i];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int l = 0; l < GJ_MAX_LEVELS; ++l) gbest_score[l] = best_score[(size_t)owner * GJ_MAX_LEVELS + l];
        *gver += 1;
        if (gedge_ver) *gedge_ver = with_edges ? *gver : 0;
        pub[1] = k;
    }
}

// ---- out-of-line slow path: generic generator / evaluator (rare) ------------------------------------
static __device__ __noinline__ void gj_tsf_generate(const GjProblemDev& P, const GjGroups& G, const GjSelectArgs& A,
                                                    const uint32_t* table, int island, int c, GjMove* out) {
    *out = gj_generate_move(P, G, A.M, A.seed, (uint32_t)(A.island_base + island), A.step, (uint32_t)c, table,
                            A.tabu_word_off);
}

static __device__ __noinline__ long long gj_tsf_slow_key(const GjProblemDev& P, const GjGroups& G, const GjSelectArgs& A,
                                                         const uint32_t* table, const int32_t* t, const int32_t* cnt,
                                                         int island, int c, GjMove* trace_out) {
    GjMove m;
    gj_tsf_generate(P, G, A, table, island, c, &m);
    if (trace_out) *trace_out = m;
    GjTspBase B{t, P.n_vars, P.D, (size_t)P.n_locations, nullptr, true};
    int d_uniq = 0; double d_dist = 0.0;
    gj_tsp_move_delta(P, G, m, A.noop != 0, true, B, cnt, d_uniq, d_dist);
    return gj_tsf_key(d_uniq, llrint(d_dist * 1000.0));
}

// TRACE: the instantiation gj_islands_trace_step uses (writes every neighbour's move and score)
// MB: resident CTAs per SM the register allocation is sized for (launch bounds)
// The draws of neighbour c in the order gj_generate_move consumes them (one Philox4x32-10 block):
// x[3] move kind, x[2] semantic group (one group: unused), x[1] and x[0] the two position ranks among
// the F free (non-tabu) positions.  Returns the kind; a0 / a1 are only meaningful for swap (1),
// insertion (4) and inverse (5), the kinds the fixed-point path scores itself.
__device__ __forceinline__ int gj_tsf_draw(uint32_t key0, uint32_t key1, uint32_t step_lo, uint32_t step_hi, int c,
                                           const uint32_t (&kind_thr)[5], uint32_t F, const int32_t* free_list,
                                           int& a0, int& a1) {
    const uint32_t ctr[4] = {step_lo, (uint32_t)c, step_hi, 0u}, key[2] = {key0, key1};
    uint32_t x[4];
    gj_philox_block(ctr, key, x);
    int kind = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) kind += (x[3] > kind_thr[i]) ? 1 : 0;
    const int r0 = (int)__umulhi(x[1], F);
    int r1 = (int)__umulhi(x[0], F - 1u);
    r1 += (r1 >= r0) ? 1 : 0;
    a0 = free_list ? free_list[r0] : r0;
    a1 = free_list ? free_list[r1] : r1;
    return kind;
}
#define GJ_TSF_FAST_KINDS 0x32u        // bit k set: kind k is scored by the fixed-point path (1, 4, 5)

template <int NT, int MB, bool TRACE>
__global__ void __launch_bounds__(NT, MB)
k_ts_step_fast(const __grid_constant__ GjProblemDev P, const __grid_constant__ GjGroups G,
               const __grid_constant__ GjTsFastArgs F) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long sh_wkey[32];
    __shared__ int sh_widx[32];
    __shared__ int sh_sel0[NT], sh_sel1[NT], sh_selinfo[NT];   // ids the last chunk's moves selected
    __shared__ int sh_scan[NT];
    __shared__ long long sh_isum[32];
    __shared__ __align__(8) uint64_t sh_mbar;
    __shared__ long long sh_curkey;
    __shared__ long long sh_bestkey;
    __shared__ int sh_accept, sh_best, sh_nwork, sh_adopt, sh_gedge;
    __shared__ GjMove sh_mv;
    __shared__ GjScore sh_gs[32];
    __shared__ int sh_gi[32];
    __shared__ int sh_gpub, sh_last;

    const GjSelectArgs& A = F.A;
    const int island = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const int K = A.K, n = P.n_vars;
    const size_t n_pad = ((size_t)n + 3) & ~(size_t)3;
    // carve
    size_t o = 0;
    int32_t* t = (int32_t*)(smem_raw + o) + 4; o += (n_pad + 8) * 4;
    int32_t* e32 = (int32_t*)(smem_raw + o); o += (((size_t)n + 1 + 3) & ~(size_t)3) * 4;
    uint32_t* table = (uint32_t*)(smem_raw + o); o += (((size_t)A.tabu_words_per_island + 3) & ~(size_t)3) * 4;
    int32_t* cnt = (int32_t*)(smem_raw + o); o += (size_t)F.cnt_stride * 4;
    o = (o + 15) & ~(size_t)15;
    double* e64 = (double*)(smem_raw + o);

    int32_t* cur_row = A.cur + (size_t)island * A.stride;
    double* edge_g = F.edge + (size_t)island * (size_t)F.edge_stride;
    const size_t L = (size_t)P.n_locations;
    auto stamp = [&](int k) {
        if (F.phase_clocks && tid == 0) F.phase_clocks[(size_t)island * 8 + k] = clock64();
    };
    stamp(0);

    // ---- P0: stage ---------------------------------------------------------------------------------
    // Everything the step's serial prologue touches is cold after another island group / kernel used
    // the L2: the lanes of warp 1 pull those lines in while thread 0 starts the bulk copies, so the
    // adoption test below runs on L2 hits instead of a chain of HBM round trips.
    const uint32_t row_bytes = (uint32_t)(n_pad * 4);
    const uint32_t tabu_bytes = A.tabu_bits ? (uint32_t)(A.tabu_words_per_island * 4) : 0u;
    const uint32_t edge_bytes = (uint32_t)((size_t)F.edge_stride * 8);
    if (tid == 64) {
        // this CTA's slice of the milli-unit matrix (the whole grid covers it once): the gathers of P1
        // then hit L2 even when another kernel evicted the table since the last step
        const size_t total = (size_t)P.n_locations * (size_t)P.n_locations * 4;
        size_t per = (total + gridDim.x - 1) / gridDim.x;
        per = (per + 127) & ~(size_t)127;
        const size_t off = (size_t)blockIdx.x * per;
        if (off < total) {
            const size_t len = min(per, total - off) & ~(size_t)15;
            if (len) gj_l2_prefetch_bulk((const char*)P.D32 + off, (uint32_t)len);
        }
        if (A.tabu_bits) {
            const int T = A.tabu_ring_per_island;
            const uint32_t bytes = (uint32_t)(((size_t)T * 4) & ~(size_t)15);     // rounded down: stays inside
            if (bytes) gj_l2_prefetch_bulk(A.tabu_ring_old + (size_t)island * T - (((size_t)island * T) & 3), bytes);
        }
    }
    if (warp == 1) {
        const void* pf = nullptr;
        switch (lane) {
            case 0: pf = A.gver; break;
            case 1: pf = A.gseen + island; break;
            case 2: pf = A.gbest_score; break;
            case 3: pf = A.best_score + (size_t)island * GJ_MAX_LEVELS; break;
            case 4: pf = A.cur_score + (size_t)island * GJ_MAX_LEVELS; break;
            case 5: pf = F.stale + island; break;
            case 6: pf = A.dirty + island; break;
            case 7: pf = A.tabu_fill ? A.tabu_fill + (size_t)island * A.n_groups : nullptr; break;
            case 8: pf = A.tabu_size; break;
            case 9: pf = A.tabu_ring_off; break;
            case 10: pf = A.tabu_word_off; break;
            default: break;
        }
        if (pf) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
    }
    if (tid == 0) {
        gj_mbar_init(&sh_mbar, 1);
        sh_nwork = 0;
        // all three bulk copies leave at once; the island's own row is loaded speculatively -- an adoption
        // (rare after the first steps) reloads the row from the global top in a second phase
        gj_mbar_expect_tx(&sh_mbar, row_bytes + tabu_bytes + edge_bytes);
        gj_tma_load_1d(t, cur_row, row_bytes, &sh_mbar);
        if (tabu_bytes) gj_tma_load_1d(table, A.tabu_bits + (size_t)island * A.tabu_words_per_island, tabu_bytes, &sh_mbar);
        gj_tma_load_1d(e64, edge_g, edge_bytes, &sh_mbar);       // overwritten below when the tour was replaced
        const bool adopt = gj_adopt_decide(A, island);            // update_global_top, adopt half
        sh_adopt = adopt ? 1 : 0;
        // the issuing thread polls the mbarrier (a spinning try_wait burns issue slots: with all eight
        // warps on it the poll loop was ~45 % of the kernel's executed instructions); the rest sleep
        gj_mbar_wait(&sh_mbar, 0);
        sh_gedge = 0;
        if (adopt) {
            // the adopted row comes with its edge lengths when the publisher could provide them
            const bool ge = F.gedge != nullptr && *F.gedge_ver == *A.gver;
            sh_gedge = ge ? 1 : 0;
            gj_mbar_expect_tx(&sh_mbar, row_bytes + (ge ? edge_bytes : 0u));
            gj_tma_load_1d(t, A.gbest, row_bytes, &sh_mbar);
            if (ge) gj_tma_load_1d(e64, F.gedge, edge_bytes, &sh_mbar);
            gj_mbar_wait(&sh_mbar, 1);
        }
    }
    __syncthreads();
    stamp(6);
    const bool adopted = sh_adopt != 0;
    const int state_stale = F.stale[island];
    if (tid == 0) { t[-1] = 0; t[n] = 0; }               // depot before the first and after the last stop
    if (adopted)
        for (int i = tid; i < n; i += NT) cur_row[i] = t[i];
    __syncthreads();
    if (state_stale && adopted && sh_gedge) {
        // the tour was replaced by the global top, whose edge lengths arrived with it
        for (int i = tid; i <= n; i += NT) edge_g[i] = e64[i];
        if (tid == 0) F.stale[island] = 0;
        __syncthreads();
    } else if (state_stale) {
        // the tour was replaced since the last step: all n + 1 edge lengths from the matrix
        for (int i = tid; i <= n; i += NT) {
            const double d = __ldg(&P.D[(size_t)t[i - 1] * L + (size_t)t[i]]);
            e64[i] = d;
            edge_g[i] = d;
        }
        if (tid == 0) F.stale[island] = 0;
        __syncthreads();
    }
    for (int i = tid; i <= n; i += NT) e32[i] = (int32_t)__double2int_rn(e64[i] * 1000.0);
    if (tid == 0) sh_curkey = gj_tsf_milli_index(A.cur_score[(size_t)island * GJ_MAX_LEVELS + 1]);
    sh_selinfo[tid] = 0;
    __syncthreads();

    const uint32_t* table_ro = A.tabu_bits ? table : nullptr;
    const int W = (F.glen + 31) >> 5;
    const int32_t* free_list = table_ro ? (const int32_t*)(table + 2 * (W + 1)) : nullptr;
    const int n_free = table_ro ? ((const int32_t*)table)[2 * W + 1] : 0;
    // Mover::select_non_tabu_ids: a fully tabu group falls back to plain choice (gj_pick_positions)
    const bool use_tabu = free_list != nullptr && n_free >= 2;
    const uint32_t Fd = use_tabu ? (uint32_t)n_free : (uint32_t)F.glen;
    const uint32_t gisl = (uint32_t)(A.island_base + island);
    const uint32_t key0 = (uint32_t)A.seed ^ (gisl * 0x9E3779B1u), key1 = (uint32_t)(A.seed >> 32) + gisl;
    const uint32_t step_lo = (uint32_t)A.step, step_hi = (uint32_t)(A.step >> 32);
    const int32_t* __restrict__ D32 = P.D32;
    const int Li = P.n_locations;
    const double cur_h = A.cur_score[(size_t)island * GJ_MAX_LEVELS + 0];
    const int n_chunks = (K + NT - 1) / NT;
    int* worklist = F.worklist + (size_t)island * K;

    stamp(1);
    // ---- P1: generate + score ------------------------------------------------------------------------
    // the hot loop's neighbours never change the duplicate count: their key is the 32-bit length change
    int best32 = 0x7fffffff;
    int best_idx = -1;
#pragma unroll 2
    for (int c = tid; c < K; c += NT) {
        int a0, a1;
        const int kind = gj_tsf_draw(key0, key1, step_lo, step_hi, c, F.kind_thr, Fd, use_tabu ? free_list : nullptr, a0, a1);
        const bool last_chunk = c >= (n_chunks - 1) * NT;
        if (!((GJ_TSF_FAST_KINDS >> kind) & 1u)) {           // not swap / insertion / inverse
            worklist[atomicAdd(&sh_nwork, 1)] = c;
            if (last_chunk) sh_selinfo[tid] = 0xff;         // P4 regenerates it
            continue;
        }
        if (last_chunk) { sh_sel0[tid] = a0; sh_sel1[tid] = a1; sh_selinfo[tid] = 3; }   // two ids, group 0
        const int p = F.first + min(a0, a1), q = F.first + max(a0, a1);
        const int pm = t[p - 1], tp = t[p], pn = t[p + 1];
        const int qm = t[q - 1], tq = t[q], qp = t[q + 1];
        const bool adj = (q == p + 1);
        const bool inv = (kind == 5);
        const bool ins_l = (kind == 4) && (a0 < a1);        // t[p] travels to the end
        const bool ins_r = (kind == 4) && !ins_l;           // t[q] travels to the front
        const bool swp_far = (kind == 1) && !adj;
        const int x0 = ins_l ? pn : tq;
        const int y1 = ins_r ? qm : tp;
        const int a2b = swp_far ? pn : tp;
        // removed edges: the island's own edge cache; added edges: D32 in L2 (see gj_tsp_move_delta)
        int removed = e32[p] + e32[q + 1];
        int added = __ldg(&D32[pm * Li + x0]) + __ldg(&D32[y1 * Li + qp]);
        if (!inv) { removed += e32[ins_r ? q : p + 1]; added += __ldg(&D32[tq * Li + a2b]); }
        if (swp_far) { removed += e32[q]; added += __ldg(&D32[qm * Li + tp]); }
        const int k32 = added - removed;
        const long long k64 = (long long)k32;
        if (TRACE && F.moves_out) {
            GjMove m;
            m.kind = (uint8_t)kind; m.group = 0; m.k = 2; m.pad = 0;
#pragma unroll
            for (int i = 0; i < GJ_MOVE_MAXK; ++i) { m.a[i] = 0; m.v[i] = 0; }
            m.a[0] = a0; m.a[1] = a1;
            F.moves_out[(size_t)island * K + c] = m;
        }
        if (TRACE && F.scores_out) {
            double* so = F.scores_out + ((size_t)island * K + c) * 2;
            so[0] = cur_h; so[1] = gj_tsf_from_index(sh_curkey + k64);
        }
        if (k32 < best32) { best32 = k32; best_idx = c; }      // increasing c: the first minimum stays
    }
    long long best_key = best_idx >= 0 ? (long long)best32 : 0x7fffffffffffffffll;
    __syncthreads();

    // ---- P1b: the queued neighbours (generic generator + delta evaluator, out of line) -----------------
    const int n_work = sh_nwork;
    if (n_work > 0) {
        for (int i = tid; i < F.cnt_stride; i += NT) cnt[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += NT) atomicAdd(&cnt[t[i] - P.val_lo], 1);
        __syncthreads();
        for (int w = tid; w < n_work; w += NT) {
            const int c = worklist[w];
            const long long k64 = gj_tsf_slow_key(P, G, A, table_ro, t, cnt, island, c,
                                                  (TRACE && F.moves_out) ? F.moves_out + (size_t)island * K + c : nullptr);
            if (TRACE && F.scores_out) {
                int du; long long dm;
                gj_tsf_unkey(k64, du, dm);
                double* so = F.scores_out + ((size_t)island * K + c) * 2;
                so[0] = cur_h - (double)du; so[1] = gj_tsf_from_index(sh_curkey + dm);
            }
            if (k64 < best_key || (k64 == best_key && c < best_idx)) { best_key = k64; best_idx = c; }
        }
    }

    stamp(2);
    // ---- P2: first minimum (tabu_search_base.rs:166-171) + acceptance (:174) ---------------------------
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const long long ok = __shfl_xor_sync(GJ_FULL_MASK, best_key, off);
        const int oi = __shfl_xor_sync(GJ_FULL_MASK, best_idx, off);
        if (oi >= 0 && (best_idx < 0 || ok < best_key || (ok == best_key && oi < best_idx))) { best_key = ok; best_idx = oi; }
    }
    if (lane == 0) { sh_wkey[warp] = best_key; sh_widx[warp] = best_idx; }
    __syncthreads();
    if (tid == 0) {
        long long bk = sh_wkey[0]; int bi = sh_widx[0];
        for (int w = 1; w < nwarps; ++w) {
            const long long ok = sh_wkey[w]; const int oi = sh_widx[w];
            if (oi >= 0 && (bi < 0 || ok < bk || (ok == bk && oi < bi))) { bk = ok; bi = oi; }
        }
        // candidate <= current (tabu_search_base.rs:174): hard level first, then the milli-unit index
        const bool accept = bi >= 0 && bk <= 0;
        sh_accept = accept ? 1 : 0;
        sh_best = bi;
        sh_bestkey = bk;
        if (accept) {
            // the winning move again (a pure function of its index): inline for the common kinds
            int a0, a1;
            const int kind = gj_tsf_draw(key0, key1, step_lo, step_hi, bi, F.kind_thr, Fd, use_tabu ? free_list : nullptr, a0, a1);
            if ((GJ_TSF_FAST_KINDS >> kind) & 1u) {
                sh_mv.kind = (uint8_t)kind; sh_mv.group = 0; sh_mv.k = 2; sh_mv.pad = 0;
#pragma unroll
                for (int i = 0; i < GJ_MOVE_MAXK; ++i) { sh_mv.a[i] = 0; sh_mv.v[i] = 0; }
                sh_mv.a[0] = a0; sh_mv.a[1] = a1;
            } else {
                gj_tsf_generate(P, G, A, table_ro, island, bi, &sh_mv);
            }
        }
        if (A.selected_out) { A.selected_out[island] = bi; A.accepted_out[island] = accept ? 1 : 0; }
        atomicAdd(&A.counters[0], (unsigned long long)K);
        if (island == 0) atomicAdd(&A.counters[1], 1ull);
        if (accept) atomicAdd(&A.counters[2], 1ull);
    }
    __syncthreads();

    stamp(3);
    // ---- P3: apply, write back, patch the edges, exact re-score, update_top_individual ----------------
    if (sh_accept) {
        const GjMove m = sh_mv;
        const bool fast_kind = ((GJ_TSF_FAST_KINDS >> m.kind) & 1u) && m.k == 2 && m.kind != GJ_MOVE_NULL;
        if (fast_kind) {
            // swap / 2-opt / insertion on the staged tour itself (the group's columns are consecutive
            // and share their bounds: no clamp, no group table); only the touched range goes back to HBM
            const int c0 = F.first + m.a[0], c1 = F.first + m.a[1];
            const int p = min(c0, c1), q = max(c0, c1);
            const int len = q - p + 1;
            if (m.kind == 1) {
                if (tid == 0) { const int x = t[p]; t[p] = t[q]; t[q] = x; }
            } else if (m.kind == 5) {
                for (int j = tid; j < len / 2; j += NT) { const int x = t[p + j]; t[p + j] = t[q - j]; t[q - j] = x; }
            } else {
                int32_t* scratch = e32;                      // dead since the end of P1
                for (int i = tid; i < len; i += NT) scratch[i] = t[p + i];
                __syncthreads();
                for (int i = tid; i < len; i += NT) t[p + i] = scratch[gj_segment_src_slot(m, true, i, len)];
            }
            __syncthreads();
            if (m.kind == 1) { if (tid < 2) cur_row[tid ? q : p] = t[tid ? q : p]; }
            else for (int i = p + tid; i <= q; i += NT) cur_row[i] = t[i];
        } else {
            gj_apply_move(P, m, G, true, A.noop != 0, tid, NT,
                          [&](int id) { return cur_row[id]; }, [&](int id, int v) { t[id] = v; });
            __syncthreads();
            for (int i = tid; i < n; i += NT) cur_row[i] = t[i];
        }
        const bool identity = m.kind == GJ_MOVE_NULL || (A.noop && (m.kind == 3 || (m.kind == 2 && m.k == 2)));
        if (!identity) {
            const int c0 = F.first + m.a[0], c1 = F.first + m.a[1];
            const int p = min(c0, c1), q = max(c0, c1);
            auto regather = [&](int i) {
                if (i >= 0 && i <= n) {
                    const double d = __ldg(&P.D[(size_t)t[i - 1] * L + (size_t)t[i]]);
                    e64[i] = d; edge_g[i] = d;
                }
            };
            if (m.kind == 1 && m.k == 2) {
                if (tid == 0) regather(p);
                if (tid == 1) regather(p + 1);
                if (tid == 2) regather(q);
                if (tid == 3) regather(q + 1);
            } else if (m.kind == 5) {
                // symmetric matrix: the interior edges p+1 .. q keep their lengths in reverse order
                const int len = q - p;
                for (int j = tid; j < len / 2; j += NT) {
                    const double xa = e64[p + 1 + j], ya = e64[q - j];
                    e64[p + 1 + j] = ya; e64[q - j] = xa;
                    edge_g[p + 1 + j] = ya; edge_g[q - j] = xa;
                }
                if (tid == 0) regather(p);
                if (tid == 1) regather(q + 1);
            } else if (m.kind == 4) {
                for (int i = p + tid; i <= q + 1; i += NT) regather(i);
            } else {
                for (int i = tid; i <= n; i += NT) regather(i);       // generic small move
            }
        }
        __syncthreads();
        // the accepted neighbour's stored score = FULL evaluation of the stored vector: hard from the
        // exact duplicate count, soft as the reference's own fold (tsp ISC :76-80) or the integer sum
        int du; long long dm;
        gj_tsf_unkey(sh_bestkey, du, dm);
        const double hard = cur_h - (double)du;
        if (P.exact_sums) {
            if (tid == 0) {
                double fold = 0.0;
#pragma unroll 8
                for (int i = 1; i < n; ++i) fold = fold + e64[i];
                double sample_distance = 0.0;
                sample_distance += e64[0];                 // D[0][s0]
                sample_distance += e64[n];                 // D[s_last][0]
                sample_distance += fold;
                GjScore sc;
                sc.v[0] = hard; sc.v[1] = sample_distance; sc.v[2] = 0.0;
                // ISC: both levels carry weights[0] == 1 (checked by the host): 0.0 + 1.0 * x == x
                gj_score_round(sc, P);
                A.cur_score[(size_t)island * GJ_MAX_LEVELS + 0] = sc.v[0];
                A.cur_score[(size_t)island * GJ_MAX_LEVELS + 1] = sc.v[1];
                A.cur_score[(size_t)island * GJ_MAX_LEVELS + 2] = 0.0;
                A.dirty[island] = 1;
            }
        } else {
            long long acc = 0;
            for (int i = tid; i <= n; i += NT) acc += (long long)__double2int_rn(e64[i] * 1000.0);
            acc = gj_warp_sum(acc);
            if (lane == 0) sh_isum[warp] = acc;
            __syncthreads();
            if (tid == 0) {
                long long tot = 0;
                for (int w = 0; w < nwarps; ++w) tot += sh_isum[w];
                A.cur_score[(size_t)island * GJ_MAX_LEVELS + 0] = hard;
                A.cur_score[(size_t)island * GJ_MAX_LEVELS + 1] = gj_tsf_from_index(tot);
                A.cur_score[(size_t)island * GJ_MAX_LEVELS + 2] = 0.0;
                A.dirty[island] = 1;
            }
        }
        __syncthreads();
    }
    gj_update_top(island, A.levels, A.stride, n, A.cur, A.cur_score, A.best, A.best_score, A.dirty);

    stamp(4);
    // ---- P4: tabu deque update (mover.rs:75-96; see gj_tabu_deque_advance) -------------------------------
    if (A.tabu_bits) {
        // the new deque and the new table are built in shared memory (e32 and the staged table are dead
        // by now) and leave as coalesced copies: no global round trip between the phases of the rebuild
        uint32_t* bits_rw = A.tabu_bits + (size_t)island * A.tabu_words_per_island;
        const int32_t* ring_old = A.tabu_ring_old + (size_t)island * A.tabu_ring_per_island + A.tabu_ring_off[0];
        int32_t* ring_new = A.tabu_ring_new + (size_t)island * A.tabu_ring_per_island + A.tabu_ring_off[0];
        int32_t* ring_s = e32;                               // T <= glen <= n ints
        const int T = A.tabu_size[0];
        const int fill_old = A.tabu_fill[island * A.n_groups];
        int collected = 0;
        for (int chunk = n_chunks - 1; chunk >= 0 && collected < T; --chunk) {
            const int j = chunk * NT + tid;
            int sel[GJ_MOVE_MAXK]; int cntsel = 0;
            if (j < K) {
                const int info = sh_selinfo[tid];
                if (chunk == n_chunks - 1 && info != 0xff) {
                    if (info == 3) { cntsel = 2; sel[0] = sh_sel0[tid]; sel[1] = sh_sel1[tid]; }
                } else {
                    int a0, a1;
                    const int kind = gj_tsf_draw(key0, key1, step_lo, step_hi, j, F.kind_thr, Fd, use_tabu ? free_list : nullptr, a0, a1);
                    if ((GJ_TSF_FAST_KINDS >> kind) & 1u) {
                        cntsel = 2; sel[0] = a0; sel[1] = a1;
                    } else {
                        GjMove m;
                        gj_tsf_generate(P, G, A, table_ro, island, j, &m);
                        if (m.kind != GJ_MOVE_NULL) cntsel = gj_move_selected(m, sel);
                    }
                }
            }
            int total;
            const int incl = gj_block_scan_incl(cntsel, sh_scan, &total);
            const int after = total - incl;
#pragma unroll
            for (int i = 0; i < GJ_MOVE_MAXK; ++i) {
                if (i < cntsel) {
                    const int rank = collected + after + (cntsel - 1 - i);
                    if (rank < T) ring_s[rank] = sel[i];
                }
            }
            __syncthreads();
            collected += total;
        }
        for (int r = collected + tid; r < T; r += NT) {
            const int rho = r - collected;
            if (rho < fill_old) ring_s[r] = ring_old[rho];
        }
        const int fill = min(T, fill_old + collected);
        __syncthreads();
        if (tid == 0) A.tabu_fill[island * A.n_groups] = fill;
        for (int r = tid; r < fill; r += NT) ring_new[r] = ring_s[r];
        gj_tabu_table_rebuild(table + A.tabu_word_off[0], F.glen, ring_s, fill, sh_scan);
        for (int w = tid; w < A.tabu_words_per_island; w += NT) bits_rw[w] = table[w];
    }
    stamp(5);
    // ---- update_global_top, publish half (agent_base.rs:451-461) ------------------------------------------
    // The global top is the best agent top, lowest island on ties, replaced only by a strictly better
    // one (k_global_top).  In fixed point an agent top packs into 64 bits -- hard level | milli-unit
    // index | island -- whose unsigned order IS that rule, so the reduction over all islands is one
    // atomicMin per island here; k_ts_publish (one small CTA after the step) compares the winner with
    // the published top and copies its row.  (The row cannot be copied from inside this kernel: islands
    // of a later wave may be reading the published row for their adoption at that very moment.)
    if (F.pub) {
        if (tid == 0) {
            const double th = A.best_score[(size_t)island * GJ_MAX_LEVELS + 0];
            const double ts = A.best_score[(size_t)island * GJ_MAX_LEVELS + 1];
            atomicMin(&F.pub[0], gj_tsf_top_key(th, ts, island));
        }
    } else if (F.done_counter) {
        // fallback (more than 4096 islands): the LAST island to finish runs k_global_top's reduction
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned prev = atomicAdd(F.done_counter, 1u);
            sh_last = (prev == gridDim.x - 1) ? 1 : 0;
            if (sh_last) *F.done_counter = 0u;
        }
        __syncthreads();
        if (sh_last) {
            __threadfence();
            gj_global_top_cta((int)gridDim.x, A.levels, A.stride, n, A.best, A.best_score, A.gbest, A.gbest_score, A.gver,
                              sh_gs, sh_gi, &sh_gpub);
        }
    }
}
