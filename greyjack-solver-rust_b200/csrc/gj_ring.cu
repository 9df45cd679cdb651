// gj_ring.cu -- the island ring and the shared global top ACROSS GPUs over CUDA peer memory
// (NVLink / NVSwitch), one process per GPU.
//
// Reference: agents form a ring i -> (i + 1) mod n over crossbeam channels (solver/solver.rs:85-92,
// send_updates / receive_updates agent_base.rs:322-444) and share ONE global_top_individual behind a
// mutex (update_global_top, agent_base.rs:446-490).  Inside a GPU both live in gj_islands.cu; this
// file closes them across ranks without a collective library on the data path:
//
//   * every rank owns an INBOX in its own HBM (cudaMalloc + cudaIpcGetMemHandle): two migrant slots
//     (exchange parity) and a board with one global-top record per rank and parity;
//   * an exchange = (1) the rank's outgoing migrants are packed straight into the NEXT rank's inbox
//     and its global top into EVERY rank's board -- plain stores / copies on peer-mapped pointers that
//     travel over NVLink -- followed by a system-scope fence and a sequence flag; (2) a kernel on the
//     receiving rank waits (bounded) for the flags of this exchange, applies the reference's
//     acceptance rule to the migrants (gj_islands_import_migrants) and adopts the best record of the
//     board when it is strictly better than the rank's own global top.
//   Nothing blocks the host; the only cross-GPU traffic is ~4 KB per neighbour and per board entry.
//
// The senders of an exchange never wait (every rank issues its send kernels before its receive
// kernel, in stream order), so the bounded wait on the receiving side cannot deadlock; a flag that
// does not arrive within the time-out (a rank that fell behind by more than that) is skipped and
// counted -- the exchange is then simply missed, like a channel message that arrives after the
// agent has moved on.
#include <cstring>
#include <memory>

#include "gj_islands_dev.cuh"

static constexpr int kRingMaxWorld = 64;
static constexpr unsigned long long kRingTimeoutNs = 200ull * 1000ull * 1000ull;     // 200 ms

struct gj_ring {
    gj_islands* g = nullptr;
    int rank = 0, world = 1;
    size_t migrant_bytes = 0, record_bytes = 0, slot_bytes = 0, board_off = 0, inbox_bytes = 0;
    unsigned char* inbox = nullptr;                 // own (cudaMalloc)
    unsigned char* peer[kRingMaxWorld] = {};        // peer[r] = rank r's inbox mapped here (peer[rank] = inbox)
    bool opened[kRingMaxWorld] = {};
    unsigned long long seq = 0;                     // exchanges done
    unsigned int* missed = nullptr;                 // device counter: flags that timed out
    unsigned char* staging = nullptr;               // own global-top record, packed before the peer copies
};

// inbox layout: [2][flag 16 B | migrants]  then  [world][2][flag 16 B | row stride*4 | score 24 B]
static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// The whole send side of an exchange as ONE launch: CTA `world` stores the outgoing migrants into the next
// rank's inbox, CTA r (r != rank) stores this rank's global-top record into rank r's board -- plain
// 16-byte stores on peer-mapped pointers -- each followed by a system-scope fence and the sequence flag.
struct GjRingPeers { unsigned char* inbox[kRingMaxWorld]; };

__global__ void __launch_bounds__(256)
k_ring_send(GjRingPeers peers, int rank, int world, int parity, unsigned long long value, size_t slot_bytes,
            size_t board_off, size_t record_bytes, const unsigned char* __restrict__ migrants, size_t migrant_bytes,
            const int32_t* __restrict__ gbest, const double* __restrict__ gbest_score, int stride, int send_gtop) {
    const int r = blockIdx.x;
    unsigned char* dst;
    if (r == world) {
        dst = peers.inbox[(rank + 1) % world] + (size_t)parity * slot_bytes;
        if (((reinterpret_cast<uintptr_t>(migrants) | migrant_bytes) & 15) == 0) {
            const uint4* src = reinterpret_cast<const uint4*>(migrants);
            uint4* out = reinterpret_cast<uint4*>(dst + 16);
            for (size_t i = threadIdx.x; i < migrant_bytes / 16; i += blockDim.x) out[i] = src[i];
        } else {                                                             // rows are int32, scores f64: 4-byte units
            const uint32_t* src = reinterpret_cast<const uint32_t*>(migrants);
            uint32_t* out = reinterpret_cast<uint32_t*>(dst + 16);
            for (size_t i = threadIdx.x; i < migrant_bytes / 4; i += blockDim.x) out[i] = src[i];
        }
    } else {
        if (r == rank || !send_gtop) return;
        dst = peers.inbox[r] + board_off + ((size_t)rank * 2 + parity) * record_bytes;
        int32_t* row = reinterpret_cast<int32_t*>(dst + 16);
        for (int i = threadIdx.x; i < stride; i += blockDim.x) row[i] = gbest[i];
        if (threadIdx.x < GJ_MAX_LEVELS) reinterpret_cast<double*>(dst + 16 + (size_t)stride * 4)[threadIdx.x] = gbest_score[threadIdx.x];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *(volatile unsigned long long*)dst = value;
        __threadfence_system();
    }
}

__device__ __forceinline__ bool gj_ring_wait_flag(const unsigned long long* flag, unsigned long long value) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        const unsigned long long v = *(volatile const unsigned long long*)flag;
        if (v >= value) { __threadfence_system(); return true; }
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > kRingTimeoutNs) return false;
        __nanosleep(200);
    }
}

// waits for the migrant flag of this exchange; *ok = 1 when the payload may be imported
__global__ void k_ring_wait_migrants(const unsigned long long* flag, unsigned long long value, int* ok,
                                     unsigned int* missed) {
    const bool got = gj_ring_wait_flag(flag, value);
    *ok = got ? 1 : 0;
    if (!got) atomicAdd(missed, 1u);
}

// packs this group's global top (row + score) into `out`
__global__ void k_ring_pack_gtop(const int32_t* __restrict__ gbest, const double* __restrict__ gbest_score,
                                 int stride, unsigned char* out) {
    int32_t* row = (int32_t*)out;
    for (int i = threadIdx.x; i < stride; i += blockDim.x) row[i] = gbest[i];
    if (threadIdx.x < GJ_MAX_LEVELS) ((double*)(out + (size_t)stride * 4))[threadIdx.x] = gbest_score[threadIdx.x];
}

// update_global_top across ranks: the best record of the board (lowest rank on ties) replaces this
// group's global top when it is STRICTLY better (agent_base.rs:451); islands adopt it through the
// usual version bump.  One CTA.
__global__ void __launch_bounds__(256)
k_ring_merge_gtop(const unsigned char* board, int world, int me, int parity, unsigned long long value, size_t record_bytes,
                  int stride, int n_vars, int levels, int32_t* gbest, double* gbest_score, int* gver,
                  unsigned long long* ts_pub, unsigned long long step, unsigned int* missed) {
    __shared__ int sh_win;
    if (threadIdx.x == 0) {
        GjScore best = gj_load_score(gbest_score, levels);
        int win = -1;
        for (int r = 0; r < world; ++r) {
            if (r == me) continue;
            const unsigned char* rec = board + ((size_t)r * 2 + parity) * record_bytes;
            if (!gj_ring_wait_flag((const unsigned long long*)rec, value)) { atomicAdd(missed, 1u); continue; }
            const double* sc = (const double*)(rec + 16 + (size_t)stride * 4);
            GjScore s;
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) s.v[l] = (l < levels) ? sc[l] : 0.0;
            if (!gj_score_le(best, s, levels)) { best = s; win = r; }      // strictly better; first rank on ties
        }
        sh_win = win;
        if (win >= 0) {
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) gbest_score[l] = best.v[l];
            *gver += 1;
            if (ts_pub) {
                // keep the fixed-point step's published key in step (gj_islands_tsfast.cuh): an island id
                // beyond every local one marks "owned by another rank"
                ts_pub[1] = ((unsigned long long)llrint(best.v[0]) << 48) |
                            ((unsigned long long)llrint(best.v[1] * 1000.0) << 12) | 0xfffull;
            }
        }
    }
    __syncthreads();
    if (sh_win >= 0) {
        const int32_t* row = (const int32_t*)(board + ((size_t)sh_win * 2 + parity) * record_bytes + 16);
        for (int i = threadIdx.x; i < n_vars; i += blockDim.x) gbest[i] = row[i];
    }
}

extern "C" gj_status gj_ring_create(gj_islands* g, int32_t rank, int32_t world, gj_ring** out) {
    if (!g || !out || world < 1 || world > kRingMaxWorld || rank < 0 || rank >= world)
        return gj_fail(GJ_ERR_INVALID, "bad ring arguments");
    *out = nullptr;
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    std::unique_ptr<gj_ring> r(new gj_ring());
    r->g = g; r->rank = rank; r->world = world;
    r->migrant_bytes = (size_t)gj_islands_migrant_bytes(g);
    r->slot_bytes = align16(16 + r->migrant_bytes);
    r->record_bytes = align16(16 + (size_t)g->stride * 4 + GJ_MAX_LEVELS * 8);
    r->board_off = 2 * r->slot_bytes;
    r->inbox_bytes = r->board_off + (size_t)world * 2 * r->record_bytes;
    GJ_CUDA_TRY(cudaMalloc((void**)&r->inbox, r->inbox_bytes));
    GJ_CUDA_TRY(cudaMemset(r->inbox, 0, r->inbox_bytes));
    GJ_CUDA_TRY(cudaMalloc((void**)&r->missed, sizeof(unsigned int) + sizeof(int)));
    GJ_CUDA_TRY(cudaMemset(r->missed, 0, sizeof(unsigned int) + sizeof(int)));
    GJ_CUDA_TRY(cudaMalloc((void**)&r->staging, r->record_bytes));
    GJ_CUDA_TRY(cudaMemset(r->staging, 0, r->record_bytes));
    r->peer[rank] = r->inbox;
    GJ_CUDA_TRY(cudaDeviceSynchronize());
    *out = r.release();
    return GJ_OK;
}

extern "C" void gj_ring_destroy(gj_ring* r) {
    if (!r) return;
    cudaSetDevice(r->g ? r->g->device : 0);
    cudaDeviceSynchronize();
    for (int i = 0; i < r->world; ++i)
        if (r->opened[i] && r->peer[i]) cudaIpcCloseMemHandle(r->peer[i]);
    if (r->inbox) cudaFree(r->inbox);
    if (r->missed) cudaFree(r->missed);
    if (r->staging) cudaFree(r->staging);
    delete r;
}

extern "C" gj_status gj_ring_handle(gj_ring* r, gj_peer_handle* out) {
    if (!r || !out) return gj_fail(GJ_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(gj_peer_handle), "gj_peer_handle too small");
    GJ_CUDA_TRY(cudaSetDevice(r->g->p->device));
    cudaIpcMemHandle_t h;
    GJ_CUDA_TRY(cudaIpcGetMemHandle(&h, r->inbox));
    std::memset(out, 0, sizeof(*out));
    std::memcpy(out, &h, sizeof(h));
    return GJ_OK;
}

extern "C" gj_status gj_ring_connect(gj_ring* r, const gj_peer_handle* handles) {
    if (!r || !handles) return gj_fail(GJ_ERR_INVALID, "null argument");
    GJ_CUDA_TRY(cudaSetDevice(r->g->p->device));
    for (int i = 0; i < r->world; ++i) {
        if (i == r->rank || r->opened[i]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, &handles[i], sizeof(h));
        void* p = nullptr;
        GJ_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        r->peer[i] = (unsigned char*)p;
        r->opened[i] = true;
    }
    return GJ_OK;
}

extern "C" gj_status gj_ring_exchange(gj_ring* r, void* stream) {
    if (!r) return gj_fail(GJ_ERR_INVALID, "null ring");
    gj_islands* g = r->g;
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    if (r->world == 1) return GJ_OK;
    for (int i = 0; i < r->world; ++i)
        if (!r->peer[i]) return gj_fail(GJ_ERR_INVALID, "gj_ring_connect has not been called with every rank's handle");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long value = r->seq + 1;
    const int parity = (int)(r->seq & 1ull);
    gj_status rc;
    // ---- send: migrants -> next rank's inbox; global top -> every rank's board (one launch) ------------------
    const unsigned char* out_migrants = nullptr;
    if ((rc = gj_islands_pack_outgoing(g, st, &out_migrants))) return rc;
    const bool local_search = g->prm.agent != GJ_AGENT_GENETIC_ALGORITHM;
    // the record sent must be current (the fixed-point step publishes inside its own launches)
    if (local_search && !g->ts_fast && (rc = gj_ls_global_top(g, st))) return rc;
    GjRingPeers peers{};
    for (int i = 0; i < r->world; ++i) peers.inbox[i] = r->peer[i];
    k_ring_send<<<r->world + 1, 256, 0, st>>>(peers, r->rank, r->world, parity, value, r->slot_bytes, r->board_off,
                                              r->record_bytes, out_migrants, r->migrant_bytes, g->gbest, g->gbest_score,
                                              g->stride, local_search ? 1 : 0);
    GJ_LAUNCH_CHECK();
    // ---- receive ------------------------------------------------------------------------------------------
    unsigned char* in_slot = r->inbox + (size_t)parity * r->slot_bytes;
    int* ok = (int*)(r->missed + 1);
    k_ring_wait_migrants<<<1, 1, 0, st>>>((const unsigned long long*)in_slot, value, ok, r->missed);
    GJ_LAUNCH_CHECK();
    // a missed flag leaves the slot as the previous exchange of this parity wrote it: importing it again
    // is harmless (the acceptance rule decides), so the import is unconditional
    if ((rc = gj_islands_import_migrants(g, in_slot + 16, st))) return rc;
    if (local_search) {
        k_ring_merge_gtop<<<1, 256, 0, st>>>(r->inbox + r->board_off, r->world, r->rank, parity, value, r->record_bytes,
                                             g->stride, g->n_vars, g->levels, g->gbest, g->gbest_score, g->gver,
                                             g->ts_fast ? g->ts_pub : nullptr, g->step, r->missed);
        GJ_LAUNCH_CHECK();
        // (VRP chains re-index the global top lazily, at the start of their next launch: gj_launch_vrp_chains)
    }
    r->seq += 1;
    return GJ_OK;
}

extern "C" gj_status gj_ring_stats(gj_ring* r, int64_t* exchanges, int64_t* missed) {
    if (!r) return gj_fail(GJ_ERR_INVALID, "null ring");
    GJ_CUDA_TRY(cudaSetDevice(r->g->p->device));
    unsigned int m = 0;
    GJ_CUDA_TRY(cudaMemcpy(&m, r->missed, sizeof(m), cudaMemcpyDeviceToHost));
    if (exchanges) *exchanges = (int64_t)r->seq;
    if (missed) *missed = (int64_t)m;
    return GJ_OK;
}

// ---- the same two steps with the transport left to the caller (NCCL / gloo): device buffers ------------
extern "C" int64_t gj_islands_global_top_bytes(const gj_islands* g) {
    if (!g) return 0;
    return (int64_t)align16((size_t)g->stride * 4 + GJ_MAX_LEVELS * 8);
}

extern "C" gj_status gj_islands_export_global_top(gj_islands* g, void* d_buffer, void* stream) {
    if (!g || !d_buffer) return gj_fail(GJ_ERR_INVALID, "bad argument");
    if (g->prm.agent == GJ_AGENT_GENETIC_ALGORITHM) return gj_fail(GJ_ERR_UNSUPPORTED, "local-search agents only");
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    cudaStream_t st = (cudaStream_t)stream;
    gj_status rc;
    if ((rc = gj_ls_global_top(g, st))) return rc;
    k_ring_pack_gtop<<<1, 256, 0, st>>>(g->gbest, g->gbest_score, g->stride, (unsigned char*)d_buffer);
    GJ_LAUNCH_CHECK();
    return GJ_OK;
}

__global__ void __launch_bounds__(256)
k_import_gtop(const unsigned char* recs, int count, size_t record_bytes, int stride, int n_vars, int levels,
              int32_t* gbest, double* gbest_score, int* gver, unsigned long long* ts_pub, unsigned long long step) {
    __shared__ int sh_win;
    if (threadIdx.x == 0) {
        GjScore best = gj_load_score(gbest_score, levels);
        int win = -1;
        for (int r = 0; r < count; ++r) {
            const double* sc = (const double*)(recs + (size_t)r * record_bytes + (size_t)stride * 4);
            GjScore s;
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) s.v[l] = (l < levels) ? sc[l] : 0.0;
            if (!gj_score_le(best, s, levels)) { best = s; win = r; }
        }
        sh_win = win;
        if (win >= 0) {
            for (int l = 0; l < GJ_MAX_LEVELS; ++l) gbest_score[l] = best.v[l];
            *gver += 1;
            if (ts_pub) {
                ts_pub[1] = ((unsigned long long)llrint(best.v[0]) << 48) |
                            ((unsigned long long)llrint(best.v[1] * 1000.0) << 12) | 0xfffull;
            }
        }
    }
    __syncthreads();
    if (sh_win >= 0) {
        const int32_t* row = (const int32_t*)(recs + (size_t)sh_win * record_bytes);
        for (int i = threadIdx.x; i < n_vars; i += blockDim.x) gbest[i] = row[i];
    }
}

extern "C" gj_status gj_islands_import_global_top(gj_islands* g, const void* d_records, int32_t count, void* stream) {
    if (!g || !d_records || count < 0) return gj_fail(GJ_ERR_INVALID, "bad argument");
    if (g->prm.agent == GJ_AGENT_GENETIC_ALGORITHM) return gj_fail(GJ_ERR_UNSUPPORTED, "local-search agents only");
    GJ_CUDA_TRY(cudaSetDevice(g->p->device));
    cudaStream_t st = (cudaStream_t)stream;
    k_import_gtop<<<1, 256, 0, st>>>((const unsigned char*)d_records, count, (size_t)gj_islands_global_top_bytes(g), g->stride,
                                     g->n_vars, g->levels, g->gbest, g->gbest_score, g->gver,
                                     g->ts_fast ? g->ts_pub : nullptr, g->step);
    GJ_LAUNCH_CHECK();
    // (VRP chains re-index the global top lazily, at the start of their next launch: gj_launch_vrp_chains)
    return GJ_OK;
}
