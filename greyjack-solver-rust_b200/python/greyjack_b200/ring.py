"""Cross-GPU island ring: the reference's agent ring `i -> (i + 1) mod n` (solver/solver.rs:85-92,
AgentToAgentUpdate over crossbeam channels, agent_base.rs:322-444) with one process per GPU.

Each rank owns a group of islands (a gj_islands handle).  Inside a group the ring is handled on
the device by gj_islands_step; this module closes the ring ACROSS ranks: every
`migration_frequency` steps the migrants of rank r's last island travel to rank r+1's first
island (torch.distributed send/recv: NCCL over NVLink on GPUs, gloo in the CPU tests), where the
reference's acceptance rule (agent_base.rs:414-440) is applied by gj_islands_import_migrants.
There is no other data-path collective: islands are independent between migrations.

The `islands` object only needs migrant_bytes / export_migrants / import_migrants /
set_external_ring (greyjack_b200.Islands, or a stand-in in tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def ring_neighbours(rank: int, world: int):
    """(destination, source) of `rank` in the ring i -> (i + 1) mod world."""
    return (rank + 1) % world, (rank - 1) % world


def global_island_base(rank: int, islands_per_rank: int) -> int:
    """Global id of a rank's island 0 (RNG key and ring position, solver.rs:92)."""
    return rank * islands_per_rank


class RingMigrator:
    """Ring (and, for local-search agents, the shared global top) with the transport in
    torch.distributed: NCCL on GPUs, gloo in the CPU tests."""

    def __init__(self, islands, rank: int, world: int, islands_per_rank: int, device="cuda", group=None,
                 share_global_top: bool = False):
        self.islands, self.rank, self.world, self.group = islands, rank, world, group
        self.dst, self.src = ring_neighbours(rank, world)
        islands.set_external_ring(world > 1, global_island_base(rank, islands_per_rank))
        n = int(islands.migrant_bytes())
        self.out = torch.empty(n, dtype=torch.uint8, device=device)
        self.inp = torch.empty(n, dtype=torch.uint8, device=device)
        self.share_global_top = share_global_top and world > 1
        if self.share_global_top:
            gb = int(islands.global_top_bytes())
            self.g_out = torch.zeros(gb, dtype=torch.uint8, device=device)
            self.g_all = torch.zeros(gb * world, dtype=torch.uint8, device=device)
        self.exchanges = 0

    def exchange(self, stream: int = 0):
        """One ring exchange; call it every migration_frequency steps on every rank."""
        if self.world == 1:
            return
        # `stream` must be torch's current stream: NCCL orders its send after the work queued on
        # it (the export kernels), and req.wait() orders the import after the receive
        self.islands.export_migrants(self.out.data_ptr(), stream)
        ops = [dist.P2POp(dist.isend, self.out, self.dst, group=self.group),
               dist.P2POp(dist.irecv, self.inp, self.src, group=self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self.islands.import_migrants(self.inp.data_ptr(), stream)
        if self.share_global_top:
            # update_global_top across ranks (agent_base.rs:446-490): every rank's record to every rank,
            # the best one adopted on the device when strictly better -- no host decision
            self.islands.export_global_top(self.g_out.data_ptr(), stream)
            dist.all_gather_into_tensor(self.g_all, self.g_out, group=self.group)
            self.islands.import_global_top(self.g_all.data_ptr(), self.world, stream)
        self.exchanges += 1


class PeerRing:
    """Ring + shared global top over CUDA peer memory (csrc/gj_ring.cu): migrants and global-top
    records are stored straight into the neighbours' inboxes over NVLink, nothing waits on the host
    and no collective runs on the data path.  torch.distributed is only used ONCE, to hand the IPC
    handles of the inboxes around."""

    def __init__(self, islands, rank: int, world: int, islands_per_rank: int, group=None):
        import ctypes as C
        from . import _lib
        self.islands, self.rank, self.world = islands, rank, world
        self._L = _lib.load()
        islands.set_external_ring(world > 1, global_island_base(rank, islands_per_rank))
        self.handle = C.c_void_p()
        _lib.check(self._L.gj_ring_create(islands.handle, C.c_int32(rank), C.c_int32(world), C.byref(self.handle)))
        mine = (C.c_ubyte * 64)()
        _lib.check(self._L.gj_ring_handle(self.handle, mine))
        table = [None] * world
        if world > 1:
            dist.all_gather_object(table, bytes(mine), group=group)
            flat = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(table))
            _lib.check(self._L.gj_ring_connect(self.handle, flat))
            dist.barrier(group=group)
        self.exchanges = 0

    def exchange(self, stream: int = 0):
        from . import _lib
        import ctypes as C
        if self.world == 1:
            return
        _lib.check(self._L.gj_ring_exchange(self.handle, C.c_void_p(stream)))
        self.exchanges += 1

    def stats(self):
        from . import _lib
        import ctypes as C
        e, m = C.c_int64(0), C.c_int64(0)
        _lib.check(self._L.gj_ring_stats(self.handle, C.byref(e), C.byref(m)))
        return {"exchanges": e.value, "missed": m.value}

    def close(self):
        if getattr(self, "handle", None):
            self._L.gj_ring_destroy(self.handle)
            self.handle = None


class HybridRing:
    """ONE ring over island groups of different agent kinds (BASELINE config 5: GeneticAlgorithm +
    TabuSearch hybrid islands).  On every rank the groups are chained g0 -> g1 -> ... -> g_last and
    g_last feeds g0 of the next rank: the reference ring i -> (i + 1) mod n (solver.rs:85-92) over ALL
    agents.  A link carries the sender group's migrants (its last island's individual; a GA island's
    best `migrants` individuals) into the receiver group, which applies ITS OWN acceptance rule
    (receive_updates, agent_base.rs:405-440: LocalSearch agents compare with population[0], Population
    agents with their worst individuals).  Groups must agree on the bytes per exchange (give the GA a
    migration_rate with ceil(rate * pop) == 1 next to local-search groups).  The cross-rank link uses
    torch.distributed point-to-point (NCCL on GPUs)."""

    def __init__(self, groups, rank: int, world: int, device="cuda", group=None):
        self.groups, self.rank, self.world, self.group = list(groups), rank, world, group
        per_rank = sum(g.n_islands for g in self.groups)
        base = rank * per_rank
        for g in self.groups:
            g.set_external_ring(True, base)          # the ring is closed here, not inside gj_islands_step
            base += g.n_islands
        sizes = {int(g.migrant_bytes()) for g in self.groups}
        if len(sizes) != 1:
            raise ValueError(f"groups exchange different migrant sizes {sorted(sizes)}: set the GeneticAlgorithm's "
                             "migration_rate so that ceil(rate * population) == 1")
        n = sizes.pop()
        self.out = [torch.empty(n, dtype=torch.uint8, device=device) for _ in self.groups]
        self.inp = torch.empty(n, dtype=torch.uint8, device=device)
        self.dst, self.src = ring_neighbours(rank, world)
        self.exchanges = 0

    def exchange(self, stream: int = 0):
        # every agent sends the individual it held BEFORE receiving: export everything first
        for g, buf in zip(self.groups, self.out):
            g.export_migrants(buf.data_ptr(), stream)
        for i in range(1, len(self.groups)):
            self.groups[i].import_migrants(self.out[i - 1].data_ptr(), stream)
        if self.world == 1:
            self.groups[0].import_migrants(self.out[-1].data_ptr(), stream)
        else:
            ops = [dist.P2POp(dist.isend, self.out[-1], self.dst, group=self.group),
                   dist.P2POp(dist.irecv, self.inp, self.src, group=self.group)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            self.groups[0].import_migrants(self.inp.data_ptr(), stream)
        self.exchanges += 1

    def describe(self) -> str:
        kinds = " -> ".join(f"{type(g.builder).__name__} x{g.n_islands}" for g in self.groups)
        return f"per GPU: {kinds} -> next GPU ({self.world} GPU{'s' if self.world > 1 else ''})"

    def close(self):
        pass


def run_steps(islands, migrator: RingMigrator, n_steps: int, migration_frequency: int, stream: int = 0,
              first_step: int = 0):
    """Agent::solve's loop across ranks: step, and every migration_frequency steps exchange
    (agent_base.rs:161-183)."""
    for s in range(first_step, first_step + n_steps):
        islands.step(1, stream)
        if (s + 1) % migration_frequency == 0:
            migrator.exchange(stream)


def global_best(score, world: int, device="cpu", group=None):
    """Lexicographic minimum (lower is better) of every rank's best score -> (score, owner rank).
    Replaces the Arc<Mutex<global_top_individual>> across processes (agent_base.rs:446-490)."""
    t = torch.tensor([float(x) for x in score], dtype=torch.float64, device=device)
    if world == 1:
        return t.tolist(), 0
    allv = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allv, t, group=group)
    rows = [tuple(v.tolist()) for v in allv]
    owner = min(range(world), key=lambda r: (rows[r], r))
    return list(rows[owner]), owner
