"""World-size-2 (and 3) gloo tests of the cross-GPU island ring (greyjack_b200/ring.py) on CPU.
The islands are stand-ins backed by the ORACLE's acceptance rule (the real gj_islands needs a
GPU); what is under test is the host logic: ring topology, export -> send/recv -> import order,
the exchange cadence of run_steps and the cross-rank global best."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from greyjack_b200 import ring


class FakeIslands:
    """One island per rank; a migrant = [n_vars int32 | levels f64] like the device layout."""

    def __init__(self, rank, n_vars=8, levels=2):
        from oracle import gj_oracle
        self.o = gj_oracle
        self.n_vars, self.levels = n_vars, levels
        self.vec = np.full(n_vars, rank, dtype=np.int32)
        self.score = np.array([0.0, 100.0 - 10.0 * rank])      # higher rank = better (lower) score
        self.external = None
        self.steps = 0
        self.log = []

    def set_external_ring(self, on, island_base):
        self.external = (bool(on), island_base)

    def migrant_bytes(self):
        return self.n_vars * 4 + self.levels * 8

    def step(self, n, stream=0):
        self.steps += n
        self.log.append("step")

    def export_migrants(self, ptr, stream=0):
        buf = self.vec.tobytes() + self.score.tobytes()
        ctypes.memmove(ptr, buf, len(buf))
        self.log.append("export")

    # the group's global top as a device record (row int32 | score f64), gj_islands_*_global_top
    def global_top_bytes(self):
        return (self.n_vars * 4 + self.levels * 8 + 15) // 16 * 16

    def export_global_top(self, ptr, stream=0):
        gv, gs = getattr(self, "gtop", (self.vec, self.score))
        buf = gv.tobytes() + gs.tobytes()
        ctypes.memmove(ptr, buf, len(buf))
        self.log.append("export_gtop")

    def import_global_top(self, ptr, count, stream=0):
        gv, gs = getattr(self, "gtop", (self.vec, self.score))
        for r in range(count):
            raw = ctypes.string_at(ptr + r * self.global_top_bytes(), self.n_vars * 4 + self.levels * 8)
            vec = np.frombuffer(raw[: self.n_vars * 4], dtype=np.int32)
            score = np.frombuffer(raw[self.n_vars * 4:], dtype=np.float64)
            if self.o.score_cmp(score, gs) < 0:          # strictly better (agent_base.rs:451)
                gv, gs = vec.copy(), score.copy()
        self.gtop = (gv, gs)
        self.log.append("import_gtop")

    def import_migrants(self, ptr, stream=0):
        raw = ctypes.string_at(ptr, self.migrant_bytes())
        vec = np.frombuffer(raw[: self.n_vars * 4], dtype=np.int32)
        score = np.frombuffer(raw[self.n_vars * 4:], dtype=np.float64)
        # agent_base.rs:429-434: migrant replaces the current individual iff migrant <= current
        if self.o.score_le(score, self.score):
            self.vec, self.score = vec.copy(), score.copy()
        self.log.append("import")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    isl = FakeIslands(rank)
    mig = ring.RingMigrator(isl, rank, world, islands_per_rank=5, device="cpu")
    assert isl.external == (True, 5 * rank)
    assert (mig.dst, mig.src) == ((rank + 1) % world, (rank - 1) % world)
    res = {"rank": rank}
    # one exchange: rank r receives the individual of rank r-1, keeps it iff it is <= its own
    ring.run_steps(isl, mig, n_steps=3, migration_frequency=3)
    res["after1_vec0"] = int(isl.vec[0])
    res["log1"] = list(isl.log)
    # two more exchanges: the best individual (owned by the last rank) moves one hop per exchange
    ring.run_steps(isl, mig, n_steps=4, migration_frequency=2, first_step=0)
    res["after3_vec0"] = int(isl.vec[0])
    res["exchanges"] = mig.exchanges
    res["steps"] = isl.steps
    best, owner = ring.global_best(isl.score, world)
    res["best"], res["owner"] = best, owner
    out.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_exchange_and_global_best(world, oracle):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((out.get(timeout=120) for _ in range(world)), key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    last = world - 1
    for r in res:
        rank = r["rank"]
        # exchange #1: rank 0 receives from the last rank (the best) and takes it; every other
        # rank receives a WORSE individual (from rank-1) and keeps its own
        assert r["after1_vec0"] == (last if rank == 0 else rank)
        assert r["log1"] == ["step", "step", "step", "export", "import"]
        assert r["exchanges"] == 3 and r["steps"] == 7
        # after 3 exchanges the best individual has travelled 3 hops from the last rank
        reached = {(last + h) % world for h in range(0, 4)}
        assert r["after3_vec0"] == (last if rank in reached else rank)
        assert r["best"] == [0.0, 100.0 - 10.0 * last]
    # the best individual has spread around the ring, so several ranks tie; the lowest rank wins
    assert len({r["owner"] for r in res}) == 1
    assert res[0]["after3_vec0"] == last and res[0]["owner"] == 0


def _gtop_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    isl = FakeIslands(rank)
    mig = ring.RingMigrator(isl, rank, world, islands_per_rank=5, device="cpu", share_global_top=True)
    ring.run_steps(isl, mig, n_steps=2, migration_frequency=2)
    out.put({"rank": rank, "gtop_vec0": int(isl.gtop[0][0]), "gtop_score": isl.gtop[1].tolist(), "log": list(isl.log)})
    dist.barrier()
    dist.destroy_process_group()


def test_global_top_is_shared_across_ranks_in_one_exchange(oracle):
    """update_global_top across ranks (agent_base.rs:446-490): after ONE exchange every rank holds the
    best rank's individual as its global top (not one ring hop per exchange like the migrants)."""
    world = 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gtop_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((out.get(timeout=120) for _ in range(world)), key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r["gtop_vec0"] == world - 1 and r["gtop_score"] == [0.0, 100.0 - 10.0 * (world - 1)]
        assert r["log"] == ["step", "step", "export", "import", "export_gtop", "import_gtop"]


def test_single_rank_is_a_no_op():
    isl = FakeIslands(0)
    mig = ring.RingMigrator(isl, 0, 1, islands_per_rank=4, device="cpu")
    assert isl.external == (False, 0)
    ring.run_steps(isl, mig, 5, 2)
    assert isl.log == ["step"] * 5 and mig.exchanges == 0
    assert ring.global_best([1.0, 2.0], 1) == ([1.0, 2.0], 0)


def test_ring_helpers():
    assert ring.ring_neighbours(0, 4) == (1, 3)
    assert ring.ring_neighbours(3, 4) == (0, 2)
    assert ring.global_island_base(2, 148) == 296
