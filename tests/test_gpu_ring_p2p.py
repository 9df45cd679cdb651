"""The cross-GPU ring and shared global top on REAL peer memory (csrc/gj_ring.cu) and over NCCL
(ring.RingMigrator): two ranks on two GPUs of one box, one process each.  Skipped on a single-GPU
box (the driver's test run); run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_ring_p2p.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, transport, out):
    import torch.distributed as dist
    from greyjack_b200 import Problem, TabuSearch, instances as inst, ring
    from oracle import gj_oracle as oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    spec = inst.tsp(100, seed=4)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec, device=rank)
    I = 4
    # rank 0 starts from the greedy tour, rank 1 from the identity tour (much worse)
    start = spec.initial if rank == 0 else np.arange(1, 100, dtype=np.float64)
    isl = TabuSearch(64, 0.0, True, None, [0, 0.5, 0, 0, 0, 0.5], 1, scoring="delta").build_agent(
        gp, n_islands=I, seed=50 + rank, initial=np.stack([start] * I))
    stream = torch.cuda.current_stream().cuda_stream
    if transport == "p2p":
        mig = ring.PeerRing(isl, rank, world, I)
    else:
        mig = ring.RingMigrator(isl, rank, world, I, device="cuda", share_global_top=True)
    res = {"rank": rank}
    s0 = isl.best(-1)[1]
    isl.step(1, stream)
    mig.exchange(stream)
    torch.cuda.synchronize()
    gv, gs = isl.best(-1)
    res["gtop_after_1"] = gs.tolist()
    res["gtop_is_scored"] = bool(np.array_equal(gs, oracle.score_round(op.score_incremental(gv, [[]])[0], spec.score_precision))
                                 or np.array_equal(gs, op.score_incremental(gv, [[]])[0]))
    res["cur0_after_1"] = isl.current(0)[1].tolist()
    for k in range(6):
        isl.step(1, stream)
        mig.exchange(stream)
    torch.cuda.synchronize()
    res["gtop_end"] = isl.best(-1)[1].tolist()
    res["tops_end"] = [isl.best(i)[1].tolist() for i in range(I)]
    res["start"] = s0.tolist()
    if transport == "p2p":
        res["stats"] = mig.stats()
        mig.close()
    out.put(res)
    dist.barrier()
    isl.close(); gp.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_two_gpu_ring_and_global_top(transport):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, transport, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((out.get(timeout=300) for _ in range(2)), key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0, r1 = res
    # one exchange: rank 1's global top is now rank 0's (strictly better than anything rank 1 had),
    # and it is a real scored individual
    assert r1["gtop_after_1"] == r0["gtop_after_1"]
    assert r0["gtop_is_scored"] and r1["gtop_is_scored"]
    assert r1["gtop_after_1"][1] < r1["start"][1]
    # rank 1's first island received rank 0's last island's individual through the ring
    assert r1["cur0_after_1"][1] < r1["start"][1]
    # compare_to_global: after a few more steps every island of rank 1 has adopted / improved on it
    for t in r1["tops_end"]:
        assert t[1] <= r1["gtop_after_1"][1]
    assert r0["gtop_end"] == r1["gtop_end"]
    if transport == "p2p":
        assert r0["stats"] == {"exchanges": 7, "missed": 0} and r1["stats"] == {"exchanges": 7, "missed": 0}
