#!/usr/bin/env python
"""Times gj_islands_step for a few configurations (CUDA events, no L2 flush) -- a development
probe, not a bench line.  usage: python tools/step_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greyjack-solver-rust_b200", "python"))
import torch
import greyjack_b200 as gj
from greyjack_b200 import instances as inst

def run(name, spec, builder, islands, steps=50, exact=False):
    prob = gj.Problem(spec)
    prob.set_exact_sums(exact)
    isl = builder.build_agent(prob, n_islands=islands, seed=1)
    st = torch.cuda.current_stream().cuda_stream
    isl.step(5, st); torch.cuda.synchronize()
    c0 = isl.stats()["candidates"]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); isl.step(steps, st); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    c = isl.stats()["candidates"] - c0
    print(f"{name:<44} {1e3*ms/steps:9.1f} us/step  {c/ms/1e6:9.3f} G cand/s  best={isl.best(-1)[1]}", flush=True)
    isl.close(); prob.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["tsp"]
    if "tsp" in which:
        spec = inst.tsp(1000, seed=1)
        P2 = [0.0, 0.5, 0.0, 0.0, 0.0, 0.5]
        for sc in ("delta", "delta_unfused", "full"):
            run(f"tsp1000 TS K=4096 I=148 {sc}", spec, gj.TabuSearch(4096, 0.5, True, None, P2, 10, scoring=sc), 148, steps=20 if sc == "full" else 50)
        run("tsp1000 TS K=4096 I=148 delta exact", spec, gj.TabuSearch(4096, 0.5, True, None, P2, 10, scoring="delta"), 148, exact=True)
        run("tsp1000 TS K=4096 I=148 delta notabu", spec, gj.TabuSearch(4096, 0.0, True, None, P2, 10, scoring="delta"), 148)
        run("tsp1000 TS K=1024 I=148 delta", spec, gj.TabuSearch(1024, 0.5, True, None, P2, 10, scoring="delta"), 148)
        run("tsp1000 TS K=1024 I=592 delta", spec, gj.TabuSearch(1024, 0.5, True, None, P2, 10, scoring="delta"), 592)
        run("tsp1000 TS K=4096 I=296 delta", spec, gj.TabuSearch(4096, 0.5, True, None, P2, 10, scoring="delta"), 296)
        run("tsp1000 TS K=16384 I=148 delta", spec, gj.TabuSearch(16384, 0.5, True, None, P2, 10, scoring="delta"), 148)
        run("tsp1000 TS K=4096 I=148 delta mix", spec, gj.TabuSearch(4096, 0.5, True, None, [0, .2, .2, .2, .2, .2], 10, scoring="delta"), 148)
        run("tsp1000 LA I=4096 delta", spec, gj.LateAcceptance(32, 0.2, None, P2, 100, scoring="delta"), 4096, steps=200)
    if "nq" in which:
        spec = inst.nqueens(256)
        SW = [0, 1.0, 0, 0, 0, 0]
        run("nq256 TS K=4096 I=148 delta", spec, gj.TabuSearch(4096, 0.2, True, None, SW, 10, scoring="delta"), 148)
        run("nq256 LA I=4096 delta", spec, gj.LateAcceptance(32, 0.2, None, SW, 100, scoring="delta"), 4096, steps=200)
        run("nq256 LA I=4096 full", spec, gj.LateAcceptance(32, 0.2, None, SW, 100, scoring="full"), 4096, steps=200)
    if "vrp" in which:
        spec = inst.cvrp(2000, 50, seed=2, greedy=False)
        import numpy as np
        spec.initial = np.full(spec.n_vars, np.nan)
        run("cvrp2000x50 GA pop=8192 I=1", spec, gj.GeneticAlgorithm(8192, 0.5, 0.2, 0.05, 1.0, None, 0.00001, 10), 1, steps=20)
        run("cvrp2000x50 GA pop=8192 I=8", spec, gj.GeneticAlgorithm(8192, 0.5, 0.2, 0.05, 1.0, None, 0.00001, 10), 8, steps=10)
        spec = inst.cvrp(2000, 50, seed=2, greedy=True)
        run("cvrp2000x50 TS K=4096 I=8 full", spec, gj.TabuSearch(4096, 0.2, True, None, [0.5, 0.5, 0, 0, 0, 0], 10), 8, steps=10)
        spec = inst.vrptw(5000, 125, n_depots=5, seed=3, service_variant=True, greedy=False)
        run("vrptw5000 LA I=592 full", spec, gj.LateAcceptance(32, 0.2, None, [0.5, 0.5, 0, 0, 0, 0], 50), 592, steps=50)
        run("vrptw5000 LA I=4096 full", spec, gj.LateAcceptance(32, 0.2, None, [0.5, 0.5, 0, 0, 0, 0], 50), 4096, steps=20)
