#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greyjack-solver-rust_b200", "python"))
import numpy as np
import greyjack_b200 as gj
from greyjack_b200 import instances as inst

ALL = [0.2, 0.16, 0.16, 0.16, 0.16, 0.16]
rng = np.random.default_rng(0)
for spec in (inst.nqueens(37), inst.tsp(61, seed=3), inst.cvrp(30, 4, seed=2), inst.vrptw(30, 4, n_depots=2, seed=3)):
    p = gj.Problem(spec)
    x = np.stack([spec.initial + rng.integers(-2, 3, size=spec.n_vars) for _ in range(9)])
    p.request_score_plain(x)
    deltas = [[(int(a), float(spec.initial[b]))] for a, b in rng.integers(0, spec.n_vars, size=(13, 2))]
    p.request_score_incremental(spec.initial, deltas)
    modes = ("full", "delta", "delta_unfused") if spec.kind <= inst.TSP else ("full",)
    for mode in modes:
        for tabu in (0.0, 0.3):
            ts = gj.TabuSearch(100, tabu, True, 1.5, ALL, 3, scoring=mode).build_agent(p, n_islands=3, seed=1)
            ts.step(7); ts.trace_step(1); ts.best(-1); ts.close()
            la = gj.LateAcceptance(5, tabu, None, ALL, 3, scoring=mode).build_agent(p, n_islands=5, seed=2)
            la.step(9); la.trace_step(2); la.best(-1); la.close()
        sa = gj.SimulatedAnnealing([1.0, 10.0, 10.0], 0.99, 0.2, None, ALL, 3, scoring=mode).build_agent(p, n_islands=3, seed=3)
        sa.step(9); sa.trace_step(0); sa.close()
    if spec.kind > inst.TSP:
        # VRP chains (route index, CTA re-index of migrated chains, global top's index) and VRP TabuSearch deltas
        small = [0.5, 0.5, 0.0, 0.0, 0.0, 0.0]
        la = gj.LateAcceptance(5, 0.2, None, small, 4, scoring="delta", chain_steps_per_launch=4).build_agent(p, n_islands=9, seed=2)
        la.step(13); la.trace_step(2); la.best(-1); la.close()
        ts = gj.TabuSearch(64, 0.2, True, None, small, 3, scoring="delta").build_agent(p, n_islands=3, seed=1)
        ts.step(5); ts.best(-1); ts.close()
    elif os.environ.get("GJ_SANITIZE_WIDE", "1") != "0":
        # the wide LateAcceptance launch shape (more than 20 chains per SM)
        import torch
        sms = torch.cuda.get_device_properties(0).multi_processor_count
        la = gj.LateAcceptance(5, 0.2, None, [0.0, 1.0, 0.0, 0.0, 0.0, 0.0] if spec.kind == inst.NQUEENS else ALL, 4, scoring="delta",
                               chain_steps_per_launch=4).build_agent(p, n_islands=21 * sms + 3, seed=2)
        la.step(9); la.best(-1); la.close()
    ga = gj.GeneticAlgorithm(50, 0.5, 0.2, 0.05, 1.0, None, 0.05, 2).build_agent(p, n_islands=2, seed=4)
    ga.step(5); ga.best(-1); ga.close()
    p.close()
print("sanitize probe done")
