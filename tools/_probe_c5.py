import os, sys
sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo/greyjack-solver-rust_b200/python")
import step_probe as sp
import greyjack_b200 as gj
from greyjack_b200 import instances as inst
import torch
spec = inst.tsp(20000, seed=4, with_matrix=False)
P2 = [0.0, 0.5, 0.0, 0.0, 0.0, 0.5]
def run(name, builder, islands, steps):
    prob = gj.Problem(spec, use_coords=True)
    prob.set_exact_sums(False)
    isl = builder.build_agent(prob, n_islands=islands, seed=1)
    st = torch.cuda.current_stream().cuda_stream
    isl.step(3, st); torch.cuda.synchronize()
    c0 = isl.stats()["candidates"]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); isl.step(steps, st); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b); c = isl.stats()["candidates"] - c0
    print(f"{name:<44} {1e3*ms/steps:9.1f} us/step  {c/ms/1e6:9.3f} G cand/s  best={isl.best(-1)[1]}", flush=True)
    isl.close(); prob.close()
run("tsp20000 TS K=4096 I=148 delta", gj.TabuSearch(4096, 0.2, True, None, P2, 10, scoring="delta"), 148, 30)
run("tsp20000 TS K=4096 I=592 delta", gj.TabuSearch(4096, 0.2, True, None, P2, 10, scoring="delta"), 592, 20)
run("tsp20000 TS K=4096 I=16 full", gj.TabuSearch(4096, 0.2, True, None, P2, 10, scoring="full"), 16, 3)
run("tsp20000 GA pop=1024 I=8", gj.GeneticAlgorithm(1024, 0.5, 0.2, 0.0, 1.0, P2, 0.01, 5), 8, 10)
