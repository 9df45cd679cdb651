import os, sys
sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo/greyjack-solver-rust_b200/python")
import greyjack_b200 as gj
from greyjack_b200 import instances as inst
import torch
spec = inst.vrptw(5000, 125, n_depots=5, seed=3, service_variant=True, greedy=False)
P2 = [0.5, 0.5, 0.0, 0.0, 0.0, 0.0]
def run(name, builder, islands, steps):
    prob = gj.Problem(spec)
    isl = builder.build_agent(prob, n_islands=islands, seed=1)
    st = torch.cuda.current_stream().cuda_stream
    isl.step(max(8, steps // 10), st); torch.cuda.synchronize()
    c0 = isl.stats()["candidates"]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); isl.step(steps, st); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b); c = isl.stats()["candidates"] - c0
    print(f"{name:<52} [{isl.step_path}] {1e3*ms/steps:9.1f} us/step  {c/ms/1e3:9.3f} M cand/s  best={isl.best(-1)[1]}", flush=True)
    isl.close(); prob.close()
for I, spl, mig in ((4096, 128, 50), (4096, 128, 1000000), (8192, 128, 1000000), (16384, 128, 1000000), (16384, 32, 50)):
    run(f"C4 LA I={I} steps/launch={spl} mig={mig}", gj.LateAcceptance(32, 0.2, None, P2, mig, scoring="delta", chain_steps_per_launch=spl), I, 512)
