import os, sys
sys.path.insert(0, "/root/repo/greyjack-solver-rust_b200/python")
import greyjack_b200 as gj
from greyjack_b200 import instances as inst
import torch
spec = inst.vrptw(5000, 125, n_depots=5, seed=3, service_variant=True, greedy=False)
prob = gj.Problem(spec)
isl = gj.LateAcceptance(32, 0.2, None, [0.5, 0.5, 0, 0, 0, 0], 50, scoring="delta", chain_steps_per_launch=64).build_agent(prob, n_islands=4096, seed=1)
isl.step(64 * 10)
torch.cuda.synchronize()
print(isl.step_path, isl.best(-1)[1])
