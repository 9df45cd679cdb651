"""development probe: C4 (VRPTW-5000 vrp_service, LateAcceptance chains) step timing"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greyjack-solver-rust_b200", "python"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import greyjack_b200 as gj
from greyjack_b200 import instances as inst

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n_isl = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
spl = int(sys.argv[3]) if len(sys.argv) > 3 else 64
spec = inst.vrptw(5000, 125, n_depots=5, seed=3, service_variant=True, greedy=False)
p = gj.Problem(spec)
isl = gj.LateAcceptance(32, 0.2, None, [0.5, 0.5, 0, 0, 0, 0], 64, scoring="delta",
                        chain_steps_per_launch=spl).build_agent(p, n_islands=n_isl, seed=3)
isl.step(3 * spl)
torch.cuda.synchronize()
t0 = time.perf_counter()
isl.step(calls * spl)
torch.cuda.synchronize()
t = time.perf_counter() - t0
print("C4 %s: %.1f us per step of all chains, %.2f M candidates/s, best %s" % (
    isl.step_path, 1e6 * t / (calls * spl), calls * spl * n_isl / t / 1e6, isl.best(-1)[1]))
