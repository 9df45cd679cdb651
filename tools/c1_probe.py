"""development probe: C1 (N-Queens-256 LateAcceptance chains) step timing"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greyjack-solver-rust_b200", "python"))
sys.path.insert(0, ROOT)
import torch
import greyjack_b200 as gj
from greyjack_b200 import instances as inst

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_isl = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
spl = int(sys.argv[3]) if len(sys.argv) > 3 else 50
spec = inst.nqueens(256, seed=45)
p = gj.Problem(spec)
isl = gj.LateAcceptance(32, 0.2, None, [0, 1.0, 0, 0, 0, 0], 100, scoring="delta",
                        chain_steps_per_launch=spl).build_agent(p, n_islands=n_isl, seed=1)
isl.step(3 * spl)
torch.cuda.synchronize()
t0 = time.perf_counter()
isl.step(calls * spl)
torch.cuda.synchronize()
t = time.perf_counter() - t0
print("C1 %s: %.2f us per step of all chains, %.1f M candidates/s, best %s" % (
    isl.step_path, 1e6 * t / (calls * spl), calls * spl * n_isl / t / 1e6, isl.best(-1)[1]))
