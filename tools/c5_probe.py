"""development probe: C5 (TSP-20000 TabuSearch, 4096 moves per step, lean fused step) step timing"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greyjack-solver-rust_b200", "python"))
sys.path.insert(0, ROOT)
import torch
import greyjack_b200 as gj
from greyjack_b200 import instances as inst

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_isl = int(sys.argv[2]) if len(sys.argv) > 2 else 148
exact = int(sys.argv[3]) if len(sys.argv) > 3 else 0
spec = inst.tsp(20000, seed=4, with_matrix=False)
p = gj.Problem(spec, use_coords=True)
p.set_exact_sums(bool(exact))
isl = gj.TabuSearch(4096, 0.2, True, None, [0.0, 0.5, 0.0, 0.0, 0.0, 0.5], 10, scoring="delta").build_agent(p, n_islands=n_isl, seed=4)
isl.step(3)
torch.cuda.synchronize()
t0 = time.perf_counter()
isl.step(steps)
torch.cuda.synchronize()
t = time.perf_counter() - t0
print("C5 %s islands=%d exact=%d: %.1f us per step, %.2f G candidates/s" % (isl.step_path, n_isl, exact, 1e6 * t / steps, steps * n_isl * 4096 / t / 1e9))
isl.close()
