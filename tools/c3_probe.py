"""development probe: C3 (CVRP-2000x50 GeneticAlgorithm, population 8192) generation timing"""
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

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
spec = inst.cvrp(2000, 50, seed=2, greedy=False)
spec.initial = np.full(spec.n_vars, np.nan)
p = gj.Problem(spec)
isl = gj.GeneticAlgorithm(8192, 0.5, 0.2, 0.05, 1.0, None, 0.00001, 10).build_agent(p, n_islands=1, seed=2)
isl.step(3)
torch.cuda.synchronize()
isl.set_profiling(True)
t0 = time.perf_counter()
isl.step(steps)
torch.cuda.synchronize()
t = time.perf_counter() - t0
ms, n = isl.profile_read()
print("C3 generation %.1f us, scorer kernel %.1f us, %.2f M candidates/s" % (1e6 * t / steps, 1e3 * ms / n, steps * 8192 / t / 1e6))

import ctypes
L = ctypes.CDLL(gj.LIB_PATH)
if hasattr(L, "gj_debug_vrp_phases"):
    buf = (ctypes.c_ulonglong * 16)()
    L.gj_debug_vrp_phases(buf, 1)
    isl.step(4)
    torch.cuda.synchronize()
    L.gj_debug_vrp_phases(buf, 0)
    names = ["load", "apply", "-", "zero", "pass1", "o5", "o6", "o7", "scan", "scatter", "legs", "fold", "final", "store"]
    tot = sum(buf)
    print("phase cycles per CTA (thread 0):", ", ".join("%s %.0f" % (nm, buf[i] / (4 * 8192)) for i, nm in enumerate(names) if buf[i]),
          "| total %.0f" % (tot / (4 * 8192)))
