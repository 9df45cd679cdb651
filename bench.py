#!/usr/bin/env python
"""bench.py -- candidate moves scored per second on BASELINE.json's headline configuration.

Workload (config C2): TSP, 1 000 synthetic cities, TabuSearch, swap / 2-opt neighbourhood of
4 096 moves per step per island.  A "step" is one TabuSearch step of every island resident on
the GPU (move generation -> scoring -> selection, all on the device).

  value  : candidates scored / s, whole job, islands resident in HBM (device-timed, CUDA events)
  e2e    : the same metric through the reference-facing call gj_score_incremental
           (OOPScoreRequester::request_score_incremental) with HOST buffers: every step copies
           the base + the delta lists host->device and the scores device->host.
  cpu_baseline / --impl reference : the oracle's restatement of the reference's own CPU path
           (one island per host thread, pseudo-incremental scoring), timed on this box.

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "greyjack-solver-rust_b200", "python"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_CITIES = 1000
NEIGHBOURS = 4096
MOVE_PROBAS = [0.0, 0.5, 0.0, 0.0, 0.0, 0.5]       # swap + inverse (2-opt)
TABU_RATE = 0.5                                     # examples/tsp/src/main.rs:47
MIGRATION_FREQUENCY = 10
METRIC = "candidate moves scored/sec (whole box)"
UNIT = "candidates/s"
# algorithmic bytes per candidate
A_FULL = 4 * (N_CITIES - 1) + 8 * N_CITIES + 16     # 12 012 B: full (pseudo-incremental) evaluation
A_DELTA = 0.5 * 96 + 0.5 * 56                       # swap 96 B / 2-opt 56 B, 50:50 mix
SCORING_DESC = {
    "full": "full re-evaluation of base+move per candidate (the reference's pseudo-incremental ISC semantics)",
    "delta": "fused island step: move generated in registers, delta-scored against the island's state staged in "
             "shared memory (tour, edge lengths, value counts, tabu table), added edges gathered from the "
             "L2-resident matrix; selection, apply, full re-score of the accepted neighbour and tabu update in "
             "the same kernel",
}
SCORING_KERNEL = {"full": "k_score_moves_warp<GJ_TSP>", "delta": "k_ls_step_fused<GJ_TSP,256>"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed regions (B200_PROFILING.md): one
    `nvidia-smi -lms 50` process streams samples while the bench runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                line = line.strip()
                if line:
                    self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def summary(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.join(timeout=6)
        rows = [r for r in self.rows if len(r) >= 6]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(rows[0][1]) if rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(rows)}


def host_moves(base, n_moves, rng):
    """swap / 2-opt delta lists in the reference's incremental form (mover.rs:180-219, 378-420),
    built with numpy on the host -- the input of the e2e (host-buffer) leg."""
    n = len(base)
    offs = np.zeros(n_moves + 1, dtype=np.uint64)
    ids, vals = [], []
    for j in range(n_moves):
        a, b = rng.choice(n, size=2, replace=False)
        if rng.random() < 0.5:
            ids.append(np.array([a, b], dtype=np.uint64))
            vals.append(np.array([base[b], base[a]]))
        else:
            lo, hi = (a, b) if a < b else (b, a)
            ids.append(np.arange(lo, hi + 1, dtype=np.uint64))
            vals.append(base[lo:hi + 1][::-1].copy())
        offs[j + 1] = offs[j] + len(ids[-1])
    return offs, np.concatenate(ids), np.concatenate(vals).astype(np.float64)


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port), all host threads."""
    if rank != 0:
        return
    from greyjack_b200 import instances as inst
    from oracle import gj_oracle
    spec = inst.tsp(N_CITIES, seed=1)
    op = gj_oracle.OracleProblem(spec)
    cores = os.cpu_count() or 1
    base = spec.initial
    for _ in range(args.warmup):
        op.bench_ts(base, NEIGHBOURS, 1, cores, 7, MOVE_PROBAS, [3, 3])
    # each bench step = one TabuSearch step of `cores` islands (one per thread, solver.rs:94)
    n, secs, best = op.bench_ts(base, NEIGHBOURS, args.steps, cores, 11, MOVE_PROBAS, [3, 3])
    value = n / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C2: TSP 1000 cities, TabuSearch, swap/2-opt, 4096 moves per step",
                   "islands": cores, "moves_per_island_step": NEIGHBOURS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} TabuSearch steps x {cores} islands x {NEIGHBOURS} moves; "
                                   "oracle port of the reference ISC path WITHOUT its Polars marshalling "
                                   "(CPU-favouring)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def other_configs(gj, inst, torch):
    """BASELINE.json's other configurations (parity-test cases, not bench lines): device-timed
    candidates/s of the agent each one names, islands resident in HBM, no L2 flush.  Reported for
    context next to the headline; a failure here never touches the headline numbers."""
    out = {}
    stream = torch.cuda.current_stream().cuda_stream

    def timed(name, make, steps):
        try:
            prob, isl = make()
            isl.step(max(1, steps // 10), stream)
            torch.cuda.synchronize()
            c0 = isl.stats()["candidates"]
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); isl.step(steps, stream); b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            out[name] = {"candidates_per_s": (isl.stats()["candidates"] - c0) / (ms * 1e-3),
                         "us_per_step": 1e3 * ms / steps, "best": [float(x) for x in isl.best(-1)[1]]}
            isl.close(); prob.close()
        except Exception as e:                      # noqa: BLE001
            out[name] = {"error": str(e)[:200]}

    def c1():
        p = gj.Problem(inst.nqueens(256, seed=45))
        return p, gj.LateAcceptance(32, 0.2, None, [0, 1.0, 0, 0, 0, 0], 100, scoring="delta").build_agent(p, n_islands=4096, seed=1)

    def c3():
        spec = inst.cvrp(2000, 50, seed=2, greedy=False)
        spec.initial = np.full(spec.n_vars, np.nan)
        p = gj.Problem(spec)
        return p, gj.GeneticAlgorithm(8192, 0.5, 0.2, 0.05, 1.0, None, 0.00001, 10).build_agent(p, n_islands=1, seed=2)

    def c4():
        p = gj.Problem(inst.vrptw(5000, 125, n_depots=5, seed=3, service_variant=True, greedy=False))
        return p, gj.LateAcceptance(32, 0.2, None, [0.5, 0.5, 0, 0, 0, 0], 50, scoring="delta",
                                    chain_steps_per_launch=64).build_agent(p, n_islands=4096, seed=3)

    def c5():
        p = gj.Problem(inst.tsp(20000, seed=4, with_matrix=False), use_coords=True)
        p.set_exact_sums(False)
        return p, gj.TabuSearch(4096, 0.2, True, None, MOVE_PROBAS, 10, scoring="delta").build_agent(p, n_islands=148, seed=4)

    timed("C1 nqueens-256 LateAcceptance x4096 chains (k_la_chains)", c1, 1000)
    timed("C3 cvrp-2000x50 GeneticAlgorithm pop 8192 x1 island", c3, 10)
    timed("C4 vrptw-5000 (vrp_service) LateAcceptance x4096 chains (k_vrp_chains, route-level delta)", c4, 640)
    timed("C5 tsp-20000 TabuSearch 4096 moves x148 islands (fused step, lean layout)", c5, 20)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--islands", type=int, default=int(os.environ.get("GJ_BENCH_ISLANDS", "592")))
    ap.add_argument("--e2e-agents", type=int, default=4)
    ap.add_argument("--scoring", default="delta", choices=["delta", "full"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="development: device-timed value only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import greyjack_b200 as gj
    from greyjack_b200 import instances as inst

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    spec = inst.tsp(N_CITIES, seed=1)
    prob = gj.Problem(spec, device=local_rank)
    prob.set_exact_sums(False)
    SCORING = args.scoring
    builder = gj.TabuSearch(NEIGHBOURS, TABU_RATE, True, None, MOVE_PROBAS, MIGRATION_FREQUENCY, scoring=SCORING)
    isl = builder.build_agent(prob, n_islands=args.islands, seed=1000 + rank)
    from greyjack_b200 import ring
    migrator = ring.RingMigrator(isl, rank, world, args.islands, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def one_step(i):
        isl.step(1, stream)
        if (i + 1) % MIGRATION_FREQUENCY == 0:
            # AgentToAgentUpdate ring i -> i+1 across GPUs (agent_base.rs:161-183) over NCCL
            migrator.exchange(stream)

    for i in range(args.warmup):
        one_step(i)
    migrator.exchange(stream)      # untimed: sets up the NCCL point-to-point connections (no-op at N=1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    isl.set_profiling(True)
    c0 = isl.stats()["candidates"]
    launches0 = gj.load().gj_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for i in range(args.steps):
        if not os.environ.get("GJ_BENCH_NO_FLUSH"):      # development only: the reported protocol flushes
            flush.fill_(i & 0xFF)                # L2 flush between timed iterations (untimed)
        evs[i][0].record()
        one_step(args.warmup + i)
        evs[i][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    cands = isl.stats()["candidates"] - c0
    gpu_launches = int(gj.load().gj_launch_count() - launches0)     # counted by the library itself
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        c = torch.tensor([cands], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        cands = float(c.item())
    value = cands / (total_ms * 1e-3)

    # dominant kernel (the scorer): CUDA events on its launch stream, inside the timed region
    k_total_ms, k_launches = isl.profile_read()
    isl.set_profiling(False)
    kernel_ms = k_total_ms / max(1, k_launches)
    per_launch_cands = args.islands * NEIGHBOURS
    a_bytes = A_FULL if SCORING == "full" else A_DELTA
    peak, peak_src = peaks()
    achieved = per_launch_cands * a_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: TSP 1000 cities, TabuSearch, swap/2-opt, 4096 moves per step",
                   "islands_per_gpu": args.islands, "moves_per_island_step": NEIGHBOURS,
                   "move_probas": MOVE_PROBAS, "tabu_entity_rate": TABU_RATE,
                   "migration_frequency": MIGRATION_FREQUENCY, "score_precision": [3, 3],
                   "scoring": SCORING_DESC[SCORING], "float_sums": "tree (gj_problem_set_exact_sums(0))", "l2": "flushed between timed steps (256 MiB write)",
                   "parallelism": f"islands x{world}"},
        "clocks": None, "gpu_launches": gpu_launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel": SCORING_KERNEL[SCORING], "kernel_ms": kernel_ms,
                     "kernel_share_of_step": kernel_ms / (total_ms / args.steps) if world == 1 else None,
                     "algorithmic_bytes_per_candidate": a_bytes,
                     "candidates_per_launch": per_launch_cands},
    }

    if args.no_e2e:
        line["clocks"] = sampler.summary()
        line["step_path"] = isl.step_path
        if rank == 0:
            print(json.dumps(line))
        return
    if rank != 0:
        sampler.summary()
    if rank == 0:
        # ---- e2e: reference-facing call, host buffers, H2D + D2H inside the timed region ------
        A = args.e2e_agents
        rng = np.random.default_rng(5)
        base = spec.initial.copy()
        probs = [prob] + [gj.Problem(spec, device=local_rank) for _ in range(A - 1)]
        # inputs and outputs live in page-locked host memory (gj_host_alloc): the copies inside
        # gj_score_incremental are then asynchronous DMA at PCIe speed
        keep = []

        def pin(a):
            holder, view = gj.pinned_copy(a)
            keep.append(holder)
            return view

        base = pin(base)
        sets = [tuple(pin(x) for x in host_moves(base, NEIGHBOURS, rng)) for _ in range(A)]
        outs = [pin(np.empty((NEIGHBOURS, 2))) for _ in range(A)]
        h2d = sum(base.nbytes + o.nbytes + i.nbytes + v.nbytes for o, i, v in sets)
        d2h = sum(o.nbytes for o in outs)

        def agent(a, n):
            o, i, v = sets[a]
            for _ in range(n):
                probs[a].request_score_incremental_csr(base, o, i, v, out=outs[a])

        def run_agents(n):
            th = [threading.Thread(target=agent, args=(a, n)) for a in range(A)]
            for t in th:
                t.start()
            for t in th:
                t.join()

        run_agents(args.warmup)
        l0 = gj.load().gj_launch_count()
        t0 = time.perf_counter()
        run_agents(args.steps)
        e2e_s = time.perf_counter() - t0
        e2e_launches = int(gj.load().gj_launch_count() - l0)
        # the sampler has been running since before the device-timed region: both timed regions
        line["clocks"] = sampler.summary()
        line["e2e"] = {"value": A * NEIGHBOURS * args.steps / e2e_s, "unit": UNIT,
                       "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "agents": A, "host_memory": "pinned (gj_host_alloc)",
                       "call": "gj_score_incremental (request_score_incremental), host CSR deltas, "
                               "one call per agent per step, agents on separate host threads / CUDA streams"}
        line["gpu_launches"] = gpu_launches + e2e_launches      # both timed regions, counted by the library

        # ---- additional figure: the same call with the packed wire format (u32 ids, i32 values) ----
        psets = []
        for o, i, v in sets:
            psets.append((o, pin(i.astype(np.uint32)), pin(np.rint(v).astype(np.int32))))
        ph2d = sum(base.nbytes + o.nbytes + i.nbytes + v.nbytes for o, i, v in psets)

        def agent_packed(a, n):
            o, i, v = psets[a]
            for _ in range(n):
                probs[a].request_score_incremental_packed(base, o, i, v, out=outs[a])

        def run_packed(n):
            th = [threading.Thread(target=agent_packed, args=(a, n)) for a in range(A)]
            for t in th:
                t.start()
            for t in th:
                t.join()

        run_packed(args.warmup)
        t0 = time.perf_counter()
        run_packed(args.steps)
        p_s = time.perf_counter() - t0
        line["e2e"]["packed"] = {"value": A * NEIGHBOURS * args.steps / p_s, "unit": UNIT,
                                 "h2d_bytes_per_step": int(ph2d),
                                 "call": "gj_score_incremental_packed: 8 B per delta instead of the "
                                         "reference layout's 16 B (not the headline)"}

        # ---- cpu_baseline: oracle port on the host cores, bounded sample ------------------------
        if world == 1 and not args.no_cpu_baseline:
            from oracle import gj_oracle
            op = gj_oracle.OracleProblem(spec)
            cores = os.cpu_count() or 1
            n1, s1, _ = op.bench_ts(base, NEIGHBOURS, 1, cores, 3, MOVE_PROBAS, [3, 3])
            steps = max(2, min(800, int(12.0 / max(s1, 1e-3))))
            n, secs, _ = op.bench_ts(base, NEIGHBOURS, steps, cores, 4, MOVE_PROBAS, [3, 3])
            line["cpu_baseline"] = {
                "value": n / secs, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{steps} TabuSearch steps x {cores} islands x {NEIGHBOURS} moves ({secs:.1f} s); "
                          "oracle port of the reference ISC path without its Polars marshalling (CPU-favouring)"}
            # ---- metric (ii): best score at fixed wall time, same instance, same move mix ---------
            wall = float(os.environ.get("GJ_BENCH_QUALITY_WALL_S", "2.0"))
            q_isl = builder.build_agent(prob, n_islands=args.islands, seed=4242)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            q_steps = 0
            while time.perf_counter() - t0 < wall:
                q_isl.step(MIGRATION_FREQUENCY, stream)
                torch.cuda.synchronize()
                q_steps += MIGRATION_FREQUENCY
            gpu_wall = time.perf_counter() - t0
            _, gpu_best = q_isl.best(-1)
            q_isl.close()
            cpu_steps = max(1, int(wall * (n / secs) / (cores * NEIGHBOURS)))
            _, cpu_wall, cpu_best = op.bench_ts(base, NEIGHBOURS, cpu_steps, cores, 5, MOVE_PROBAS, [3, 3])
            line["quality"] = {
                "metric": "best score (hard, soft) at fixed wall time; lower is better",
                "start": [float(x) for x in op.score_incremental(base, [[]])[0]],
                "gpu": {"wall_s": gpu_wall, "best": [float(x) for x in gpu_best], "steps": q_steps,
                        "islands": args.islands},
                "cpu": {"wall_s": cpu_wall, "best": [float(x) for x in cpu_best], "steps": cpu_steps,
                        "islands": cores, "note": "oracle port, one TabuSearch island per host thread"}}
        if world == 1 and os.environ.get("GJ_BENCH_EXTRAS", "1") != "0":
            line["other_configs"] = other_configs(gj, inst, torch)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
