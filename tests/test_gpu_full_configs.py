"""BASELINE.json's configurations at FULL size on the GPU, through the agents that the configs
name.  The oracle cannot score whole neighbourhoods at these sizes in seconds, so each test pairs
oracle spot checks (a handful of candidates / the carried individuals) with size-independent
properties: scores never get worse, the carried score is the oracle's score of the carried
vector, permutation moves keep a feasible solution feasible, counters add up."""
import numpy as np
import pytest

from greyjack_b200 import (GeneticAlgorithm, LateAcceptance, Problem, TabuSearch, instances as inst)
from test_gpu_islands import _same_score

pytestmark = pytest.mark.gpu


def test_c1_nqueens256_late_acceptance_single_agent(oracle):
    """C1: N-Queens N=256, LateAcceptance(size 32), one agent, swap moves (examples/nqueens/src/main.rs)."""
    spec = inst.nqueens(256, seed=45)
    op = oracle.OracleProblem(spec)
    for scoring in ("full", "delta"):
        gp = Problem(spec)
        isl = LateAcceptance(32, 0.2, None, [0, 1.0, 0, 0, 0, 0], 10, scoring=scoring).build_agent(gp, n_islands=1, seed=45)
        v0, s0 = isl.best(0)
        assert np.array_equal(s0, op.score_incremental(v0, [[]])[0])
        prev = s0
        for _ in range(4):
            isl.step(2500)
            v, s = isl.best(0)
            assert sorted(v.tolist()) == list(range(256))           # swaps keep the permutation
            assert np.array_equal(s, op.score_incremental(v, [[]])[0])   # integer level: bit-exact
            assert oracle.score_cmp(s, prev) <= 0
            prev = s
            cv, cs = isl.current(0)
            assert np.array_equal(cs, op.score_incremental(cv, [[]])[0])
        assert prev[0] < 0.6 * s0[0]                                 # 10 000 LA steps make real progress
        assert isl.stats() == {"candidates": 10000, "steps": 10000, "accepted": isl.stats()["accepted"]}
        isl.close(); gp.close()


def test_c3_cvrp2000_genetic_algorithm_pop8192(oracle):
    """C3: CVRP 2000 customers / 50 vehicles, GeneticAlgorithm population 8192 (one island here;
    8 islands on 8 GPUs in the scaling run)."""
    spec = inst.cvrp(2000, 50, seed=2, greedy=False)
    spec.initial = np.full(spec.n_vars, np.nan)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    ga = GeneticAlgorithm(8192, 0.5, 0.2, 0.05, 1.0, None, 0.00001, 10).build_agent(gp, n_islands=1, seed=2)
    prev = None
    for _ in range(3):
        ga.step(2)
        v, s = ga.best(0)
        assert _same_score(s, op.score_plain(v)[0], spec, oracle)    # PSC semantics, exact route folds
        if prev is not None:
            assert oracle.score_cmp(s, prev) <= 0
        prev = s
    assert ga.stats()["candidates"] == 6 * 8192
    ga.close(); gp.close()


def test_c4_vrptw5000_late_acceptance_islands(oracle):
    """C4: vrp_service time-window VRP, 5000 stops / 125 vehicles / 5 depots, LateAcceptance islands."""
    spec = inst.vrptw(5000, 125, n_depots=5, seed=3, service_variant=True, greedy=False)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    rng = np.random.default_rng(4)
    # scorer spot check at full size: plain (PSC) and incremental (ISC) forms, mixed hard/medium/soft
    x = np.stack([spec.initial.copy() for _ in range(4)])
    x[1:, 0::2] = rng.integers(0, 125, size=(3, 5000))
    got, want = gp.request_score_plain(x), op.score_plain(x)
    assert np.array_equal(got, want)
    deltas = [[(int(2 * i), float(rng.integers(0, 125)))] for i in rng.integers(0, 5000, size=16)]
    assert np.array_equal(gp.request_score_incremental(spec.initial, deltas), op.score_incremental(spec.initial, deltas))
    isl = LateAcceptance(32, 0.2, None, [0.5, 0.5, 0, 0, 0, 0], 50).build_agent(gp, n_islands=64, seed=3)
    s0 = isl.best(-1)[1]
    isl.step(300)
    v, s = isl.best(-1)
    assert _same_score(s, op.score_incremental(v, [[]])[0], spec, oracle)
    assert oracle.score_cmp(s, s0) < 0
    for i in (0, 31, 63):
        cv, cs = isl.current(i)
        assert _same_score(cs, op.score_incremental(cv, [[]])[0], spec, oracle)
    assert isl.stats()["candidates"] == 300 * 64
    isl.close(); gp.close()


def test_c5_tsp20000_ga_and_tabu_islands(oracle):
    """C5: TSP 20 000 cities (3.2 GB matrix built on the device from coordinates), a GA island group
    and a TabuSearch island group on the same problem, elite exchange between them."""
    spec = inst.tsp(20000, seed=4, with_matrix=False)
    gp = Problem(spec, use_coords=True)
    D = gp.distance_matrix()
    # the device-built matrix obeys the examples' formula (spot rows; full equality is tested small)
    rows = [0, 1, 9999, 19999]
    want = inst.distance_matrix(np.vstack([spec.coords[rows], spec.coords]))[:4, 4:]
    assert np.array_equal(D[rows], want)
    spec.distance_matrix = D
    op = oracle.OracleProblem(spec)
    ts = TabuSearch(4096, 0.2, True, None, [0, 0.5, 0, 0, 0, 0.5], 10, scoring="delta").build_agent(gp, n_islands=8, seed=4)
    tr = ts.trace_step(3)
    base = spec.initial
    want = oracle.score_round(op.score_incremental(base, tr["deltas"][:64]), spec.score_precision)
    assert np.array_equal(tr["scores"][:64, 0], want[:, 0])
    assert np.max(np.abs(tr["scores"][:64, 1] - want[:, 1])) <= 1.001e-3
    # the reference-facing incremental call and the full-evaluation islands at this size (one
    # 80 KB shared-memory clone per warp: the kernels drop to two warps per CTA)
    got = gp.request_score_incremental(base, tr["deltas"][:48])
    assert np.array_equal(got, op.score_incremental(base, tr["deltas"][:48]))
    tf = TabuSearch(64, 0.2, True, None, [0, 0.5, 0, 0, 0, 0.5], 10, scoring="full").build_agent(gp, n_islands=2, seed=4)
    trf = tf.trace_step(1)
    assert np.array_equal(trf["scores"], oracle.score_round(op.score_incremental(base, trf["deltas"]), spec.score_precision))
    tf.close()
    s0 = ts.best(-1)[1]
    ts.step(20)
    v, s = ts.best(-1)
    assert sorted(v.tolist()) == list(range(1, 20000))
    assert _same_score(s, op.score_incremental(v, [[]])[0], spec, oracle)
    assert oracle.score_cmp(s, s0) < 0
    ga = GeneticAlgorithm(1024, 0.5, 0.2, 0.0, 1.0, [0, 0.5, 0, 0, 0, 0.5], 0.01, 5).build_agent(
        gp, n_islands=2, seed=5, initial=np.stack([v, v]))
    ga.step(6)
    gv, gs = ga.best(-1)
    assert _same_score(gs, op.score_plain(gv)[0], spec, oracle)
    assert oracle.score_cmp(gs, s) <= 0        # seeded with the tabu islands' elite, never worse
    assert ga.stats()["candidates"] == 6 * 1024 * 2
    ga.close(); ts.close(); gp.close()
