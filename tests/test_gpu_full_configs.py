"""BASELINE.json's configurations at FULL size on the GPU, through the agents that the configs
name.  The oracle cannot score whole neighbourhoods at these sizes in seconds, so each test pairs
oracle spot checks (a handful of candidates / the carried individuals) with size-independent
properties: scores never get worse, the carried score is the oracle's score of the carried
vector, permutation moves keep a feasible solution feasible, counters add up."""
import numpy as np
import pytest

from greyjack_b200 import (GeneticAlgorithm, LateAcceptance, Problem, TabuSearch, instances as inst)
from test_gpu_islands import _final_state, _oracle_move, _same_score
from test_gpu_delta import QUANTUM, _check_delta_scores

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


def _selected_ids(d):
    """positions a move pushed into the tabu deque (Mover::select_non_tabu_ids, mover.rs:75-96)"""
    kind, k = int(d[0]), int(d[2])
    if kind == 255:
        return []
    if kind == 3:
        return [int(d[4])]
    return [int(x) for x in d[4:4 + (2 if kind >= 4 else k)]]


def _replay_ts_step(isl, island, spec, op, oracle, exact):
    """One traced TabuSearch step of `island`, everything replayed through the oracle: the K moves
    (mover.rs:145-421), their scores (ISC scorer + ScoreTrait::round), the selection
    (tabu_search_base.rs:157-188), the stored individual and the tabu deque (mover.rs:75-96)."""
    K = isl.K
    base, cur_score = isl.current(island)
    full = oracle.score_round(op.score_incremental(base, [[]])[0], spec.score_precision)
    unr = op.score_incremental(base, [[]])[0]
    assert cur_score[0] in (full[0], unr[0]) and min(abs(cur_score[1] - full[1]), abs(cur_score[1] - unr[1])) <= (0.0 if exact else QUANTUM)
    tabu_before, T = isl.trace_tabu(island, 0)
    assert T == int(np.ceil(0.5 * spec.n_vars))
    tr = isl.trace_step(island)
    banned = set(tabu_before.tolist())
    pushed = []
    for j in range(K):
        d = tr["desc"][j]
        want = _oracle_move(op, spec, base, d)
        assert _final_state(spec.n_vars, tr["deltas"][j]) == _final_state(spec.n_vars, want), (island, j, d)
        sel = _selected_ids(d)
        assert not (set(sel) & banned), (island, j, sel)       # every neighbour sees the step-start deque
        pushed += sel
    _check_delta_scores(tr["scores"], op.score_incremental(base, tr["deltas"]), spec, oracle)
    sel, acc = oracle.ts_select(tr["scores"], cur_score)
    assert (tr["selected"], tr["accepted"]) == (sel, acc)
    new, new_score = isl.current(island)
    want_vec = base.copy()
    if acc:
        for c, v in tr["deltas"][sel]:
            want_vec[c] = v
        want_score = oracle.score_round(op.score_incremental(new, [[]])[0], spec.score_precision)
        assert new_score[0] == want_score[0]
        if exact:
            assert np.array_equal(new_score, want_score)        # stored score = full evaluation, reference order
        else:
            assert abs(new_score[1] - want_score[1]) <= QUANTUM
    else:
        assert np.array_equal(new_score, cur_score)
    assert np.array_equal(new, want_vec)
    # the deque advanced once, by the step's ids in candidate order (newest first), older ids behind
    tabu_after, _ = isl.trace_tabu(island, 0)
    assert tabu_after.tolist() == (pushed[::-1] + tabu_before.tolist())[:T]
    return acc


@pytest.mark.parametrize("scoring", ["delta", "delta_f64"], ids=["fixed-point", "f64"])
@pytest.mark.parametrize("exact", [True, False], ids=["exact-sums", "tree-sums"])
def test_c2_tsp1000_tabu_fused_bench_shape(exact, scoring, oracle):
    """C2 exactly as bench.py runs it: TSP-1000 seed 1, TabuSearch(4096 neighbours, tabu 0.5,
    compare_to_global, swap + 2-opt, migration every 10), 2368 islands per GPU (bench.ISLANDS_PER_GPU:
    four waves of CTAs), the fused step with 16 neighbours per thread -- in fixed point (what bench.py
    times) and in f64 -- traced on the first, a middle and the last island, before and after a
    migration / global-top adoption."""
    spec = inst.tsp(1000, seed=1)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    gp.set_exact_sums(exact)
    isl = TabuSearch(4096, 0.5, True, None, [0.0, 0.5, 0.0, 0.0, 0.0, 0.5], 10, scoring=scoring).build_agent(
        gp, n_islands=2368, seed=1000)
    assert isl.step_path == ("fused_fixed" if scoring == "delta" else "fused")
    islands = (0, 1183, 2367)
    accepted = 0
    for i in islands:                                   # fresh islands, empty deques
        accepted += _replay_ts_step(isl, i, spec, op, oracle, exact)
    isl.step(10)                                        # one ring migration + ten global-top publications
    g_vec, g_score = isl.best(-1)
    # the published global top is adopted by every island whose own top it beats (compare_to_global)
    adopted = sum(int(np.array_equal(isl.current(i)[0], g_vec)) for i in islands)
    for i in islands:
        accepted += _replay_ts_step(isl, i, spec, op, oracle, exact)
    isl.step(7)
    for i in islands:                                   # full deques (500 ids), mid-run
        assert len(isl.trace_tabu(i, 0)[0]) == 500
        accepted += _replay_ts_step(isl, i, spec, op, oracle, exact)
    assert adopted >= 2 and accepted >= 6
    st = isl.stats()
    assert st["steps"] == 26 and st["candidates"] == 26 * 4096 * 2368
    v, s = isl.best(-1)
    assert sorted(v.tolist()) == list(range(1, 1000))
    want = oracle.score_round(op.score_incremental(v, [[]])[0], spec.score_precision)
    assert s[0] == want[0] and abs(s[1] - want[1]) <= (0.0 if exact else QUANTUM)
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
    # bench.py runs C5 with tree sums: the same trace in that mode (lean layout: edge lengths in HBM)
    gp.set_exact_sums(False)
    tt = TabuSearch(4096, 0.2, True, None, [0, 0.5, 0, 0, 0, 0.5], 10, scoring="delta").build_agent(gp, n_islands=4, seed=9)
    assert tt.step_path == "fused_lean"
    for _ in range(2):
        tbase, tcur = tt.current(1)
        ttr = tt.trace_step(1)
        pick = list(range(0, 4096, 64))
        twant = oracle.score_round(op.score_incremental(tbase, [ttr["deltas"][j] for j in pick]), spec.score_precision)
        assert np.array_equal(ttr["scores"][pick, 0], twant[:, 0])
        assert np.max(np.abs(ttr["scores"][pick, 1] - twant[:, 1])) <= 1.001e-3
        tsel, tacc = oracle.ts_select(ttr["scores"], tcur)
        assert (ttr["selected"], ttr["accepted"]) == (tsel, tacc)
    tt.close()
    gp.set_exact_sums(True)
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
