"""GPU parity of the device-resident agents: every generated move, every candidate score and
every selection decision of a TabuSearch / LateAcceptance step is replayed through the CPU
oracle (mover.rs, the ISC scorers, tabu_search_base.rs / late_acceptance_base.rs); GA
generations are checked through their invariants and the oracle's plain scorer."""
import numpy as np
import pytest

from greyjack_b200 import (GeneticAlgorithm, LateAcceptance, Problem, TabuSearch, instances as inst)

pytestmark = pytest.mark.gpu

MIX = [0.0, 0.2, 0.2, 0.2, 0.2, 0.2]      # examples/tsp/src/main.rs:47
ALL = [0.2, 0.16, 0.16, 0.16, 0.16, 0.16]


def _same_score(sc, unrounded, spec, oracle):
    """A stored score is either a rounded candidate score (agent_base.rs:311-314) or the
    never-rounded score of an initial individual (agent_base.rs:190-218)."""
    return np.array_equal(sc, unrounded) or np.array_equal(sc, oracle.score_round(unrounded, spec.score_precision))


def _oracle_move(op, spec, base, d, noop=True):
    """Replays one device move descriptor through the oracle mover (incremental form)."""
    kind, group, k = int(d[0]), int(d[1]), int(d[2])
    a, v = d[4:12], d[12:20]
    names = list(spec.groups.keys())
    g = np.asarray(spec.groups[names[group]], dtype=np.int32)
    if kind == 255:
        return []
    if kind == 0:
        res = op.move_change(base, g, a[:k], v[:k].astype(np.float64), True)
    elif kind == 1:
        res = op.move_swap(base, g, a[:k], True)
    elif kind == 2:
        res = op.move_swap_edges(base, g, a[:k], True)
    elif kind == 3:
        res = op.move_scramble(base, g, int(a[0]), v[:k], True)
    elif kind == 4:
        res = op.move_insertion(base, g, int(a[0]), int(a[1]), True)
    else:
        res = op.move_inverse(base, g, int(a[0]), int(a[1]), True)
    assert res is not None
    cols, vals = res
    vals = op.fix_deltas(cols, vals)
    return [(int(c), float(x)) for c, x in zip(cols, vals)]


def _final_state(n, pairs):
    out = {}
    for c, v in pairs:
        out[c] = v
    return out


def _check_rounded(got, want_unrounded, spec, oracle):
    """Default exact-sums mode: the rounded candidate scores are bit-identical to the oracle's
    (agent_base.rs:311-314 rounding included)."""
    want = oracle.score_round(want_unrounded, spec.score_precision)
    assert np.array_equal(got, want)


CASES = [
    ("nq64", lambda: inst.nqueens(64), [0.0, 1.0, 0.0, 0.0, 0.0, 0.0], 0.0, None),
    ("nq64-all", lambda: inst.nqueens(64), ALL, 0.2, 1.0),
    ("tsp200-mix", lambda: inst.tsp(200, seed=3), MIX, 0.5, None),
    ("tsp200-all", lambda: inst.tsp(200, seed=3), ALL, 0.0, 1.0),
    ("cvrp60", lambda: inst.cvrp(60, 6, seed=2), [0.5, 0.5, 0.0, 0.0, 0.0, 0.0], 0.8, None),
    ("cvrp60-all", lambda: inst.cvrp(60, 6, seed=2), ALL, 0.2, 1.0),
    ("vrpsvc80", lambda: inst.vrptw(80, 6, n_depots=2, seed=3), [0.5, 0.5, 0.0, 0.0, 0.0, 0.0], 0.2, None),
    ("vrptw80-all", lambda: inst.vrptw(80, 6, n_depots=2, seed=3, service_variant=False), ALL, 0.2, 1.0),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_tabu_search_step_replay(case, oracle):
    _, mk, probas, tabu, mult = case
    spec = mk()
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    K = 96
    isl = TabuSearch(K, tabu, True, mult, probas, 10).build_agent(gp, n_islands=2, seed=1234)
    for island in (0, 1):
        for _ in range(4):
            base, cur_score = isl.current(island)
            tr = isl.trace_step(island)
            # (1) every move expands to exactly what mover.rs would emit for the same choices
            for j in range(K):
                want = _oracle_move(op, spec, base, tr["desc"][j])
                assert _final_state(spec.n_vars, tr["deltas"][j]) == _final_state(spec.n_vars, want), (j, tr["desc"][j])
                assert [c for c, _ in tr["deltas"][j]] == [c for c, _ in want]
            # (2) every candidate score equals the ISC oracle on the same delta list
            _check_rounded(tr["scores"], op.score_incremental(base, tr["deltas"]), spec, oracle)
            # (3) selection: first minimum, accept iff best <= current
            sel, acc = oracle.ts_select(tr["scores"], cur_score)
            assert (tr["selected"], tr["accepted"]) == (sel, acc)
            # (4) the stored individual
            new, new_score = isl.current(island)
            want_vec = base.copy()
            if acc:
                for c, v in tr["deltas"][sel]:
                    want_vec[c] = v
                assert np.array_equal(new_score, tr["scores"][sel])
            else:
                assert np.array_equal(new_score, cur_score)
            assert np.array_equal(new, want_vec)
    isl.close(); gp.close()


def test_moves_cover_every_kind_and_are_distinct(oracle):
    spec = inst.tsp(300, seed=5)
    gp = Problem(spec)
    isl = TabuSearch(512, 0.0, True, 1.0, ALL, 10, reference_noop_moves=False).build_agent(gp, seed=7)
    tr = isl.trace_step(0)
    kinds = set(int(k) for k in tr["kinds"])
    assert kinds >= {0, 1, 2, 3, 4, 5}
    # chosen positions of a move are distinct (choice without replacement, math_utils.rs:43-45)
    for d in tr["desc"]:
        if d[0] in (0, 1, 2):
            assert len(set(d[4:4 + d[2]])) == d[2]
        if d[0] in (4, 5):
            assert d[4] != d[5]
    # without the reference's no-op quirk scramble really permutes
    base = spec.initial
    scr = [j for j, k in enumerate(tr["kinds"]) if k == 3]
    changed = sum(any(base[c] != v for c, v in tr["deltas"][j]) for j in scr)
    assert changed > len(scr) // 2
    isl.close(); gp.close()


def test_late_acceptance_chain_replay(oracle):
    spec = inst.tsp(120, seed=9)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    size = 5
    isl = LateAcceptance(size, 0.2, None, MIX, 10000).build_agent(gp, n_islands=3, seed=99)
    late = []
    for step in range(40):
        base, cur_score = isl.current(1)
        tr = isl.trace_step(1)
        want = _oracle_move(op, spec, base, tr["desc"][0])
        assert _final_state(spec.n_vars, tr["deltas"][0]) == _final_state(spec.n_vars, want)
        _check_rounded(tr["scores"], op.score_incremental(base, tr["deltas"]), spec, oracle)
        acc, late = oracle.la_accept(tr["scores"][0], cur_score, late, size)
        assert tr["accepted"] == acc, step
        new, new_score = isl.current(1)
        assert np.array_equal(new_score, tr["scores"][0] if acc else cur_score)
    isl.close(); gp.close()


@pytest.mark.parametrize("mk", [lambda: inst.tsp(150, seed=2), lambda: inst.nqueens(48),
                                lambda: inst.cvrp(50, 5, seed=6),
                                lambda: inst.vrptw(50, 5, n_depots=2, seed=6)],
                         ids=["tsp", "nqueens", "cvrp", "vrpsvc"])
def test_tabu_search_run_is_consistent(mk, oracle):
    spec = mk()
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    probas = [0.0, 1.0, 0.0, 0.0, 0.0, 0.0] if spec.kind == inst.NQUEENS else [0.1, 0.3, 0.1, 0.1, 0.2, 0.2]
    isl = TabuSearch(128, 0.2, True, None, probas, 5, reference_noop_moves=False).build_agent(gp, n_islands=4, seed=5)
    _, s0 = isl.best(0)
    prev = None
    for _ in range(6):
        isl.step(10)
        vec, sc = isl.best(-1)
        # the carried score is the (rounded) ISC score of the carried vector
        assert _same_score(sc, op.score_incremental(vec, [[]])[0], spec, oracle)
        if prev is not None:
            assert oracle.score_cmp(sc, prev) <= 0          # global best never gets worse
        prev = sc
        for i in range(4):
            cv, cs = isl.current(i)
            assert _same_score(cs, op.score_incremental(cv, [[]])[0], spec, oracle)
    assert oracle.score_cmp(prev, s0) < 0                   # and it did improve
    st = isl.stats()
    assert st["steps"] == 60 and st["candidates"] == 60 * 128 * 4
    isl.close(); gp.close()


def test_migration_and_global_best(oracle):
    spec = inst.tsp(100, seed=4)
    gp = Problem(spec)
    # island 0 starts from the greedy tour, the others from a bad (identity) tour
    init = np.stack([spec.initial] + [np.arange(1, 100, dtype=np.float64)] * 3)
    isl = TabuSearch(8, 0.0, False, None, [0, 1, 0, 0, 0, 0], 1).build_agent(gp, n_islands=4, seed=3, initial=init)
    scores0 = [isl.current(i)[1] for i in range(4)]
    assert oracle.score_cmp(scores0[0], scores0[1]) < 0
    isl.step(1)      # migration_frequency = 1: island 1 receives island 0's individual (<= rule)
    s1 = isl.current(1)[1]
    assert oracle.score_cmp(s1, scores0[1]) < 0
    # agent_base.rs:161-183: odd agents receive BEFORE they send, so island 1 forwards what it just took to
    # island 2; island 2 (even) had already sent its own individual to island 3
    assert oracle.score_cmp(isl.current(2)[1], s1) == 0
    assert oracle.score_cmp(s1, isl.current(3)[1]) < 0
    isl.step(3)
    gv, gs = isl.best(-1)
    for i in range(4):
        assert oracle.score_cmp(gs, isl.best(i)[1]) <= 0
    isl.close()
    # compare_to_global = true: every island adopts the global best right after it appears
    isl = TabuSearch(8, 0.0, True, None, [0, 1, 0, 0, 0, 0], 1000).build_agent(gp, n_islands=4, seed=3, initial=init)
    isl.step(1)
    gs = isl.best(-1)[1]
    for i in range(1, 4):
        assert oracle.score_cmp(isl.current(i)[1], scores0[1]) < 0
    isl.close(); gp.close()


@pytest.mark.parametrize("mk", [lambda: inst.cvrp(60, 6, seed=2, greedy=False), lambda: inst.tsp(80, seed=3, greedy=False),
                                lambda: inst.vrptw(40, 4, n_depots=2, seed=8, greedy=False)],
                         ids=["cvrp", "tsp", "vrpsvc"])
def test_genetic_algorithm_generations(mk, oracle):
    spec = mk()
    spec.initial = np.full(spec.n_vars, np.nan)     # None: population sampled uniformly (gj_integer.rs:98-112)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    ga = GeneticAlgorithm(256, 0.5, 0.2, 0.05, 1.0, None, 0.02, 3).build_agent(gp, n_islands=3, seed=11)
    prev = None
    for _ in range(5):
        ga.step(4)
        for i in (0, 2):
            vec, sc = ga.best(i)
            assert _same_score(sc, op.score_plain(vec)[0], spec, oracle)   # PSC semantics
        gv, gs = ga.best(-1)
        if prev is not None:
            assert oracle.score_cmp(gs, prev) <= 0
        prev = gs
        cv, cs = ga.current(1)
        assert _same_score(cs, op.score_plain(cv)[0], spec, oracle)
    assert ga.stats()["candidates"] == 20 * 256 * 3
    ga.close(); gp.close()


def test_hybrid_ring_moves_elites_between_tabu_and_ga_groups(oracle):
    """BASELINE config 5's mixed ring: a TabuSearch group and a GeneticAlgorithm group on one problem,
    chained by ring.HybridRing.  The GA group starts from random tours, the TabuSearch group from the
    greedy tour: after one exchange the GA population holds the tabu group's individual (Population
    rule: the migrant replaces the worst individual when it is <= it, agent_base.rs:405-412, 435-439);
    the tabu group's first island keeps its own (LocalSearch rule: migrant <= population[0], :429-434)."""
    torch = pytest.importorskip("torch")
    from greyjack_b200 import ring
    spec = inst.tsp(120, seed=6)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    ts = TabuSearch(256, 0.2, True, None, [0, 0.5, 0, 0, 0, 0.5], 10, scoring="delta").build_agent(gp, n_islands=3, seed=1)
    rnd = inst.tsp(120, seed=6, greedy=False)
    rnd.initial = np.full(spec.n_vars, np.nan)
    gp2 = Problem(rnd)
    ga = GeneticAlgorithm(64, 0.5, 0.2, 0.0, 1.0, [0, 0.5, 0, 0, 0, 0.5], 0.01, 10).build_agent(gp2, n_islands=2, seed=2)
    assert ts.migrant_bytes() == ga.migrant_bytes()          # ceil(0.01 * 64) == 1 individual per exchange
    hyb = ring.HybridRing([ts, ga], 0, 1)
    ts_cur = [ts.current(i) for i in range(3)]
    ga_best_before = ga.best(0)[1]
    assert oracle.score_cmp(ts_cur[2][1], ga_best_before) < 0
    hyb.exchange(0)
    torch.cuda.synchronize()
    rows, scores, order = ga.ga_population(0)
    # the last tabu island's individual now sits in GA island 0 (it beat that island's worst individual)
    hit = [k for k in range(64) if np.array_equal(rows[k], ts_cur[2][0])]
    assert len(hit) == 1 and np.array_equal(scores[hit[0]], ts_cur[2][1])
    assert _same_score(scores[hit[0]], op.score_plain(rows[hit[0]])[0], spec, oracle)
    # ... and is GA island 0's best individual after the re-sort
    assert order[0] == hit[0]
    # GA island 1 received GA island 0's best (in-group link), tabu island 0 was offered GA island 1's best
    # (a random tour, far worse than the greedy one): rejected
    assert np.array_equal(ts.current(0)[0], ts_cur[0][0])
    assert "TabuSearch x3 -> GeneticAlgorithm x2" in hyb.describe()
    ga.step(3); ts.step(3)
    hyb.exchange(0)
    assert oracle.score_cmp(ga.best(-1)[1], ga_best_before) < 0
    ga.close(); ts.close(); gp.close(); gp2.close()
