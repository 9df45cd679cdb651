"""GeneticAlgorithm islands replayed decision by decision through the oracle's restatement of
genetic_algorithm_base.rs: population.sort() (:157), select_p_best x2 (:83-92), cross (:105-134),
Mover::do_move(plain form) + fix_variables (:168-180), request_score_plain + round
(agent_base.rs:283-287), build_updated_population with select_p_worst (:198-213), and the mover's
tabu deque (mover.rs:75-96).  The device exposes its random draws (gj_islands_ga_trace_generation);
every decision that follows from them is recomputed on the CPU."""
import functools

import numpy as np
import pytest

from greyjack_b200 import GeneticAlgorithm, Problem, instances as inst

pytestmark = pytest.mark.gpu


def _plain_move(op, spec, cand, d):
    """Replays one device move descriptor through the oracle mover, PLAIN form -> (cols, candidate)."""
    kind, group, k = int(d[0]), int(d[1]), int(d[2])
    a, v = d[4:12], d[12:20]
    g = np.asarray(spec.groups[list(spec.groups.keys())[group]], dtype=np.int32)
    if kind == 255:
        return np.zeros(0, dtype=np.int32), np.array(cand, dtype=np.float64)
    if kind == 0:
        res = op.move_change(cand, g, a[:k], v[:k].astype(np.float64), False)
    elif kind == 1:
        res = op.move_swap(cand, g, a[:k], False)
    elif kind == 2:
        res = op.move_swap_edges(cand, g, a[:k], False)
    elif kind == 3:
        res = op.move_scramble(cand, g, int(a[0]), v[:k], False)
    elif kind == 4:
        res = op.move_insertion(cand, g, int(a[0]), int(a[1]), False)
    else:
        res = op.move_inverse(cand, g, int(a[0]), int(a[1]), False)
    assert res is not None
    cols, out = res
    return cols, op.fix_variables(out, cols)


def _selected(d):
    kind, k = int(d[0]), int(d[2])
    if kind == 255:
        return []
    if kind == 3:
        return [int(d[4])]
    return [int(x) for x in d[4:4 + (2 if kind >= 4 else k)]]


CASES = [
    ("cvrp60", lambda: inst.cvrp(60, 6, seed=2, greedy=False), None, 0.05),
    ("tsp80-2opt", lambda: inst.tsp(80, seed=3, greedy=False), [0, 0.5, 0, 0, 0, 0.5], 0.2),
    ("vrpsvc40-all", lambda: inst.vrptw(40, 4, n_depots=2, seed=8, greedy=False), [0.2, 0.16, 0.16, 0.16, 0.16, 0.16], 0.0),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_ga_generation_replay(case, oracle):
    _, mk, probas, tabu_rate = case
    spec = mk()
    spec.initial = np.full(spec.n_vars, np.nan)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    pop, cx, p_best = 192, 0.5, 0.2
    ga = GeneticAlgorithm(pop, cx, p_best, tabu_rate, 1.0, probas, 0.02, 1000).build_agent(gp, n_islands=2, seed=21)
    half = (pop + 1) // 2
    names = list(spec.groups.keys())
    crossed = swapped = taken = 0
    for gen in range(3):
        rows, scores, order = ga.ga_population(1)
        # population.sort(): stable, by Ord::cmp of the score
        idx = sorted(range(pop), key=functools.cmp_to_key(lambda x, y: oracle.score_cmp(scores[x], scores[y])))
        assert order.tolist() == idx
        tabu_before = [ga.trace_tabu(1, gi) for gi in range(len(names))]
        tr = ga.ga_trace_generation(1)
        assert np.array_equal(tr["order_before"], order)
        pushed = [[] for _ in names]
        for q in range(half):
            p1, lt1, id1, p2, lt2, id2, u, w = tr["pairs"][q]
            r1, l1 = oracle.ga_select(p1, id1, pop)
            r2, l2 = oracle.ga_select(p2, id2, pop)
            assert (l1, l2) == (lt1, lt2) and r1 >= 0 and r2 >= 0
            assert 0.000001 <= p1 < p_best and 0.000001 <= p2 < p_best
            c1, c2 = rows[order[r1]], rows[order[r2]]
            if u <= cx:                                   # genetic_algorithm_base.rs:164-166
                assert 0.0 <= w <= 1.0
                n1, n2 = oracle.ga_cross(c1, c2, w)
                crossed += 1
                swapped += int(np.array_equal(n1, c2) and not np.array_equal(c1, c2))
                c1, c2 = n1, n2
            else:
                assert w == -1.0
            for child, c in ((0, c1), (1, c2)):
                d = tr["desc"][2 * q + child]
                _, want = _plain_move(op, spec, c, d)
                assert np.array_equal(tr["cand_rows"][2 * q + child], want), (gen, q, child, d)
                sel = _selected(d)
                if tabu_rate and sel:
                    assert not (set(sel) & set(tabu_before[int(d[1])][0].tolist())), (gen, q, child)
                    pushed[int(d[1])] += sel
        # request_score_plain + ScoreTrait::round (agent_base.rs:283-287)
        want_sc = oracle.score_round(op.score_plain(tr["cand_rows"]), spec.score_precision)
        if spec.kind == inst.TSP:
            assert np.array_equal(tr["cand_scores"][:, 0], want_sc[:, 0])
            assert np.array_equal(tr["cand_scores"], want_sc)          # exact sums (default): bit-exact
        else:
            assert np.array_equal(tr["cand_scores"], want_sc)
        # build_updated_population (:198-213)
        worst = np.zeros(pop, dtype=np.int64)
        for i in range(pop):
            p, lt, idd = tr["replace"][i]
            rk, l = oracle.ga_select(p, idd, pop, worst=True)
            assert l == lt and rk >= 0
            assert tr["src"][i] == i or tr["src"][i] == -(rk + 1)
            worst[i] = order[rk]
        src = oracle.ga_replace(tr["cand_scores"][:pop], scores, worst)
        rows2, scores2, order2 = ga.ga_population(1)
        for i in range(pop):
            if src[i] >= 0:
                assert tr["src"][i] == i
                assert np.array_equal(rows2[i], tr["cand_rows"][i]) and np.array_equal(scores2[i], tr["cand_scores"][i])
                taken += 1
            else:
                assert tr["src"][i] < 0
                assert np.array_equal(rows2[i], rows[worst[i]]) and np.array_equal(scores2[i], scores[worst[i]])
        # the mover's tabu deque advanced by the generation's ids, newest first (declared batching:
        # every offspring saw the generation-start deque)
        if tabu_rate:
            for gi in range(len(names)):
                before, size = tabu_before[gi]
                after, _ = ga.trace_tabu(1, gi)
                want = (pushed[gi][::-1] + before.tolist())[:size]
                assert after.tolist() == want
    assert crossed > 0 and swapped > 0 and taken > 0
    ga.close(); gp.close()
