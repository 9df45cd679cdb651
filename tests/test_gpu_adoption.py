"""update_global_top's adopt half (agent_base.rs:446-490) for LateAcceptance / SimulatedAnnealing: the
reference re-tests `global.score < agent_top.score` after EVERY iteration and puts population[0] back
on the global top each time (LateAcceptance pushing the score it leaves behind), so an agent that
accepts a worse neighbour right after adopting is reset again until a step rejects or improves.  The
traced island is modelled on the CPU iteration by iteration (step -> update_top_individual ->
update_global_top) and compared with the device after every step, on the chain kernels (which make
the test inside a launch) and on the per-step kernels."""
import numpy as np
import pytest

from greyjack_b200 import LateAcceptance, Problem, SimulatedAnnealing, instances as inst

pytestmark = pytest.mark.gpu


def _publish(oracle, G, tops):
    """k_global_top: the best agent top (first index on ties) replaces the global top when strictly better."""
    best = 0
    for j in range(1, len(tops)):
        if oracle.score_cmp(tops[j][1], tops[best][1]) < 0:
            best = j
    if G is None or oracle.score_cmp(tops[best][1], G[1]) < 0:
        return (tops[best][0].copy(), tops[best][1].copy())
    return G


@pytest.mark.parametrize("agent", ["la", "sa"])
@pytest.mark.parametrize("scoring,mk", [("delta", lambda: inst.nqueens(48)), ("full", lambda: inst.tsp(60, seed=7))],
                         ids=["chain-nqueens", "perstep-tsp"])
def test_global_top_adoption_follows_the_reference_every_step(scoring, mk, agent, oracle):
    spec = mk()
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    n = spec.n_vars
    size = 4
    rng = np.random.default_rng(3)
    if spec.kind == inst.TSP:         # island 0: the greedy tour; the others: random tours
        good = spec.initial
        bad = [rng.permutation(n).astype(np.float64) + 1 for _ in range(2)]
    else:                             # island 0: a permutation; the others: every queen on one row
        good = spec.initial
        bad = [np.zeros(n), np.full(n, float(n - 1))]
    init = np.stack([good, bad[0], bad[1]])
    probas = [0.0, 1.0, 0.0, 0.0, 0.0, 0.0]
    if agent == "la":
        isl = LateAcceptance(size, 0.0, None, probas, 10 ** 6, scoring=scoring).build_agent(gp, n_islands=3, seed=17, initial=init)
    else:
        t0 = [2.0] if spec.levels == 1 else [2.0, 500.0]
        isl = SimulatedAnnealing(t0, 0.995, 0.0, None, probas, 10 ** 6, scoring=scoring).build_agent(gp, n_islands=3, seed=17, initial=init)
        temp = np.array(t0, dtype=np.float64)
    assert isl.step_path == ("chain" if scoring == "delta" else "full")
    T = 1                                            # the traced island starts far worse than island 0
    cur = list(isl.current(T)); top = [cur[0].copy(), cur[1].copy()]
    late = []
    G = None
    bounces = holds = 0

    def end_of_iteration(G):
        """update_global_top after the iteration: publish, then the adopt half for island T."""
        nonlocal cur, late
        isl.best(-1)                                  # device: publish (+ adopt on the per-step path)
        tops = [isl.best(j) for j in range(3)]
        assert np.array_equal(tops[T][1], top[1]) and np.array_equal(tops[T][0], top[0])
        G = _publish(oracle, G, tops)
        if oracle.score_cmp(G[1], top[1]) < 0:        # agent_base.rs:465-489
            if agent == "la":
                late = ([cur[1].copy()] + late)[:size]
            cur = [G[0].copy(), G[1].copy()]
        return G

    G = end_of_iteration(G)
    assert np.array_equal(cur[0], good)               # adopted island 0's start right away
    for step in range(70):
        dv, ds = isl.current(T)
        assert np.array_equal(dv, cur[0]) and np.array_equal(ds, cur[1]), step
        standing_on_g = np.array_equal(cur[0], G[0]) and oracle.score_cmp(G[1], top[1]) < 0
        holds += int(standing_on_g)
        tr = isl.trace_step(T)
        want = oracle.score_round(op.score_incremental(cur[0], tr["deltas"]), spec.score_precision)
        assert np.array_equal(tr["scores"], want)
        sc = tr["scores"][0]
        if agent == "la":
            acc, late = oracle.la_accept(sc, cur[1], late, size)
        else:
            aux = isl.trace_aux(T)
            acc, temp, _ = oracle.sa_accept(sc, cur[1], temp, 0.995, 1.0, aux["random"])
        assert tr["accepted"] == acc, step
        if acc:
            vec = cur[0].copy()
            for c, v in tr["deltas"][0]:
                vec[c] = v
            cur = [vec, sc.copy()]
            bounces += int(standing_on_g and oracle.score_cmp(sc, G[1]) > 0)
        if oracle.score_le(cur[1], top[1]):           # update_top_individual (agent_base.rs:220-224)
            top = [cur[0].copy(), cur[1].copy()]
        G = end_of_iteration(G)
    assert holds > 0
    if agent == "la":
        assert bounces > 0                            # a worse neighbour was accepted on the global top and undone
    isl.close(); gp.close()
