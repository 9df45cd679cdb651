"""GPU parity of DELTA scoring (GJ_SCORING_DELTA, csrc/gj_delta.cuh): islands that score a
neighbour from the O(k) constraint terms its move touches instead of re-scoring the whole vector.

Bar (north_star): integer score levels and integer move deltas bit-exact with the reference's
scorer (oracle ISC on the same delta lists); the float level (TSP tour length) within a stated
tolerance: one 10^-3 quantum of ScoreTrait::round after rounding (= 1e-12 relative before it;
the exact tour length sits ON the truncation boundary, see gj_eval.cuh).  The stored current /
best scores are re-scored by the full evaluator after every accepted move and must be bit-exact.
"""
import numpy as np
import pytest

from greyjack_b200 import LateAcceptance, Problem, SimulatedAnnealing, TabuSearch, instances as inst
from test_gpu_islands import ALL, MIX, _final_state, _oracle_move, _same_score

pytestmark = pytest.mark.gpu

QUANTUM = 1.0e-3 + 1e-8      # one ScoreTrait::round quantum (+ f64 noise of subtracting two ~1e4 scores)
# "delta": the fused single-kernel step (gj_islands_fused.cuh); "delta_unfused": the same
# arithmetic as separate kernels (what DELTA uses when an island does not fit in shared memory)
# "delta_f64": the fused step in f64 where "delta" would pick the fixed-point TSP step (gj_islands_tsfast.cuh)
SCORINGS = ["delta", "delta_f64", "delta_unfused"]


def _check_delta_scores(got, want_unrounded, spec, oracle):
    want = oracle.score_round(want_unrounded, spec.score_precision)
    L = spec.levels
    if L == 1:
        assert np.array_equal(got, want)
        return
    for l in range(L - 1):
        assert np.array_equal(got[:, l], want[:, l]), f"integer level {l} differs"
    assert np.max(np.abs(got[:, L - 1] - want[:, L - 1])) <= QUANTUM
    # and before rounding the two sums agree to 1e-12 relative: the rounded values can only
    # differ when the unrounded oracle value is within 1e-9 of a quantum boundary
    diff = got[:, L - 1] != want[:, L - 1]
    frac = np.abs(want_unrounded[diff, L - 1] * 1000.0 - np.rint(want_unrounded[diff, L - 1] * 1000.0))
    assert np.all(frac < 1e-6)


def _tsp_fractional_hard_weight():
    # a fractional weight on the hard level with precision 0 there: 1..3 duplicates all round to hard 0 and
    # the decision falls to the soft level (the in-order shortcut of the fused step must not drop them)
    spec = inst.tsp(120, seed=6)
    spec.weights = np.array([0.3, 1.0, 1.0, 1.0])
    spec.score_precision = [0, 3]
    return spec


CASES = [
    ("nq64-swap", lambda: inst.nqueens(64), [0.0, 1.0, 0.0, 0.0, 0.0, 0.0], 0.0, None),
    ("nq64-all", lambda: inst.nqueens(64), ALL, 0.2, 1.0),
    ("nq97-small", lambda: inst.nqueens(97), [0.3, 0.3, 0.2, 0.2, 0.0, 0.0], 0.1, 2.0),
    ("tsp200-mix", lambda: inst.tsp(200, seed=3), MIX, 0.5, None),
    ("tsp200-all", lambda: inst.tsp(200, seed=3), ALL, 0.0, 1.0),
    ("tsp131-all-mult", lambda: inst.tsp(131, seed=8), ALL, 0.3, 3.0),
    ("tsp64-2opt", lambda: inst.tsp(64, seed=5), [0.0, 0.5, 0.0, 0.0, 0.0, 0.5], 0.5, None),
    ("tsp120-fractional-hard", _tsp_fractional_hard_weight, [0.5, 0.25, 0.0, 0.0, 0.0, 0.25], 0.0, None),
]


@pytest.mark.parametrize("scoring", SCORINGS)
@pytest.mark.parametrize("noop", [True, False], ids=["refquirk", "plainform"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_delta_step_replay(case, noop, scoring, oracle):
    _, mk, probas, tabu, mult = case
    spec = mk()
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    K = 160
    isl = TabuSearch(K, tabu, True, mult, probas, 10, reference_noop_moves=noop,
                     scoring=scoring).build_agent(gp, n_islands=2, seed=4321)
    for island in (0, 1):
        for _ in range(5):
            base, cur_score = isl.current(island)
            assert _same_score(cur_score, op.score_incremental(base, [[]])[0], spec, oracle)
            tr = isl.trace_step(island)
            if noop:
                for j in range(K):
                    want = _oracle_move(op, spec, base, tr["desc"][j])
                    assert _final_state(spec.n_vars, tr["deltas"][j]) == _final_state(spec.n_vars, want)
            _check_delta_scores(tr["scores"], op.score_incremental(base, tr["deltas"]), spec, oracle)
            sel, acc = oracle.ts_select(tr["scores"], cur_score)
            assert (tr["selected"], tr["accepted"]) == (sel, acc)
            new, new_score = isl.current(island)
            want_vec = base.copy()
            if acc:
                for c, v in tr["deltas"][sel]:
                    want_vec[c] = v
            assert np.array_equal(new, want_vec)
            # after an accepted move the stored score is the FULL evaluation of the new vector
            want_score = oracle.score_round(op.score_incremental(new, [[]])[0], spec.score_precision)
            if acc:
                assert np.array_equal(new_score, want_score)
            else:
                assert np.array_equal(new_score, cur_score)
    isl.close(); gp.close()


@pytest.mark.parametrize("scoring", SCORINGS)
@pytest.mark.parametrize("mk", [lambda: inst.tsp(150, seed=2), lambda: inst.nqueens(48)], ids=["tsp", "nqueens"])
def test_delta_and_full_islands_see_the_same_neighbourhood(mk, scoring, oracle):
    """Same seed -> same moves (the generator is a pure function of seed/island/step/candidate);
    integer levels identical, float level within one quantum."""
    spec = mk()
    gp = Problem(spec)
    probas = ALL
    a = TabuSearch(256, 0.2, True, 1.0, probas, 10, scoring="full").build_agent(gp, n_islands=2, seed=77)
    b = TabuSearch(256, 0.2, True, 1.0, probas, 10, scoring=scoring).build_agent(gp, n_islands=2, seed=77)
    ta, tb = a.trace_step(1), b.trace_step(1)
    assert np.array_equal(ta["desc"], tb["desc"])
    L = spec.levels
    for l in range(max(1, L - 1)):
        assert np.array_equal(ta["scores"][:, l], tb["scores"][:, l])
    if L > 1:
        assert np.max(np.abs(ta["scores"][:, L - 1] - tb["scores"][:, L - 1])) <= QUANTUM
    a.close(); b.close(); gp.close()


@pytest.mark.parametrize("mk", [lambda: inst.tsp(150, seed=2), lambda: inst.nqueens(48)], ids=["tsp", "nqueens"])
@pytest.mark.parametrize("scoring", SCORINGS)
@pytest.mark.parametrize("agent", ["ts", "la"])
def test_delta_run_is_consistent(mk, agent, scoring, oracle):
    spec = mk()
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    probas = [0.0, 1.0, 0.0, 0.0, 0.0, 0.0] if spec.kind == inst.NQUEENS else [0.1, 0.3, 0.1, 0.1, 0.2, 0.2]
    if agent == "ts":
        isl = TabuSearch(128, 0.2, True, None, probas, 5, reference_noop_moves=False,
                         scoring=scoring).build_agent(gp, n_islands=4, seed=5)
        per_step = 128
    else:
        isl = LateAcceptance(8, 0.2, None, probas, 5, reference_noop_moves=False,
                             scoring=scoring).build_agent(gp, n_islands=4, seed=5)
        per_step = 1
    _, s0 = isl.best(0)
    prev = None
    for _ in range(6):
        isl.step(25)
        vec, sc = isl.best(-1)
        assert _same_score(sc, op.score_incremental(vec, [[]])[0], spec, oracle)
        if prev is not None:
            assert oracle.score_cmp(sc, prev) <= 0
        prev = sc
        for i in range(4):
            cv, cs = isl.current(i)
            assert _same_score(cs, op.score_incremental(cv, [[]])[0], spec, oracle)
            bv, bs = isl.best(i)
            assert _same_score(bs, op.score_incremental(bv, [[]])[0], spec, oracle)
    assert oracle.score_cmp(prev, s0) < 0
    st = isl.stats()
    assert st["steps"] == 150 and st["candidates"] == 150 * per_step * 4
    isl.close(); gp.close()


@pytest.mark.parametrize("scoring", SCORINGS)
def test_delta_migration_keeps_state_in_sync(scoring, oracle):
    spec = inst.tsp(100, seed=4)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    init = np.stack([spec.initial] + [np.arange(1, 100, dtype=np.float64)] * 3)
    isl = TabuSearch(64, 0.0, True, None, [0, 0.5, 0, 0, 0, 0.5], 1, scoring=scoring).build_agent(
        gp, n_islands=4, seed=3, initial=init)
    for _ in range(8):
        isl.step(1)
        for i in range(4):
            base, cur_score = isl.current(i)
            assert _same_score(cur_score, op.score_incremental(base, [[]])[0], spec, oracle)
        # the cached state must describe the (possibly migrated / adopted) vector: replay a step
        base, _ = isl.current(2)
        tr = isl.trace_step(2)
        _check_delta_scores(tr["scores"], op.score_incremental(base, tr["deltas"]), spec, oracle)
    isl.close(); gp.close()


@pytest.mark.parametrize("scoring", SCORINGS)
def test_delta_change_moves_track_duplicates(scoring, oracle):
    """change_move introduces / removes duplicate stops: the hard level must follow exactly."""
    spec = inst.tsp(60, seed=11)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    isl = TabuSearch(512, 0.0, True, 4.0, [1.0, 0, 0, 0, 0, 0], 10, scoring=scoring).build_agent(gp, seed=9)
    seen_hard = set()
    for _ in range(4):
        base, _ = isl.current(0)
        tr = isl.trace_step(0)
        want = op.score_incremental(base, tr["deltas"])
        _check_delta_scores(tr["scores"], want, spec, oracle)
        seen_hard |= set(want[:, 0].tolist())
    assert len(seen_hard) > 2
    isl.close(); gp.close()


# ---- LateAcceptance chains (gj_islands_chain.cuh): many steps per launch, one warp per chain --------
LA_CASES = [
    ("tsp120-mix", lambda: inst.tsp(120, seed=9), MIX, 0.2, None),
    ("tsp97-all", lambda: inst.tsp(97, seed=2), ALL, 0.3, 2.0),
    ("tsp64-notabu", lambda: inst.tsp(64, seed=5), [0.0, 0.5, 0.0, 0.0, 0.0, 0.5], 0.0, None),
    ("nq64-swap", lambda: inst.nqueens(64), [0.0, 1.0, 0.0, 0.0, 0.0, 0.0], 0.2, None),
    ("nq48-all", lambda: inst.nqueens(48), ALL, 0.2, 1.0),
]


@pytest.mark.parametrize("case", LA_CASES, ids=lambda c: c[0])
def test_la_chain_step_replay(case, oracle):
    """Every step of a chain replayed through the oracle: move expansion (mover.rs), candidate score
    (ISC scorer), acceptance rule with the late list (late_acceptance_base.rs:188-241), tabu ids."""
    _, mk, probas, tabu, mult = case
    spec = mk()
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    size = 5
    isl = LateAcceptance(size, tabu, mult, probas, 10000, scoring="delta").build_agent(gp, n_islands=3, seed=99)
    late = []
    recent = []
    T = max(1, int(np.ceil(tabu * spec.n_vars))) if tabu else 0
    for step in range(60):
        base, cur_score = isl.current(1)
        tr = isl.trace_step(1)
        d = tr["desc"][0]
        want = _oracle_move(op, spec, base, d)
        assert _final_state(spec.n_vars, tr["deltas"][0]) == _final_state(spec.n_vars, want)
        _check_delta_scores(tr["scores"], op.score_incremental(base, tr["deltas"]), spec, oracle)
        acc, late = oracle.la_accept(tr["scores"][0], cur_score, late, size)
        assert tr["accepted"] == acc, step
        new, new_score = isl.current(1)
        want_vec = base.copy()
        if acc:
            for c, v in tr["deltas"][0]:
                want_vec[c] = v
        assert np.array_equal(new, want_vec)
        # stored score: the full evaluation of the stored vector (bit-exact), or untouched
        if acc:
            assert np.array_equal(new_score, oracle.score_round(op.score_incremental(new, [[]])[0], spec.score_precision))
            late[0] = tr["scores"][0]                    # the late list keeps the step's own score
        else:
            assert np.array_equal(new_score, cur_score)
        # tabu: a move never selects an id that one of the last T selections holds
        if tabu and d[0] != 255:
            k = 1 if d[0] == 3 else (2 if d[0] >= 4 else d[2])
            sel = [int(x) for x in d[4:4 + k]]
            assert not (set(sel) & set(recent[-T:])), (step, sel, recent[-T:])
            recent += sel
    isl.close(); gp.close()


@pytest.mark.parametrize("mk", [lambda: inst.tsp(150, seed=2), lambda: inst.nqueens(48)], ids=["tsp", "nqueens"])
def test_la_chain_many_steps_per_launch_equal_single_steps(mk, oracle):
    """The state carried in shared memory across the steps of one launch is the state single-step
    launches rebuild from HBM: same seed -> same chain, whatever the launch granularity."""
    spec = mk()
    op = oracle.OracleProblem(spec)
    probas = [0.0, 1.0, 0.0, 0.0, 0.0, 0.0] if spec.kind == inst.NQUEENS else [0.1, 0.3, 0.1, 0.1, 0.2, 0.2]
    gp = Problem(spec)
    a = LateAcceptance(8, 0.2, None, probas, 1000, scoring="delta").build_agent(gp, n_islands=1, seed=5)
    b = LateAcceptance(8, 0.2, None, probas, 1000, scoring="delta").build_agent(gp, n_islands=1, seed=5)
    a.step(64)
    for _ in range(64):
        b.trace_step(0)
    va, sa = a.current(0)
    vb, sb = b.current(0)
    assert np.array_equal(va, vb)
    assert np.array_equal(sa, sb)
    ba, bsa = a.best(0)
    bb, bsb = b.best(0)
    assert np.array_equal(ba, bb) and np.array_equal(bsa, bsb)
    assert _same_score(bsa, op.score_incremental(ba, [[]])[0], spec, oracle)
    assert a.stats()["candidates"] == 64 and a.stats()["accepted"] == b.stats()["accepted"]
    a.close(); b.close(); gp.close()


@pytest.mark.parametrize("mk", [lambda: inst.tsp(60, seed=2, greedy=False), lambda: inst.nqueens(64)], ids=["tsp", "nqueens"])
def test_la_chain_wide_and_narrow_launch_shapes_run_the_same_chains(mk, oracle):
    """More than 20 chains per SM switch k_la_chains to its wide shape (one CTA per SM, <= 72 registers,
    every chain of the SM resident -- the C1 bench shape).  Within one launch chains are independent, and a
    chain's random stream is keyed by its island id: the first chains of a wide group must end exactly where
    the same chains of a small (narrow-shape) group end."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    spec = mk()
    op = oracle.OracleProblem(spec)
    probas = [0.0, 1.0, 0.0, 0.0, 0.0, 0.0] if spec.kind == inst.NQUEENS else [0.1, 0.3, 0.1, 0.1, 0.2, 0.2]
    gp = Problem(spec)
    n_wide = 24 * sms + 5
    mkb = lambda n: LateAcceptance(8, 0.2, None, probas, 1000, scoring="delta", chain_steps_per_launch=48).build_agent(
        gp, n_islands=n, seed=5)
    a, b = mkb(n_wide), mkb(4)
    assert a.step_path == "chain" and b.step_path == "chain"
    start = [b.best(i)[1].copy() for i in range(4)]
    a.step(48); b.step(48)
    # (the agent tops are compared: reading a current solution first lands the pending adoption of the
    # group's global top, which the 3 500 other chains of the wide group have improved)
    for i in range(4):
        ba, bsa = a.best(i)
        bb, bsb = b.best(i)
        assert np.array_equal(ba, bb) and np.array_equal(bsa, bsb)
    assert a.stats()["accepted"] >= b.stats()["accepted"] > 0
    assert any(oracle.score_cmp(b.best(i)[1], start[i]) < 0 for i in range(4))      # the chains did move
    for i in (n_wide // 2, n_wide - 1):
        v, sc = a.current(i)
        assert _same_score(sc, op.score_incremental(v, [[]])[0], spec, oracle)
        v, sc = a.best(i)
        assert _same_score(sc, op.score_incremental(v, [[]])[0], spec, oracle)
    assert a.stats()["candidates"] == 48 * n_wide
    a.step(96)                                            # adoption of the global top + migration on the wide shape
    v, sc = a.best(-1)
    assert _same_score(sc, op.score_incremental(v, [[]])[0], spec, oracle)
    a.close(); b.close(); gp.close()


# ---- SimulatedAnnealing on the same chain kernel ---------------------------
@pytest.mark.parametrize("cooling", [0.98, None], ids=["cooling", "accomplish-rate"])
@pytest.mark.parametrize("mk", [lambda: inst.tsp(90, seed=6), lambda: inst.nqueens(40)], ids=["tsp", "nqueens"])
def test_simulated_annealing_step_replay(mk, cooling, oracle):
    """Move, score and the Metropolis rule of simulated_annealing_base.rs:198-233 replayed through the
    oracle with the step's own uniform value; temperatures must follow the schedule bit for bit."""
    from greyjack_b200 import SimulatedAnnealing
    spec = mk()
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    t0 = [0.5] if spec.levels == 1 else [0.5, 3000.0]
    isl = SimulatedAnnealing(t0, cooling, 0.2, None, ALL, 10000, scoring="delta").build_agent(gp, n_islands=2, seed=31)
    temp = np.array(t0, dtype=np.float64)
    accepted_worse = 0
    metropolis_steps = 0
    for step in range(80):
        rate = step / 100.0
        if cooling is None:
            isl.set_accomplish_rate(rate)
        base, cur_score = isl.current(1)
        tr = isl.trace_step(1)
        aux = isl.trace_aux(1)
        want = _oracle_move(op, spec, base, tr["desc"][0])
        assert _final_state(spec.n_vars, tr["deltas"][0]) == _final_state(spec.n_vars, want)
        _check_delta_scores(tr["scores"], op.score_incremental(base, tr["deltas"]), spec, oracle)
        acc, temp, proba = oracle.sa_accept(tr["scores"][0], cur_score, temp, cooling, 1.0 - rate, aux["random"])
        assert 0.0 <= aux["random"] < 1.0
        assert np.array_equal(aux["temperature"][: spec.levels], temp)
        assert aux["accept_proba"] == pytest.approx(proba, rel=1e-12)
        assert tr["accepted"] == acc, step
        accepted_worse += int(acc and oracle.score_cmp(tr["scores"][0], cur_score) > 0)
        metropolis_steps += int(proba < 1.0)
        new, new_score = isl.current(1)
        want_vec = base.copy()
        if acc:
            for c, v in tr["deltas"][0]:
                want_vec[c] = v
        assert np.array_equal(new, want_vec)
    assert accepted_worse >= 0 and metropolis_steps > 0     # worsening neighbours met the exp() branch
    isl.close(); gp.close()


def test_simulated_annealing_run_and_full_mode(oracle):
    from greyjack_b200 import SimulatedAnnealing
    spec = inst.tsp(150, seed=2)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    for scoring in ("delta", "full"):
        isl = SimulatedAnnealing([1.0, 5.0], 0.999, 0.2, None, [0.1, 0.3, 0.1, 0.1, 0.2, 0.2], 5,
                                 reference_noop_moves=False, scoring=scoring).build_agent(gp, n_islands=4, seed=5)
        _, s0 = isl.best(0)
        for _ in range(4):
            isl.step(100)
            vec, sc = isl.best(-1)
            assert _same_score(sc, op.score_incremental(vec, [[]])[0], spec, oracle)
            for i in range(4):
                cv, cs = isl.current(i)
                assert _same_score(cs, op.score_incremental(cv, [[]])[0], spec, oracle)
        assert oracle.score_cmp(isl.best(-1)[1], s0) < 0
        assert isl.stats()["steps"] == 400 and isl.stats()["candidates"] == 400 * 4
        isl.close()
    gp.close()


# ---- several semantic groups, non-consecutive groups, frozen variables ---------------------------------
def _grouped_specs():
    t = inst.tsp(90, seed=12)
    n = t.n_vars
    t.groups = {"front": np.arange(0, n // 2, dtype=np.int32), "back": np.arange(n // 2, n, dtype=np.int32)}
    t.name = "tsp90-two-groups"
    u = inst.tsp(70, seed=13)
    u.groups = {"even": np.arange(0, u.n_vars, 2, dtype=np.int32), "odd": np.arange(1, u.n_vars, 2, dtype=np.int32)}
    u.name = "tsp70-strided-groups"          # segment moves are NOT runs of consecutive stops: full evaluator
    f = inst.tsp(60, seed=14)
    f.frozen = np.zeros(f.n_vars, dtype=np.uint8)
    f.frozen[[3, 4, 17, 40]] = 1              # frozen stops drop out of the group (variables_manager.rs:76-106)
    f.name = "tsp60-frozen"
    q = inst.nqueens(50)
    q.groups = {"left": np.arange(0, 20, dtype=np.int32), "right": np.arange(20, 50, dtype=np.int32)}
    q.name = "nq50-two-groups"
    return [t, u, f, q]


@pytest.mark.parametrize("scoring", SCORINGS + ["full"])
@pytest.mark.parametrize("spec", _grouped_specs(), ids=lambda s: s.name)
def test_groups_and_frozen_variables_step_replay(spec, scoring, oracle):
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    if spec.frozen is not None:
        # the mover only ever sees the unfrozen members of a group (variables_manager.rs:76-106)
        import copy
        spec = copy.copy(spec)
        spec.groups = {k: np.asarray([v for v in ids if not spec.frozen[v]], dtype=np.int32)
                       for k, ids in spec.groups.items()}
    K = 192
    isl = TabuSearch(K, 0.3, True, 1.5, ALL, 10, scoring=scoring).build_agent(gp, n_islands=2, seed=77)
    for _ in range(4):
        base, cur_score = isl.current(1)
        tr = isl.trace_step(1)
        groups_seen = set()
        for j in range(K):
            d = tr["desc"][j]
            want = _oracle_move(op, spec, base, d)
            assert _final_state(spec.n_vars, tr["deltas"][j]) == _final_state(spec.n_vars, want)
            groups_seen.add(int(d[1]))
            if spec.frozen is not None:
                assert not any(spec.frozen[c] for c, _ in tr["deltas"][j])      # frozen columns never move
        assert groups_seen == set(range(len(spec.groups)))
        want = op.score_incremental(base, tr["deltas"])
        if scoring == "full":
            assert np.array_equal(tr["scores"], oracle.score_round(want, spec.score_precision))
        else:
            _check_delta_scores(tr["scores"], want, spec, oracle)
        sel, acc = oracle.ts_select(tr["scores"], cur_score)
        assert (tr["selected"], tr["accepted"]) == (sel, acc)
    isl.close(); gp.close()


def test_tiny_instances_all_paths(oracle):
    """n = 6..9 variables: every buffer-size corner (padding, scratch, tabu tables) in all scoring modes."""
    for spec in (inst.nqueens(6), inst.tsp(8, seed=3), inst.nqueens(9)):
        op = oracle.OracleProblem(spec)
        gp = Problem(spec)
        for scoring in ("full", "delta", "delta_unfused"):
            ts = TabuSearch(40, 0.3, True, 1.0, ALL, 2, scoring=scoring).build_agent(gp, n_islands=3, seed=1)
            base, _ = ts.current(0)
            tr = ts.trace_step(0)
            _check_delta_scores(tr["scores"], op.score_incremental(base, tr["deltas"]), spec, oracle)
            ts.step(10)
            v, s = ts.best(-1)
            assert _same_score(s, op.score_incremental(v, [[]])[0], spec, oracle)
            ts.close()
            la = LateAcceptance(3, 0.3, None, ALL, 2, scoring=scoring).build_agent(gp, n_islands=3, seed=2)
            for _ in range(12):
                base, _ = la.current(1)
                tr = la.trace_step(1)
                _check_delta_scores(tr["scores"], op.score_incremental(base, tr["deltas"]), spec, oracle)
            la.step(20)
            v, s = la.best(-1)
            assert _same_score(s, op.score_incremental(v, [[]])[0], spec, oracle)
            la.close()
        gp.close()


# ---- VRP models: route-level delta evaluation (gj_vrp_delta.cuh) -- every level BIT-exact ---------------
VRP_CASES = [
    ("cvrp60", lambda: inst.cvrp(60, 6, seed=2), [0.5, 0.5, 0.0, 0.0, 0.0, 0.0], 0.8, None),
    ("cvrp60-all", lambda: inst.cvrp(60, 6, seed=2), ALL, 0.2, 1.0),
    ("cvrp45-mult", lambda: inst.cvrp(45, 5, seed=7), [0.3, 0.3, 0.2, 0.2, 0.0, 0.0], 0.1, 3.0),
    ("vrpsvc80", lambda: inst.vrptw(80, 6, n_depots=2, seed=3), [0.5, 0.5, 0.0, 0.0, 0.0, 0.0], 0.2, None),
    ("vrptw80-all", lambda: inst.vrptw(80, 6, n_depots=2, seed=3, service_variant=False), ALL, 0.2, 1.0),
    ("vrpsvc50-mult", lambda: inst.vrptw(50, 7, n_depots=3, seed=9), [0.3, 0.3, 0.2, 0.2, 0.0, 0.0], 0.0, 4.0),
]


@pytest.mark.parametrize("noop", [True, False], ids=["refquirk", "plainform"])
@pytest.mark.parametrize("case", VRP_CASES, ids=lambda c: c[0])
def test_vrp_delta_step_replay(case, noop, oracle):
    _, mk, probas, tabu, mult = case
    spec = mk()
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    K = 160
    isl = TabuSearch(K, tabu, True, mult, probas, 10, reference_noop_moves=noop,
                     scoring="delta").build_agent(gp, n_islands=2, seed=4321)
    for island in (0, 1):
        for _ in range(5):
            base, cur_score = isl.current(island)
            assert _same_score(cur_score, op.score_incremental(base, [[]])[0], spec, oracle)
            tr = isl.trace_step(island)
            if noop:
                for j in range(K):
                    want = _oracle_move(op, spec, base, tr["desc"][j])
                    assert _final_state(spec.n_vars, tr["deltas"][j]) == _final_state(spec.n_vars, want)
            want = oracle.score_round(op.score_incremental(base, tr["deltas"]), spec.score_precision)
            assert np.array_equal(tr["scores"], want)          # hard, medium AND the float soft level
            sel, acc = oracle.ts_select(tr["scores"], cur_score)
            assert (tr["selected"], tr["accepted"]) == (sel, acc)
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


@pytest.mark.parametrize("mk", [lambda: inst.cvrp(50, 5, seed=6), lambda: inst.vrptw(50, 5, n_depots=2, seed=6)],
                         ids=["cvrp", "vrpsvc"])
@pytest.mark.parametrize("agent", ["ts", "la"])
def test_vrp_delta_run_is_consistent(mk, agent, oracle):
    spec = mk()
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    probas = [0.3, 0.3, 0.1, 0.1, 0.1, 0.1]
    if agent == "ts":
        isl = TabuSearch(128, 0.2, True, None, probas, 5, reference_noop_moves=False, scoring="delta").build_agent(gp, n_islands=4, seed=5)
        per_step = 128
    else:
        isl = LateAcceptance(8, 0.2, None, probas, 5, reference_noop_moves=False, scoring="delta").build_agent(gp, n_islands=4, seed=5)
        per_step = 1
    _, s0 = isl.best(0)
    prev = None
    for _ in range(5):
        isl.step(20)
        vec, sc = isl.best(-1)
        assert _same_score(sc, op.score_incremental(vec, [[]])[0], spec, oracle)
        if prev is not None:
            assert oracle.score_cmp(sc, prev) <= 0
        prev = sc
        for i in range(4):
            cv, cs = isl.current(i)
            assert _same_score(cs, op.score_incremental(cv, [[]])[0], spec, oracle)
    assert oracle.score_cmp(prev, s0) < 0
    assert isl.stats()["candidates"] == 100 * per_step * 4
    isl.close(); gp.close()


# ---- VRP models, single-neighbour agents: chains over a route index (gj_islands_vrp_chain.cuh) ----------
VRP_CHAIN_CASES = [
    ("cvrp60", lambda: inst.cvrp(60, 6, seed=2), [0.5, 0.5, 0.0, 0.0, 0.0, 0.0], 0.3, None),
    ("cvrp45-small-moves", lambda: inst.cvrp(45, 5, seed=7), [0.3, 0.3, 0.2, 0.2, 0.0, 0.0], 0.1, 3.0),
    ("vrpsvc80", lambda: inst.vrptw(80, 6, n_depots=2, seed=3), [0.5, 0.5, 0.0, 0.0, 0.0, 0.0], 0.2, None),
    ("vrptw80-file", lambda: inst.vrptw(80, 6, n_depots=2, seed=3, service_variant=False), [0.4, 0.4, 0.1, 0.1, 0.0, 0.0], 0.2, 2.0),
    ("vrpsvc70-long-routes", lambda: inst.vrptw(70, 2, n_depots=1, seed=9), [0.5, 0.5, 0.0, 0.0, 0.0, 0.0], 0.0, 4.0),
]


@pytest.mark.parametrize("noop", [True, False], ids=["refquirk", "plainform"])
@pytest.mark.parametrize("case", VRP_CHAIN_CASES, ids=lambda c: c[0])
def test_vrp_chain_step_replay(case, noop, oracle):
    """Every step of a LateAcceptance chain on a VRP model replayed through the oracle: the move, the
    candidate score (all three levels bit-exact: the route re-walk keeps the reference's summation
    order), the acceptance rule, the stored vector and score."""
    _, mk, probas, tabu, mult = case
    spec = mk()
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    size = 4
    isl = LateAcceptance(size, tabu, mult, probas, 10000, reference_noop_moves=noop,
                         scoring="delta").build_agent(gp, n_islands=3, seed=77)
    assert isl.step_path == "vrp_chain"
    late = []
    for step in range(70):
        base, cur_score = isl.current(2)
        tr = isl.trace_step(2)
        if noop:
            want = _oracle_move(op, spec, base, tr["desc"][0])
            assert _final_state(spec.n_vars, tr["deltas"][0]) == _final_state(spec.n_vars, want)
        want_sc = oracle.score_round(op.score_incremental(base, tr["deltas"]), spec.score_precision)
        assert np.array_equal(tr["scores"], want_sc), step
        acc, late = oracle.la_accept(tr["scores"][0], cur_score, late, size)
        assert tr["accepted"] == acc, step
        new, new_score = isl.current(2)
        want_vec = base.copy()
        if acc:
            for c, v in tr["deltas"][0]:
                want_vec[c] = v
            assert np.array_equal(new_score, tr["scores"][0])
        else:
            assert np.array_equal(new_score, cur_score)
        assert np.array_equal(new, want_vec)
    vec, sc = isl.best(2)
    assert np.array_equal(sc, oracle.score_round(op.score_incremental(vec, [[]])[0], spec.score_precision))
    isl.close(); gp.close()


@pytest.mark.parametrize("agent", ["la", "sa"])
@pytest.mark.parametrize("mk", [lambda: inst.cvrp(90, 7, seed=6), lambda: inst.vrptw(90, 7, n_depots=2, seed=6)],
                         ids=["cvrp", "vrpsvc"])
def test_vrp_chain_many_steps_per_launch_equal_single_steps(mk, agent, oracle):
    """The route index carried across steps and launches == the one rebuilt from the solution row:
    same seed -> same chain whatever the launch granularity, and every stored score is the full
    evaluation of its vector."""
    spec = mk()
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    probas = [0.4, 0.4, 0.1, 0.1, 0.0, 0.0]

    def build():
        if agent == "la":
            b = LateAcceptance(8, 0.2, None, probas, 100000, reference_noop_moves=False, scoring="delta")
        else:
            b = SimulatedAnnealing([1.0, 1.0, 50.0], 0.999, 0.2, None, probas, 100000,
                                   reference_noop_moves=False, scoring="delta")
        return b.build_agent(gp, n_islands=1, seed=11)        # one chain: no global-top adoption

    a, b = build(), build()
    assert a.step_path == "vrp_chain"
    a.step(200)
    for _ in range(200):
        b.trace_step(0)
    for i in (0,):
        va, sa_ = a.current(i)
        vb, sb = b.current(i)
        assert np.array_equal(va, vb) and np.array_equal(sa_, sb)
        assert np.array_equal(sa_, oracle.score_round(op.score_incremental(va, [[]])[0], spec.score_precision))
        ba, bsa = a.best(i)
        bb, bsb = b.best(i)
        assert np.array_equal(ba, bb) and np.array_equal(bsa, bsb)
        assert np.array_equal(bsa, oracle.score_round(op.score_incremental(ba, [[]])[0], spec.score_precision))
    assert a.stats()["candidates"] == 200 and a.stats()["accepted"] == b.stats()["accepted"] > 0
    a.close(); b.close(); gp.close()


def test_vrp_chain_migration_and_global_top_keep_the_index_in_sync(oracle):
    """Migrants and adopted global tops replace a chain's solution between launches: the route index
    is rebuilt (stale flag) and every stored score stays the full evaluation of its vector."""
    spec = inst.vrptw(60, 5, n_depots=2, seed=4)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    isl = LateAcceptance(6, 0.2, None, [0.5, 0.5, 0.0, 0.0, 0.0, 0.0], 3, reference_noop_moves=False,
                         scoring="delta", chain_steps_per_launch=4).build_agent(gp, n_islands=6, seed=3)
    assert isl.step_path == "vrp_chain"
    _, s0 = isl.best(-1)
    prev = None
    for _ in range(8):
        isl.step(25)
        vec, sc = isl.best(-1)
        assert np.array_equal(sc, oracle.score_round(op.score_incremental(vec, [[]])[0], spec.score_precision))
        if prev is not None:
            assert oracle.score_cmp(sc, prev) <= 0
        prev = sc
        for i in range(6):
            cv, cs = isl.current(i)
            assert np.array_equal(cs, oracle.score_round(op.score_incremental(cv, [[]])[0], spec.score_precision))
    assert oracle.score_cmp(prev, s0) < 0
    isl.close(); gp.close()


# ---- fixed-point TabuSearch step over many steps and several waves of CTAs ------------------------------
@pytest.mark.parametrize("exact", [True, False], ids=["exact-sums", "tree-sums"])
def test_fixed_point_long_run_keeps_every_island_consistent(exact, oracle):
    """More islands than the GPU holds at once (several waves of CTAs), hundreds of steps with
    migration and global-top adoption: every island's tour stays a permutation, its stored score is
    the oracle's score of its stored tour, and the cached edge lengths never drift (the stored score
    is folded from them)."""
    spec = inst.tsp(300, seed=13)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    gp.set_exact_sums(exact)
    I = 1500
    isl = TabuSearch(256, 0.5, True, None, [0.0, 0.4, 0.0, 0.0, 0.2, 0.4], 5, scoring="delta").build_agent(
        gp, n_islands=I, seed=77)
    assert isl.step_path == "fused_fixed"
    prev = None
    for rnd in range(4):
        isl.step(150)
        gv, gs = isl.best(-1)
        assert sorted(gv.tolist()) == list(range(1, 300))
        want = oracle.score_round(op.score_incremental(gv, [[]])[0], spec.score_precision)
        assert gs[0] == 0.0 and abs(gs[1] - want[1]) <= (0.0 if exact else QUANTUM)
        if prev is not None:
            assert oracle.score_cmp(gs, prev) <= 0
        prev = gs
        for i in list(range(0, I, 97)) + [I - 1]:
            cv, cs = isl.current(i)
            assert sorted(cv.tolist()) == list(range(1, 300)), (rnd, i)
            want = oracle.score_round(op.score_incremental(cv, [[]])[0], spec.score_precision)
            assert cs[0] == 0.0 and abs(cs[1] - want[1]) <= (0.0 if exact else QUANTUM), (rnd, i, cs, want)
            bv, bs = isl.best(i)
            want = oracle.score_round(op.score_incremental(bv, [[]])[0], spec.score_precision)
            assert bs[0] == 0.0 and abs(bs[1] - want[1]) <= (0.0 if exact else QUANTUM)
            assert oracle.score_cmp(gs, bs) <= 0
    assert isl.stats()["steps"] == 600
    isl.close(); gp.close()
