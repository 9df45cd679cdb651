"""Solver.solve host loop (termination strategies, observers) over device-resident islands."""
import numpy as np
import pytest

from greyjack_b200 import (LateAcceptance, Problem, ScoreLimit, SimulatedAnnealing, Solver, StepsLimit,
                           TabuSearch, TimeSpentLimit, instances as inst)

pytestmark = pytest.mark.gpu


class Collect:
    def __init__(self):
        self.updates = []

    def update(self, payload):
        self.updates.append(payload)


def test_solve_nqueens_to_zero_conflicts(oracle):
    spec = inst.nqueens(64, seed=45)
    gp = Problem(spec)
    obs = Collect()
    vars_, score = Solver.solve(gp, TabuSearch(256, 0.2, True, None, [0, 1.0, 0, 0, 0, 0], 10, scoring="delta"),
                                n_jobs=16, termination_strategy=ScoreLimit((0.0,)), observers=[obs], seed=3)
    assert score[0] == 0.0
    assert np.array_equal(oracle.OracleProblem(spec).score_incremental(vars_, [[]])[0], score)
    assert obs.updates and obs.updates[-1]["score"] == [0.0]
    scores = [u["score"][0] for u in obs.updates]
    assert scores == sorted(scores, reverse=True)           # observers only see improvements
    gp.close()


@pytest.mark.parametrize("builder", [
    LateAcceptance(16, 0.2, None, [0, 0.5, 0, 0, 0, 0.5], 10, scoring="delta"),
    SimulatedAnnealing([1.0, 50.0], None, 0.2, None, [0, 0.5, 0, 0, 0, 0.5], 10),
], ids=["la", "sa-accomplish-rate"])
def test_solve_steps_limit_and_time_limit(builder, oracle):
    spec = inst.tsp(120, seed=5)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    start = op.score_incremental(spec.initial, [[]])[0]
    vars_, score = Solver.solve(gp, builder, n_jobs=32, termination_strategy=StepsLimit(400), seed=1)
    assert oracle.score_cmp(score, start) < 0
    assert np.array_equal(oracle.score_round(op.score_incremental(vars_, [[]])[0], spec.score_precision), score)
    vars2, score2 = Solver.solve(gp, builder, n_jobs=32, termination_strategy=TimeSpentLimit(200), seed=1)
    assert oracle.score_cmp(score2, start) < 0
    gp.close()


def test_warm_start_from_the_reference_solution_value(oracle):
    """Multi-stage solving as in the reference: the Value of one solve (Agent::convert_to_json) goes
    back in as InitialSolutionVariants::CotwinValuesVector (solver.rs:108-119)."""
    from greyjack_b200 import wire
    spec = inst.cvrp(40, 4, seed=8)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    builder = TabuSearch(128, 0.2, True, None, [0.5, 0.5, 0, 0, 0, 0], 5, scoring="delta")
    v1, s1 = Solver.solve(gp, builder, n_jobs=4, termination_strategy=StepsLimit(30), seed=1)
    text = wire.solution_to_json(spec, v1, s1)
    v2, s2 = Solver.solve(gp, builder, n_jobs=4, termination_strategy=StepsLimit(30), seed=2,
                          initial_solution=text)
    assert oracle.score_cmp(s2, s1) <= 0                     # the second stage starts where the first ended
    assert np.array_equal(oracle.score_round(op.score_incremental(v2, [[]])[0], spec.score_precision), s2)
    gp.close()


def test_frozen_replanning_keeps_pinned_stops(oracle):
    """examples/vrp/src/main.rs:120-141 on the device: a plan is rebuilt into a domain
    (vrp_routes_from_solution = build_from_solution), a vehicle is dropped, the customers of another are
    pinned, and the second solve starts from that plan -- the pinned stops never move and the scorer
    ignores candidate values on frozen variables."""
    import numpy as np
    from greyjack_b200 import Problem, Solver, StepsLimit, TabuSearch, instances as inst
    spec = inst.cvrp(40, 5, seed=9)
    gp = Problem(spec)
    v1, s1 = Solver.solve(gp, TabuSearch(128, 0.2, True, None, [0.5, 0.5, 0, 0, 0, 0], 10, scoring="delta"),
                          n_jobs=8, termination_strategy=StepsLimit(60), seed=3)
    gp.close()
    routes = inst.vrp_routes_from_solution(spec, v1)
    re = inst.vrp_replanning_spec(spec, routes, frozen_vehicles=[1], drop_vehicles=[0])
    op = oracle.OracleProblem(re)
    gp2 = Problem(re)
    v2, s2 = Solver.solve(gp2, TabuSearch(128, 0.2, True, None, [0.5, 0.5, 0, 0, 0, 0], 10, scoring="delta"),
                          n_jobs=8, termination_strategy=StepsLimit(60), seed=4)
    frozen = re.frozen.astype(bool)
    assert frozen.any() and np.array_equal(v2[frozen], re.initial[frozen])
    want = op.score_incremental(v2, [[]])[0]
    assert np.array_equal(s2, oracle.score_round(want, re.score_precision)) or np.array_equal(s2, want)
    # the dropped vehicle's customers were re-assigned: every customer appears exactly once
    assert sorted(v2[1::2].tolist()) == list(range(re.n_depots, re.n_locations))
    assert s2[0] >= 0
    gp2.close()
