"""Solution wire format: the Value of Agent::convert_to_json
(agent_base.rs:523-535) and its way back in as InitialSolutionVariants::CotwinValuesVector."""
import json

import numpy as np
import pytest

from greyjack_b200 import instances as inst, wire


def test_variable_names_follow_the_requester_enumeration():
    # oop_score_requester.rs:93-123: "<group>: <running variable index>--><attribute>"
    assert wire.variable_names(inst.nqueens(4)) == [f"queens: {i}-->row_id" for i in range(4)]
    assert wire.variable_names(inst.tsp(5, seed=1))[:2] == ["path_stops: 0-->location_vec_id",
                                                            "path_stops: 1-->location_vec_id"]
    names = wire.variable_names(inst.cvrp(3, 2, seed=1))
    assert names == ["planning_stops: 0-->vehicle_id", "planning_stops: 1-->customer_id",
                     "planning_stops: 2-->vehicle_id", "planning_stops: 3-->customer_id",
                     "planning_stops: 4-->vehicle_id", "planning_stops: 5-->customer_id"]


@pytest.mark.parametrize("mk,score", [
    (lambda: inst.nqueens(8), [3.0]),
    (lambda: inst.tsp(12, seed=3), [0.0, 1234.567]),
    (lambda: inst.vrptw(10, 3, n_depots=2, seed=5), [1000.0, 25.0, 987.25]),
], ids=["simple", "hard-soft", "hard-medium-soft"])
def test_value_round_trip(mk, score):
    spec = mk()
    text = wire.solution_to_json(spec, spec.initial, score)
    value = json.loads(text)
    # shape of json!((Vec<(String, AnyValue)>, score))
    assert isinstance(value, list) and len(value) == 2 and len(value[0]) == spec.n_vars
    assert value[0][0][1] == {"Int64": int(spec.initial[0])}
    assert list(value[1].keys()) == wire._SCORE_FIELDS[spec.levels]
    vec, sc = wire.solution_from_json(spec, text)
    assert np.array_equal(vec, spec.initial) and np.array_equal(sc, np.array(score))


def test_mismatched_solution_is_rejected():
    a, b = inst.tsp(12, seed=3), inst.tsp(13, seed=3)
    v = wire.solution_to_value(a, a.initial, [0.0, 1.0])
    with pytest.raises(ValueError):
        wire.solution_from_value(b, v)
    with pytest.raises(ValueError):
        wire.solution_from_value(inst.nqueens(11), v)
    v[0][0][1] = {"String": "x"}
    with pytest.raises(ValueError):
        wire.solution_from_value(a, v)
