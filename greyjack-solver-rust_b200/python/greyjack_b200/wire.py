"""Solution wire format of the reference (host side only).

Agent::convert_to_json (agents/base/agent_base.rs:523-535) serialises an individual as
    json!((Vec<(variable_name, AnyValue)>, score))
i.e.  [[["planning_stops: 0-->vehicle_id", {"Int64": 3}], ...], {"hard_score": 0.0, ...}]
with variable names built by the score requester (oop_score_requester.rs:93-123):
    "<entities group>: <running variable index>-->" + "<attribute>"
The same Value is what Solver::solve returns, what observers receive (ObserverTrait::update) and
what InitialSolutionVariants::CotwinValuesVector feeds back in for a warm start
(solver/solver.rs:108-119 -> DomainBuilder::build_from_solution, e.g.
examples/tsp/src/persistence/domain_builder.rs:58-77).  Nothing here touches the GPU."""
from __future__ import annotations

import json
from typing import List, Sequence, Tuple

import numpy as np

from . import instances as inst

# entity group / planning attributes of the four shipped examples, in to_vec() order
# (examples/*/src/cotwin/*.rs, examples/*/src/persistence/cotwin_builder.rs)
_NAMES = {
    inst.NQUEENS: ("queens", ["row_id"]),
    inst.TSP: ("path_stops", ["location_vec_id"]),
    inst.VRP: ("planning_stops", ["vehicle_id", "customer_id"]),
    inst.VRP_SERVICE: ("planning_stops", ["vehicle_id", "customer_id"]),
}
# field names of the score structs (score_calculation/scores/*.rs)
_SCORE_FIELDS = {1: ["simple_value"], 2: ["hard_score", "soft_score"],
                 3: ["hard_score", "medium_score", "soft_score"]}


def variable_names(spec) -> List[str]:
    """VariablesManager::get_variables_names_vec for a ProblemSpec of the shipped models."""
    group, attrs = _NAMES[spec.kind]
    return [f"{group}: {i}-->{attrs[i % len(attrs)]}" for i in range(spec.n_vars)]


def solution_to_value(spec, variable_values: Sequence[float], score: Sequence[float]):
    """The serde_json::Value of Agent::convert_to_json as plain Python lists / dicts."""
    vals = np.asarray(variable_values, dtype=np.float64)
    if vals.shape != (spec.n_vars,):
        raise ValueError(f"expected {spec.n_vars} variable values, got {vals.shape}")
    fields = _SCORE_FIELDS[spec.levels]
    pairs = [[name, {"Int64": int(v)}] for name, v in zip(variable_names(spec), np.rint(vals))]
    return [pairs, {f: float(s) for f, s in zip(fields, score)}]


def solution_to_json(spec, variable_values, score) -> str:
    return json.dumps(solution_to_value(spec, variable_values, score))


def solution_from_value(spec, value) -> Tuple[np.ndarray, np.ndarray]:
    """Inverse of solution_to_value: (variable_values f64 [n_vars], score f64 [levels]).  Names are
    checked against the spec (a solution of another model / size is rejected, where the reference's
    build_from_solution would index out of bounds)."""
    if isinstance(value, str):
        value = json.loads(value)
    pairs, score = value
    names = variable_names(spec)
    if len(pairs) != spec.n_vars:
        raise ValueError(f"solution has {len(pairs)} variables, the problem {spec.n_vars}")
    out = np.empty(spec.n_vars, dtype=np.float64)
    for i, (name, any_value) in enumerate(pairs):
        if name != names[i]:
            raise ValueError(f"variable {i} is {name!r}, expected {names[i]!r}")
        if not isinstance(any_value, dict) or len(any_value) != 1:
            raise ValueError(f"variable {i}: not an AnyValue: {any_value!r}")
        (tag, v), = any_value.items()
        if tag not in ("Int64", "Float64", "Int32", "UInt64", "UInt32"):
            raise ValueError(f"variable {i}: unsupported AnyValue::{tag}")
        out[i] = float(v)
    fields = _SCORE_FIELDS[spec.levels]
    return out, np.array([float(score[f]) for f in fields], dtype=np.float64)


def solution_from_json(spec, text: str):
    return solution_from_value(spec, json.loads(text))
