"""Constraint programs (csrc/gj_program.cu): the reference's constraint registry
(plain_score_calculator.rs:20-94) with constraints written as terms of four relational primitives.
The examples' constraints re-expressed as programs must score BIT-identically to the hard-coded
kernels and to the oracle's restatement of the Polars queries; the registry calls (add / remove /
set weights) must behave like the reference's HashMaps."""
import numpy as np
import pytest

import greyjack_b200 as gj
from greyjack_b200 import Problem, instances as inst
from greyjack_b200.program import DISTINCT_DEFICIT, GATHER_FOLD, ConstraintProgram, TermSpec
from helpers import permutation_samples, random_samples

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mk,make", [
    (lambda: inst.nqueens(64), gj.nqueens_program),
    (lambda: inst.nqueens(257, seed=3), gj.nqueens_program),
    (lambda: inst.tsp(200, seed=3), gj.tsp_program),
    (lambda: inst.tsp(1000, seed=1), gj.tsp_program),
    (lambda: inst.cvrp(60, 6, seed=2), gj.vrp_program),
    (lambda: inst.vrptw(80, 6, n_depots=2, seed=3, service_variant=False), gj.vrp_program),
    (lambda: inst.vrptw(120, 9, n_depots=3, seed=5), gj.vrp_program),
], ids=["nqueens64", "nqueens257", "tsp200", "tsp1000", "cvrp60", "vrptw80", "vrpsvc120"])
def test_example_constraints_as_programs_are_bit_identical(mk, make, oracle):
    spec = mk()
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    prog = make(gp)
    rng = np.random.default_rng(11)
    x = np.concatenate([random_samples(spec, 48, rng), permutation_samples(spec, 48, rng), spec.initial[None, :]])
    got = prog.get_score(x)
    assert np.array_equal(got, gp.request_score_plain(x))         # the hard-coded kernels
    assert np.array_equal(got, op.score_plain(x))                 # the oracle's PSC restatement
    prog.close(); gp.close()


def test_registry_semantics(oracle):
    """add_constraint gives a new name weight 1.0 and keeps the weight of a known one (:29-34);
    set_constraint_weights replaces the whole map (:40-42; a registered constraint without a weight
    is an error -- the reference panics on the missing key); remove_constraint of an unknown name is a
    no-op (:36-38)."""
    spec = inst.tsp(120, seed=4)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    rng = np.random.default_rng(2)
    x = random_samples(spec, 32, rng)
    base = op.score_plain(x)
    prog = ConstraintProgram(gp, 2)
    with pytest.raises(gj.GjError):
        prog.get_score(x)                                         # nothing registered
    prog.add_constraint("no_duplicating_stops", 0, [TermSpec(DISTINCT_DEFICIT)])
    only_hard = prog.get_score(x)
    assert np.array_equal(only_hard[:, 0], base[:, 0]) and not only_hard[:, 1].any()
    prog.add_constraint("minimize_distance", 1, [TermSpec(GATHER_FOLD)])
    assert prog.n_constraints == 2 and np.array_equal(prog.get_score(x), base)
    prog.set_constraint_weights({"no_duplicating_stops": 10.0, "minimize_distance": 0.5})
    w = prog.get_score(x)
    assert np.array_equal(w[:, 0], 10.0 * base[:, 0]) and np.array_equal(w[:, 1], 0.5 * base[:, 1])
    # the hard-coded scorer with the same weights agrees
    gp.set_constraint_weights([10.0, 0.5])
    assert np.array_equal(w, gp.request_score_plain(x))
    prog.add_constraint("minimize_distance", 1, [TermSpec(GATHER_FOLD)])     # re-registering keeps the weight
    assert np.array_equal(prog.get_score(x), w)
    with pytest.raises(gj.GjError):
        prog.set_constraint_weights({"minimize_distance": 1.0})             # a constraint left without a weight
    prog.remove_constraint("not there")
    prog.remove_constraint("no_duplicating_stops")
    assert prog.n_constraints == 1
    prog.set_constraint_weights({"minimize_distance": 2.0})
    r = prog.get_score(x)
    assert not r[:, 0].any() and np.array_equal(r[:, 1], 2.0 * base[:, 1])
    # a constraint the examples do not have: stops must differ from their predecessor's id + 1 ... any
    # key expression works -- here "distinct (value - row)" as a second hard constraint on its own name
    prog.add_constraint("distinct_offsets", 0, [TermSpec(DISTINCT_DEFICIT, key_value_coef=1, key_index_coef=-1)])
    r2 = prog.get_score(x)
    dec = np.clip(np.floor(x + 0.5), spec.lower_bounds, spec.upper_bounds)
    want = np.array([x.shape[1] - len(set((dec[j] - np.arange(x.shape[1])).tolist())) for j in range(x.shape[0])], dtype=np.float64)
    assert np.array_equal(r2[:, 0], want) and np.array_equal(r2[:, 1], r[:, 1])
    prog.close(); gp.close()


def test_unsupported_layouts_fail_loudly():
    spec = inst.tsp(50, seed=1)
    gp = Problem(spec)
    prog = ConstraintProgram(gp, 2)
    with pytest.raises(gj.GjError):
        prog.add_constraint("cap", 0, [TermSpec(2)])                       # SEGMENT_OVER_CAP without a segment column
    with pytest.raises(gj.GjError):
        prog.add_constraint("lvl", 5, [TermSpec(DISTINCT_DEFICIT)])         # level out of range
    with pytest.raises(gj.GjError):
        prog.add_constraint("seg", 0, [TermSpec(GATHER_FOLD, value_offset=1, value_stride=2, seg_stride=2)])   # not a VRP problem
    prog.close(); gp.close()
