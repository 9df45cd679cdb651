"""Pins the oracle against everything the reference's own unit tests fix for this path
 plus closed-form known answers verified by hand.

Reference tests mirrored here:
  greyjack/src/variables/gj_integer.rs:141-181   (frozen, clamp, inverse_transform)
  greyjack/src/variables/gj_float.rs:167-194     (frozen, clamp)
  greyjack/src/score_calculation/scores/simple_score.rs:108-152
  greyjack/src/score_calculation/scores/hard_soft_score.rs:130-183
"""
import math

import numpy as np
import pytest

from greyjack_b200 import instances as inst


def test_gj_integer_fix_value(oracle):
    # gj_integer.rs:157-168: GJInteger::new(Some(1), -1, 1, false): -100 -> -1, 100 -> 1
    assert oracle.fix_integer(-100.0, -1.0, 1.0) == -1.0
    assert oracle.fix_integer(100.0, -1.0, 1.0) == 1.0


def test_gj_integer_inverse_transform(oracle):
    # gj_integer.rs:170-181: 4.4 -> 4, 4.6 -> 5 with bounds [-10, 10]
    assert oracle.inverse_transform_integer(4.4, -10.0, 10.0) == 4
    assert oracle.inverse_transform_integer(4.6, -10.0, 10.0) == 5


def test_gj_integer_frozen(oracle):
    # gj_integer.rs:141-147: frozen -> initial value, whatever is proposed
    assert oracle.fix_integer(-7.0, -1.0, 1.0, frozen=True, initial=1.0) == 1.0
    assert oracle.inverse_transform_integer(0.2, -1.0, 1.0, frozen=True, initial=1.0) == 1


def test_gj_float_fix_value(oracle):
    # gj_float.rs:183-194
    assert oracle.fix_float(-100.0, -1.0, 1.0) == -1.0
    assert oracle.fix_float(100.0, -1.0, 1.0) == 1.0
    assert oracle.fix_float(0.25, -1.0, 1.0) == 0.25
    assert oracle.fix_float(5.0, -1.0, 1.0, frozen=True, initial=1.0) == 1.0


def test_rint_known_answers(oracle):
    # math_utils.rs:6-8: ties go to ceil
    assert oracle.rint(4.4) == 4.0
    assert oracle.rint(4.6) == 5.0
    assert oracle.rint(4.5) == 5.0
    assert oracle.rint(2.5) == 3.0
    r = oracle.rint(-0.5)
    assert r == 0.0 and math.copysign(1.0, r) == -1.0  # -0.0
    assert oracle.rint(-1.5) == -1.0
    assert oracle.rint(7.0) == 7.0


def test_round_known_answers(oracle):
    # math_utils.rs:10-13: truncation toward -inf
    assert oracle.round_(1.23456, 3) == 1.234
    assert oracle.round_(2.9999, 3) == 2.999
    assert oracle.round_(12.0005, 3) == 12.0
    assert oracle.round_(-1.2345, 3) == -2.0 + math.floor((-1.2345 + 2.0) * 1000.0) / 1000.0
    assert oracle.round_(5.75, 0) == 5.0
    # idempotent on its own output (location.rs:47 + domain_builder.rs:42-46 apply it twice)
    for x in (0.001, 17.329, 999.999, 123.4567):
        once = oracle.round_(x, 3)
        assert oracle.round_(once, 3) == once or abs(oracle.round_(once, 3) - once) <= 0.001


def test_simple_score_impl_and_comparison(oracle):
    # simple_score.rs:108-141
    assert oracle.fitness([9.0]) == 0.9
    small, null, large = [-10.0], [0.0], [10.0]
    assert oracle.score_cmp(small, large) < 0
    assert oracle.score_le(small, large)
    assert oracle.score_cmp(null, null) == 0
    assert oracle.score_cmp(large, null) > 0
    assert oracle.score_le(large, large)
    v = np.arange(10, dtype=np.float64)[::-1].copy()
    assert np.array_equal(oracle.sort_scores(v)[:, 0], np.arange(10, dtype=np.float64))


def test_hard_soft_score_impl_and_comparison(oracle):
    # hard_soft_score.rs:130-171
    assert oracle.fitness([0.0, 9.0]) == 0.45
    small, null, large = [-1.0, -1.0], [0.0, 0.0], [0.0, 0.1]
    assert oracle.score_cmp(small, large) < 0
    assert oracle.score_le(small, large)
    assert oracle.score_cmp(null, null) == 0
    assert oracle.score_cmp(large, null) > 0
    assert oracle.score_le(large, large)
    v1 = np.array([[i, 2 * i] for i in range(10)], dtype=np.float64)
    assert np.array_equal(oracle.sort_scores(v1[::-1].copy()), v1)
    v2 = np.array([[0, i] for i in range(10)], dtype=np.float64)
    assert np.array_equal(oracle.sort_scores(v2[::-1].copy()), v2)


def test_hard_medium_soft_lexicographic(oracle):
    # hard_medium_soft_score.rs:93-113
    assert oracle.score_cmp([0, 5, 1], [0, 5, 2]) < 0
    assert oracle.score_cmp([0, 6, 0], [0, 5, 9]) > 0
    assert oracle.score_cmp([1, 0, 0], [0, 9, 9]) > 0
    assert oracle.score_le([0, 5, 2], [0, 5, 2])
    assert not oracle.score_le([0, 5, 2.5], [0, 5, 2])


def test_score_round(oracle):
    # hard_soft_score.rs:76-79 with Solver precision [3, 3] (examples/tsp/src/main.rs:56)
    s = oracle.score_round([[2.0, 1234.56789]], [3, 3])
    assert s[0, 0] == 2.0 and s[0, 1] == 1234.567


# ---- closed-form scorer answers (SURVEY.md section 8c "golden vectors to create") ------

def _nq(oracle, n):
    spec = inst.nqueens(n)
    return spec, oracle.OracleProblem(spec)


@pytest.mark.parametrize("n", [4, 8, 64, 256])
def test_nqueens_closed_forms(oracle, n):
    spec, op = _nq(oracle, n)
    ident = np.arange(n, dtype=np.float64)
    # rows 0..n-1 on the main diagonal: all rows distinct, all desc ids distinct (2i),
    # asc ids all equal (0) -> n-1 conflicts
    assert op.score_plain(ident)[0, 0] == n - 1
    assert op.score_plain(ident[::-1].copy())[0, 0] == n - 1
    # all queens on one row: rows n-1 conflicts; both diagonals distinct
    assert op.score_plain(np.full(n, 3.0 if n > 3 else 0.0))[0, 0] == n - 1
    # incremental, no deltas == plain
    assert op.score_incremental(ident, [[]])[0, 0] == n - 1


def test_nqueens_8_solution(oracle):
    spec, op = _nq(oracle, 8)
    sol = np.array([0, 4, 7, 5, 2, 6, 1, 3], dtype=np.float64)
    assert op.score_plain(sol)[0, 0] == 0.0
    # moving one queen onto row 0 of column 1: rows dup (+1), diagonals
    bad = op.score_incremental(sol, [[(1, 0.0)]])[0, 0]
    rows = sol.copy(); rows[1] = 0
    cols = np.arange(8)
    expect = (8 - len(set(rows))) + (8 - len(set(cols + rows))) + (8 - len(set(cols - rows)))
    assert bad == expect


def test_tsp_collinear_closed_form(oracle):
    # cities on a line at x = 0,1,...,n-1: tour 1,2,..,n-1 has length 2(n-1)
    n = 12
    xy = np.stack([np.arange(n, dtype=np.float64), np.zeros(n)], axis=1)
    spec = inst.tsp(n, seed=1, greedy=False)
    spec.coords = xy
    spec.distance_matrix = inst.distance_matrix(xy)
    op = oracle.OracleProblem(spec)
    tour = np.arange(1, n, dtype=np.float64)
    s = op.score_plain(tour)[0]
    assert s[0] == 0.0 and s[1] == 2.0 * (n - 1)
    s = op.score_plain(tour[::-1].copy())[0]
    assert s[0] == 0.0 and s[1] == 2.0 * (n - 1)
    # duplicate injection: k copies of the same city -> hard = k
    dup = tour.copy(); dup[3] = dup[0]; dup[7] = dup[0]
    assert op.score_plain(dup)[0, 0] == 2.0
    assert op.score_incremental(tour, [[(3, tour[0]), (7, tour[0])]])[0, 0] == 2.0


def test_tsp_unit_square_grid(oracle):
    # 3x3 grid, boustrophedon tour from the corner depot: 8 unit edges + return sqrt(...)
    pts = [(x, y) for y in range(3) for x in range(3)]
    xy = np.array(pts, dtype=np.float64)
    spec = inst.tsp(9, seed=1, greedy=False)
    spec.distance_matrix = inst.distance_matrix(xy)
    op = oracle.OracleProblem(spec)
    order = [1, 2, 5, 4, 3, 6, 7, 8]  # snake
    s = op.score_plain(np.array(order, dtype=np.float64))[0]
    # D is truncated to 3 decimals TWICE (location.rs:47, then domain_builder.rs:42-46) and
    # the truncation is not idempotent in binary: 2.8284.. -> 2.828 -> 2.827
    back = oracle.round_(oracle.round_(math.sqrt(8.0), 3), 3)
    assert back == 2.827
    assert s[0] == 0.0
    assert abs(s[1] - (8.0 + back)) < 1e-12


def test_cvrp_hand_instance(oracle):
    # 1 depot + 4 customers on a line, 2 vehicles of capacity 10
    xy = np.array([[0, 0], [1, 0], [2, 0], [3, 0], [4, 0]], dtype=np.float64)
    spec = inst.cvrp(4, 2, seed=2, greedy=False)
    spec.distance_matrix = inst.distance_matrix(xy)
    spec.demand = np.array([0, 4, 4, 4, 4], dtype=np.uint64)
    spec.vehicle_capacity = np.array([10, 10], dtype=np.uint64)
    op = oracle.OracleProblem(spec)
    # vehicle 0: customers 1,2 ; vehicle 1: customers 3,4  -> no overflow
    x = np.array([0, 1, 0, 2, 1, 3, 1, 4], dtype=np.float64)
    s = op.score_plain(x)[0]
    assert list(s) == [0.0, 0.0, (1 + 1 + 2) + (3 + 1 + 4)]
    si = op.score_incremental(x, [[]])[0]
    assert list(si) == list(s)
    # everything on vehicle 0: load 16 > 10 -> overflow 6
    x2 = np.array([0, 1, 0, 2, 0, 3, 0, 4], dtype=np.float64)
    s2 = op.score_plain(x2)[0]
    assert list(s2) == [6.0, 0.0, 8.0]
    # duplicate customer: 1000 per duplicate, plus capacity of what is actually carried
    x3 = np.array([0, 1, 0, 1, 1, 3, 1, 4], dtype=np.float64)
    s3 = op.score_plain(x3)[0]
    assert s3[0] == 1000.0
    # route order is stop order within the vehicle, not customer order
    x4 = np.array([0, 2, 0, 1, 1, 4, 1, 3], dtype=np.float64)
    s4 = op.score_plain(x4)[0]
    assert s4[2] == (2 + 1 + 1) + (4 + 1 + 3)


def test_vrptw_hand_instance_variants(oracle):
    # Q3: the three lateness rules differ.  One vehicle, two customers.
    xy = np.array([[0, 0], [1, 0], [2, 0]], dtype=np.float64)
    for service_variant in (False, True):
        spec = inst.vrptw(2, 1, n_depots=1, seed=3, service_variant=service_variant, greedy=False)
        spec.distance_matrix = inst.distance_matrix(xy)
        spec.demand = np.array([0, 1, 1], dtype=np.uint64)
        spec.vehicle_capacity = np.array([10], dtype=np.uint64)
        spec.work_day_start = np.array([0], dtype=np.uint64)
        spec.work_day_end = np.array([100], dtype=np.uint64)
        spec.tw_start = np.array([0, 10, 0], dtype=np.uint64)
        spec.tw_end = np.array([0, 15, 20], dtype=np.uint64)
        spec.service_time = np.array([0, 10, 95], dtype=np.uint64)
        op = oracle.OracleProblem(spec)
        x = np.array([0, 1, 0, 2], dtype=np.float64)
        isc = op.score_incremental(x, [[]])[0]
        psc = op.score_plain(x)[0]
        # stop 1: arrive max(0,10)=10, service 10 -> leaves 20; stop 2: arrive 20, leaves 115
        if service_variant:
            # arrival > end + service ?  10 > 25 no ; 20 > 115 no ; day end: 115 > 100 -> 15
            assert isc[1] == 15.0
        else:
            # arrival + service > end ?  20 > 15 -> 5 ; 115 > 20 -> 95 ; day end 15
            assert isc[1] == 5.0 + 95.0 + 15.0
        # PSC skips the last stop of the route entirely (plain_score_calculator.rs:205-216):
        # only stop 1 is visited: arrival 10 > 15+10 no; leaves 20 <= 100 -> 0
        assert psc[1] == 0.0
        assert psc[0] == isc[0] == 0.0 and psc[2] == isc[2] == 4.0


def test_simulated_annealing_rule(oracle):
    """simulated_annealing_base.rs:198-233 by hand: cooling floor, accomplish-rate schedule, and
    u < e^(-dE0/T0) * e^(-dE1/T1)."""
    import math
    acc, t, p = oracle.sa_accept([0.0, 12.0], [0.0, 10.0], [1.0, 4.0], 0.5, 1.0, 0.3)
    assert list(t) == [0.5, 2.0]
    assert p == pytest.approx(math.exp(-1.0), rel=1e-15) and acc            # 0.3 < 0.3679
    acc, t, p = oracle.sa_accept([0.0, 12.0], [0.0, 10.0], [1.0, 4.0], 0.5, 1.0, 0.4)
    assert not acc
    acc, t, p = oracle.sa_accept([0.0, 9.0], [0.0, 10.0], [1.0, 4.0], None, 0.25, 0.999)
    assert list(t) == [0.25, 0.25] and p > 1.0 and acc                       # improving moves always pass
    _, t, _ = oracle.sa_accept([1.0], [1.0], [1.5e-6], 0.5, 1.0, 0.0)
    assert list(t) == [1e-7]                                                 # < 1e-6 -> 1e-7 floor


# ---- GeneticAlgorithm decisions (genetic_algorithm_base.rs:83-134) with explicit draws ----------------
def test_ga_select_and_cross_restatement(oracle):
    # last_top_id = ceil(p * pop); best: U[0, last_top) ; worst: U[pop - last_top, pop)
    assert oracle.ga_select(0.1, 3, 256) == (3, 26)
    assert oracle.ga_select(0.1, 3, 256, worst=True) == (256 - 26 + 3, 26)
    assert oracle.ga_select(0.000001, 0, 256) == (0, 1)              # the smallest p still reaches rank 0
    assert oracle.ga_select(0.1, 26, 256)[0] == -1                   # outside Uniform::new(0, last_top_id)
    # one weight for every gene, rint-ed on discrete columns (ties -> ceil): parents swap or stay
    a, b = [1.0, 2.0, 3.0], [4.0, 5.0, 6.0]
    c1, c2 = oracle.ga_cross(a, b, 0.4)
    assert c1.tolist() == b and c2.tolist() == a
    c1, c2 = oracle.ga_cross(a, b, 0.5)
    assert c1.tolist() == a and c2.tolist() == b
    c1, c2 = oracle.ga_cross(a, b, 1.0)
    assert c1.tolist() == a and c2.tolist() == b


def test_la_and_ga_baseline_drivers_make_progress(oracle):
    from greyjack_b200 import instances as inst
    spec = inst.nqueens(64)
    op = oracle.OracleProblem(spec)
    s0 = op.score_incremental(spec.initial, [[]])[0]
    n, secs, best = op.bench_la(spec.initial, 16, 4000, 2, 1, [0, 1.0, 0, 0, 0, 0], None)
    assert n == 8000 and secs > 0 and best[0] < s0[0]
    spec = inst.cvrp(40, 4, seed=2, greedy=False)
    op = oracle.OracleProblem(spec)
    n, secs, best = op.bench_ga(64, 0.5, 0.2, 40, 2, 3, [1 / 6.0] * 6, [0, 0, 3])
    assert n == 2 * 40 * 64
    n2, _, best2 = op.bench_ga(64, 0.5, 0.2, 2, 2, 3, [1 / 6.0] * 6, [0, 0, 3])
    assert oracle.score_cmp(best, best2) <= 0                        # more generations never end worse (same seed)
