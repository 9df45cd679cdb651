"""VRP files, domain round trips and frozen replanning (host side):
examples/vrp/src/persistence/domain_builder.rs:18-141 (build_domain_from_scratch, build_from_solution),
:145-331 (read_vrp_file), cotwin_builder.rs:98-140 (is_already_initialized / frozen), main.rs:120-141."""
import numpy as np
import pytest

from greyjack_b200 import instances as inst

HAND = "\n".join([
    "NAME : toy-n6-k2",
    "COMMENT : hand made",
    "TYPE : CVRP",
    "DIMENSION : 6",
    "EDGE_WEIGHT_TYPE : EUC_2D",
    "CAPACITY : 10",
    "NODE_COORD_SECTION",
    "1 0.0 0.0 depot",
    "2   3.0   4.0",
    "3 6.0 8.0",
    "4 0.0 5.0 east",
    "5 5.0 0.0",
    "6 1.0 1.0",
    "DEMAND_SECTION",
    "1 0",
    "2 4",
    "3 7",
    "4 3",
    "5 6",
    "6 2",
    "DEPOT_SECTION",
    "1",
    "-1",
    "EOF",
    "",
])


def test_read_hand_made_file():
    meta, xy, ids, matrix, demand, depots = inst.read_vrp(HAND)
    assert meta["dataset_name"] == "toy-n6-k2" and meta["vehicles_count"] == "2" and meta["vehicles_capacity"] == "10"
    assert meta["distance_type"] == "EUC_2D" and matrix is None
    assert ids == [1, 2, 3, 4, 5, 6] and xy.shape == (6, 2) and depots == [1]
    assert demand[2] == [3, 7]


def test_domain_from_hand_made_file_scores_by_hand(oracle):
    spec = inst.vrp_from_file(HAND, greedy=False)
    assert (spec.n_locations, spec.n_depots, spec.n_vehicles, spec.n_vars) == (6, 1, 2, 10)
    assert spec.distance_matrix[0, 1] == 5.0 and spec.distance_matrix[1, 2] == 5.0 and not spec.time_windowed
    assert list(spec.groups) == ["vehicle_assignment", "customer_assignment", "common"]
    # vehicle 0: depot -> 1 -> 2 -> depot (demand 11 > 10: overflow 1, distance 5 + 5 + 10 = 20)
    # vehicle 1: depot -> 3 -> 4 -> 5 -> depot (demand 3 + 6 + 2 = 11: overflow 1)
    x = np.array([0, 1, 0, 2, 1, 3, 1, 4, 1, 5], dtype=np.float64)
    sc = oracle.OracleProblem(spec).score_plain(x[None, :])[0]
    d = spec.distance_matrix
    assert sc[0] == 2.0 and sc[1] == 0.0
    assert sc[2] == pytest.approx(20.0 + d[0, 3] + d[3, 4] + d[4, 5] + d[5, 0], rel=1e-15)
    assert inst.vrp_routes_from_solution(spec, x) == [[1, 2], [3, 4, 5]]


@pytest.mark.parametrize("tw", [False, True])
def test_write_read_round_trip(tw):
    a = inst.vrptw(40, 5, n_depots=2, seed=8, service_variant=False) if tw else inst.cvrp(30, 4, seed=3)
    b = inst.vrp_from_file(inst.write_vrp(a))
    assert b.n_vehicles == a.n_vehicles and b.n_depots == a.n_depots and b.time_windowed == a.time_windowed
    for f in ("distance_matrix", "demand", "tw_start", "tw_end", "service_time", "vehicle_depot", "vehicle_capacity",
              "lower_bounds", "upper_bounds", "initial"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    if tw:
        # a vehicle's work day is its depot's window (domain_builder.rs:66-71)
        assert np.array_equal(b.work_day_start, a.tw_start[a.vehicle_depot]) and np.array_equal(b.work_day_end, a.tw_end[a.vehicle_depot])
    # the service variant keeps its two groups disjoint (vrp_service cotwin_builder.rs:128-132)
    c = inst.vrp_from_file(inst.write_vrp(a), service_variant=True)
    assert list(c.groups) == ["vehicle_assignment", "customer_assignment"] and c.kind == inst.VRP_SERVICE


def test_explicit_matrix_file():
    text = ("NAME : m-k1\nEDGE_WEIGHT_TYPE : EXPLICIT\nCAPACITY : 5\nNODE_COORD_SECTION\n1 0 0\n2 1 0\n3 0 1\nEOF\n"
            "0 1.23456 2 x\n1.23456 0 3 x\n2 3 0 x\nEOF\nDEMAND_SECTION\n1 0\n2 1\n3 1\nDEPOT_SECTION\n1\n-1\nEOF\n")
    spec = inst.vrp_from_file(text, greedy=False)
    assert spec.distance_matrix[0, 1] == 1.234 and spec.distance_matrix[1, 2] == 3.0     # round(dm, 3) truncates


def test_bad_files():
    with pytest.raises(ValueError):
        inst.read_vrp("NAME : x-k2\nCAPACITY : 1\nEDGE_WEIGHT_TYPE : EUC_2D\n")          # no NODE_COORD_SECTION
    with pytest.raises(ValueError):
        inst.vrp_from_file(HAND.replace("2 4\n", ""))                                      # a demand line missing
    with pytest.raises(ValueError):
        inst.vrp_from_file(HAND.replace("3 7", "9 7"))                                     # id mismatch


def test_replanning_spec_pins_and_drops(oracle):
    """main.rs:120-141: solve, rebuild the domain, drop vehicle 0, pin the customers of the new vehicle 0,
    solve again from that plan."""
    spec = inst.cvrp(24, 4, seed=5)
    routes = inst.vrp_routes_from_solution(spec, spec.initial)
    assert sum(len(r) for r in routes) == 24
    re = inst.vrp_replanning_spec(spec, routes, frozen_vehicles=[1], drop_vehicles=[0])
    assert re.n_vehicles == 3 and re.upper_bounds[0] == 2.0
    n_kept = sum(len(routes[k]) for k in (1, 2, 3))
    # planning stop i = the i-th (vehicle, customer) pair of the plan, vehicle-major (cotwin_builder.rs:108-119)
    assert re.initial[0] == 0.0 and re.initial[1] == routes[1][0]
    assert np.isnan(re.initial[2 * n_kept:]).all()                    # the dropped vehicle's customers: None
    frozen_stops = np.nonzero(re.frozen[0::2])[0]
    assert len(frozen_stops) == len(routes[1]) and (re.frozen[0::2] == re.frozen[1::2]).all()
    # a frozen variable decodes to its initial value whatever the candidate says (gj_integer.rs:70-75)
    op = oracle.OracleProblem(re)
    x = np.where(np.isnan(re.initial), re.lower_bounds, re.initial)
    y = x.copy()
    y[2 * frozen_stops] = 2.0
    y[2 * frozen_stops + 1] = re.lower_bounds[1]
    assert np.array_equal(op.score_plain(x[None, :]), op.score_plain(y[None, :]))
