"""GPU parity: the CUDA scorers behind the C ABI (gj_score_plain / gj_score_incremental)
against the CPU oracle on identical candidate sets.  Integer levels bit-exact; the distance
level is bit-exact too in the default exact-sums mode (reference summation order) and within
helpers.SOFT_RTOL with gj_problem_set_exact_sums(0) (tree reduction)."""
import numpy as np
import pytest

from greyjack_b200 import Problem, instances as inst
from helpers import (assert_scores_match, permutation_samples, random_moves, random_samples)

pytestmark = pytest.mark.gpu


def _specs():
    return [
        ("nq8", lambda: inst.nqueens(8)),
        ("nq256", lambda: inst.nqueens(256)),
        ("tsp12", lambda: inst.tsp(12, seed=7)),
        ("tsp100", lambda: inst.tsp(100, seed=5)),
        ("tsp1000", lambda: inst.tsp(1000, seed=1)),
        ("cvrp40x5", lambda: inst.cvrp(40, 5, seed=2)),
        ("cvrp300x12", lambda: inst.cvrp(300, 12, seed=4)),
        ("vrptw120x8", lambda: inst.vrptw(120, 8, n_depots=3, seed=3, service_variant=False)),
        ("vrpsvc120x8", lambda: inst.vrptw(120, 8, n_depots=3, seed=3, service_variant=True)),
        ("vrpsvc500x20", lambda: inst.vrptw(500, 20, n_depots=5, seed=9, service_variant=True)),
    ]


@pytest.fixture(scope="module", params=_specs(), ids=lambda p: p[0])
def triple(request, oracle):
    spec = request.param[1]()
    gp = Problem(spec)
    yield spec, oracle.OracleProblem(spec), gp
    gp.close()


def test_plain_random_candidates(triple):
    spec, op, gp = triple
    rng = np.random.default_rng(11)
    x = np.concatenate([random_samples(spec, 96, rng), permutation_samples(spec, 32, rng)])
    assert_scores_match(gp.request_score_plain(x), op.score_plain(x), spec, soft_exact=True)


def test_plain_single_and_ragged_batches(triple):
    spec, op, gp = triple
    rng = np.random.default_rng(5)
    for S in (1, 3, 5, 33):
        x = permutation_samples(spec, S, rng)
        assert_scores_match(gp.request_score_plain(x), op.score_plain(x), spec, soft_exact=True)


def test_incremental_all_moves(triple):
    spec, op, gp = triple
    rng = np.random.default_rng(23)
    base = spec.initial.copy()
    deltas, kinds = random_moves(op, spec, base, 160, rng)
    deltas.append([])                                        # empty delta list = the base itself
    deltas.append([(i, float(base[i])) for i in range(spec.n_vars)])  # init_population form
    got = gp.request_score_incremental(base, deltas)
    want = op.score_incremental(base, deltas)
    assert_scores_match(got, want, spec, soft_exact=True)
    # the no-delta candidate equals the plain score of the base for CVRP/TSP/N-Queens
    if not spec.time_windowed:
        assert_scores_match(got[-2:-1], op.score_plain(base), spec, soft_exact=True)


def test_incremental_repeated_ids_last_wins(triple):
    spec, op, gp = triple
    base = spec.initial.copy()
    lo, hi = spec.lower_bounds, spec.upper_bounds
    d = [[(0, lo[0]), (0, hi[0]), (0, lo[0] + 1.0)],
         [(spec.n_vars - 1, hi[-1]), (0, hi[0]), (spec.n_vars - 1, lo[-1])]]
    # a long list with many repeats crossing the 32-wide application chunks
    rng = np.random.default_rng(3)
    ids = rng.integers(0, min(spec.n_vars, 6), size=100)
    d.append([(int(i), float(rng.integers(lo[i], hi[i] + 1))) for i in ids])
    assert_scores_match(gp.request_score_incremental(base, d), op.score_incremental(base, d), spec,
                        soft_exact=spec.kind >= inst.VRP)


def test_incremental_packed_equals_reference_layout(triple):
    """gj_score_incremental_packed (u32 ids, i32 decoded values) == gj_score_incremental, bit for bit,
    and == the oracle; repeated ids and out-of-bounds values included."""
    from greyjack_b200.problem import deltas_to_csr
    spec, op, gp = triple
    rng = np.random.default_rng(29)
    base = spec.initial.copy()
    deltas, _ = random_moves(op, spec, base, 120, rng)
    deltas.append([])
    ids = rng.integers(0, min(spec.n_vars, 6), size=70)
    deltas.append([(int(i), float(rng.integers(spec.lower_bounds[i] - 3, spec.upper_bounds[i] + 4))) for i in ids])
    o, i, v = deltas_to_csr(deltas)
    got = gp.request_score_incremental_packed(base, o, i.astype(np.uint32), np.rint(v).astype(np.int32))
    ref = gp.request_score_incremental_csr(base, o, i, v)
    assert np.array_equal(got, ref)
    assert_scores_match(got, op.score_incremental(base, deltas), spec, soft_exact=True)


def test_fast_sums_within_tolerance(triple):
    """gj_problem_set_exact_sums(0): tree-reduced distance sums, stated tolerance 1e-12 rel."""
    spec, op, gp = triple
    rng = np.random.default_rng(17)
    x = permutation_samples(spec, 48, rng)
    try:
        gp.set_exact_sums(False)
        assert_scores_match(gp.request_score_plain(x), op.score_plain(x), spec, soft_exact=False)
        base = spec.initial.copy()
        deltas, _ = random_moves(op, spec, base, 64, rng)
        assert_scores_match(gp.request_score_incremental(base, deltas), op.score_incremental(base, deltas),
                            spec, soft_exact=False)
    finally:
        gp.set_exact_sums(True)


def test_weights(triple):
    spec, op, gp = triple
    rng = np.random.default_rng(2)
    x = permutation_samples(spec, 16, rng)
    x[:, 0] = x[:, min(3, spec.n_vars - 1)]   # force a duplicate where the model counts them
    w = np.array([2.0, 3.0, 0.5, 4.0])
    old = spec.weights.copy()
    try:
        spec.weights = w
        op2 = type(op)(spec)
        gp.set_constraint_weights(w)
        got, want = gp.request_score_plain(x), op2.score_plain(x)
        for l in range(spec.levels - 1):
            assert np.array_equal(got[:, l], want[:, l])
        np.testing.assert_allclose(got[:, -1], want[:, -1], rtol=1e-12)
    finally:
        spec.weights = old
        gp.set_constraint_weights(old)


def test_frozen_variables(oracle):
    spec = inst.tsp(60, seed=8)
    spec.frozen = np.zeros(spec.n_vars, dtype=np.uint8)
    spec.frozen[[0, 5, 17]] = 1
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    rng = np.random.default_rng(4)
    x = random_samples(spec, 40, rng)
    assert_scores_match(gp.request_score_plain(x), op.score_plain(x), spec, soft_exact=True)
    base = spec.initial.copy()
    d = [[(0, 3.0), (5, 9.0), (1, 2.0)], [(17, 1.0)]]
    assert_scores_match(gp.request_score_incremental(base, d), op.score_incremental(base, d), spec, soft_exact=True)
    gp.close()


def test_device_distance_matrix_bit_exact(oracle):
    spec = inst.tsp(257, seed=12)
    gp = Problem(spec, use_coords=True)
    D = gp.distance_matrix()
    assert np.array_equal(D, spec.distance_matrix)
    assert np.array_equal(D, oracle.distance_matrix(spec.coords))
    gp.close()


def test_create_rejects_bad_input():
    from greyjack_b200 import GjError
    spec = inst.tsp(10, seed=1)
    spec.frozen = np.ones(spec.n_vars, dtype=np.uint8)
    spec.initial = np.full(spec.n_vars, np.nan)
    with pytest.raises(GjError, match="Frozen value must be initialized"):
        Problem(spec)
    spec = inst.tsp(10, seed=1)
    spec.upper_bounds = spec.upper_bounds + 5          # ids beyond the distance matrix
    with pytest.raises(GjError):
        Problem(spec)


def test_full_size_configs_properties(oracle):
    """BASELINE configs at full size: oracle spot-check + size-independent properties."""
    rng = np.random.default_rng(99)
    # C2: TSP 1000 -- relabelling invariance: reversing a tour keeps the length (symmetric D)
    spec = inst.tsp(1000, seed=1)
    gp = Problem(spec)
    x = permutation_samples(spec, 64, rng)
    s = gp.request_score_plain(x)
    srev = gp.request_score_plain(x[:, ::-1].copy())
    assert np.array_equal(s[:, 0], np.zeros(64))
    np.testing.assert_allclose(s[:, 1], srev[:, 1], rtol=1e-12)
    assert_scores_match(s[:8], oracle.OracleProblem(spec).score_plain(x[:8]), spec, soft_exact=True)
    gp.close()
    # C3: CVRP 2000 x 50 -- moving every stop to one vehicle keeps dup=0 and overflows by sum-cap
    spec = inst.cvrp(2000, 50, seed=2, greedy=False)
    gp = Problem(spec)
    x = permutation_samples(spec, 8, rng)
    x[:, 0::2] = 7
    s = gp.request_score_plain(x)
    over = float(spec.demand.sum()) - float(spec.vehicle_capacity[7])
    assert np.array_equal(s[:, 0], np.full(8, over))
    assert_scores_match(s[:2], oracle.OracleProblem(spec).score_plain(x[:2]), spec, soft_exact=True)
    gp.close()
