"""Pinning against the REFERENCE ITSELF (oracle/reference_dump/README.md).

tests/golden/reference_inputs/ holds instances + inputs written by make_reference_inputs.py; anyone
with cargo runs the reference-side dumper (the reference's own OOPScoreRequester / PSC / ISC / Mover) on
them and commits tests/golden/reference_outputs/<example>.json.  While those files are absent the
comparisons SKIP (parity stays "unpinned") and only the part that needs no
reference runs: the inputs are well formed, the committed instance files rebuild the exact problem the
inputs were generated for, and (GPU) the CUDA path agrees with the oracle on them."""
import json
import os

import numpy as np
import pytest

from greyjack_b200 import instances as inst

HERE = os.path.dirname(os.path.abspath(__file__))
IN = os.path.join(HERE, "golden", "reference_inputs")
OUT = os.path.join(HERE, "golden", "reference_outputs")
EXAMPLES = ["nqueens", "tsp", "vrp", "vrp_service"]


def _spec(example, doc):
    i = doc["instance"]
    if example == "nqueens":
        return inst.nqueens(int(i["n_queens"]), seed=int(i["seed"]))
    path = os.path.join(os.path.dirname(HERE), i["path"])
    with open(path) as f:
        text = f.read()
    if example == "tsp":
        return inst.tsp_from_tsplib(text, greedy=False)
    if example == "vrp":
        return inst.vrp_from_file(text, greedy=False)
    return inst.vrp_service_from_json(json.loads(text), greedy=False)


def _load(example):
    with open(os.path.join(IN, f"{example}.json")) as f:
        doc = json.load(f)
    spec = _spec(example, doc)
    deltas = [[(int(c), float(v)) for c, v in d] for d in doc["deltas"]]
    return doc, spec, np.array(doc["samples"]), np.array(doc["base"]), deltas


def _reference(example):
    path = os.path.join(OUT, f"{example}.json")
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path)} not produced yet: run the reference-side dumper (oracle/reference_dump/README.md)")
    with open(path) as f:
        return json.load(f)


def expected_variable_names(example, spec):
    """oop_score_requester.rs:93-123: "<group>: <i>--><attribute>", entity by entity, field by field"""
    if example == "nqueens":
        return [f"queens: {i}-->row_id" for i in range(spec.n_vars)]
    if example == "tsp":
        return [f"path_stops: {i}-->locations_vec_id" for i in range(spec.n_vars)]
    out = []
    for i in range(spec.n_vars // 2):
        out += [f"planning_stops: {2 * i}-->vehicle_id", f"planning_stops: {2 * i + 1}-->customer_id"]
    return out


@pytest.mark.parametrize("example", EXAMPLES)
def test_inputs_are_well_formed_and_the_oracle_runs_on_them(example, oracle):
    doc, spec, samples, base, deltas = _load(example)
    assert samples.shape[1] == spec.n_vars and len(base) == spec.n_vars and len(deltas) == 96
    op = oracle.OracleProblem(spec)
    plain = op.score_plain(samples)
    incr = op.score_incremental(base, deltas)
    assert plain.shape == (len(samples), spec.levels) and incr.shape == (96, spec.levels)
    assert np.isfinite(plain).all() and np.isfinite(incr).all()
    # an empty delta list scores the base itself; for every model but time-windowed VRP the two scorers agree (Q3)
    if not spec.time_windowed:
        assert np.array_equal(op.score_incremental(base, [[]])[0], op.score_plain(base[None, :])[0])


@pytest.mark.parametrize("example", EXAMPLES)
def test_oracle_matches_the_reference(example, oracle):
    ref = _reference(example)
    doc, spec, samples, base, deltas = _load(example)
    op = oracle.OracleProblem(spec)
    assert ref["variable_names"] == expected_variable_names(example, spec)
    want_plain, want_incr = np.array(ref["plain"]), np.array(ref["incremental"])
    got_plain, got_incr = op.score_plain(samples), op.score_incremental(base, deltas)
    L = spec.levels
    for l in range(L - 1 if L > 1 else 1):
        assert np.array_equal(got_plain[:, l], want_plain[:, l])
    if L > 1:
        if spec.kind in (inst.VRP, inst.VRP_SERVICE):
            np.testing.assert_allclose(got_plain[:, L - 1], want_plain[:, L - 1], rtol=1e-12, atol=0.0)   # Q7
        else:
            assert np.array_equal(got_plain[:, L - 1], want_plain[:, L - 1])
    assert np.array_equal(got_incr, want_incr)
    _check_moves(ref["moves"], np.array(doc["mover"]["candidate"]), spec, op)


def _check_moves(moves, cand, spec, op):
    """Mover::do_move outputs of the reference, one move kind at a time, against the oracle's mover run
    on the ids recovered from the changed columns."""
    groups = [np.asarray(g, dtype=np.int32) for g in spec.groups.values()]
    seen = set()
    for m in moves:
        if m["columns"] is None:
            continue
        kind, cols, vals = int(m["kind"]), [int(c) for c in m["columns"]], [float(v) for v in m["values"]]
        seen.add(kind)
        ok = False
        for g in groups:
            pos = {int(v): k for k, v in enumerate(g)}
            if not all(c in pos for c in cols):
                continue
            if kind == 0:
                ok = all(spec.lower_bounds[c] <= v <= spec.upper_bounds[c] and v == np.floor(v) for c, v in zip(cols, vals))
            elif kind == 1:
                res = op.move_swap(cand, g, [pos[c] for c in cols], True)
            elif kind in (2, 3):
                ok = all(v == cand[c] for c, v in zip(cols, vals))          # Q8: the incremental forms are no-ops
            elif kind == 4:
                lo, hi = pos[cols[0]], pos[cols[-1]]
                lo, hi = min(lo, hi), max(lo, hi)
                for a, b in ((lo, hi), (hi, lo)):
                    r = op.move_insertion(cand, g, a, b, True)
                    if r is not None and dict(zip(r[0].tolist(), op.fix_deltas(r[0], r[1]).tolist())) == dict(zip(cols, vals)):
                        ok = True
            else:
                lo, hi = sorted((pos[cols[0]], pos[cols[-1]]))
                res = op.move_inverse(cand, g, lo, hi, True)
            if kind in (1, 5) and res is not None:
                ok = dict(zip(res[0].tolist(), op.fix_deltas(res[0], res[1]).tolist())) == dict(zip(cols, vals))
            if ok:
                break
        assert ok, m
    assert seen == {0, 1, 2, 3, 4, 5}


@pytest.mark.gpu
@pytest.mark.parametrize("example", EXAMPLES)
def test_cuda_matches_the_oracle_and_the_reference_on_the_dumper_inputs(example, oracle):
    from greyjack_b200 import Problem
    doc, spec, samples, base, deltas = _load(example)
    op = oracle.OracleProblem(spec)
    gp = Problem(spec)
    got_plain, got_incr = gp.request_score_plain(samples), gp.request_score_incremental(base, deltas)
    assert np.array_equal(got_plain, op.score_plain(samples))
    assert np.array_equal(got_incr, op.score_incremental(base, deltas))
    path = os.path.join(OUT, f"{example}.json")
    if os.path.exists(path):
        with open(path) as f:
            ref = json.load(f)
        want_plain, want_incr = np.array(ref["plain"]), np.array(ref["incremental"])
        assert np.array_equal(got_incr, want_incr)
        L = spec.levels
        for l in range(L - 1 if L > 1 else 1):
            assert np.array_equal(got_plain[:, l], want_plain[:, l])
        if L > 1:
            np.testing.assert_allclose(got_plain[:, L - 1], want_plain[:, L - 1], rtol=1e-12, atol=0.0)
    gp.close()
