"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from an independent
Python / pandas restatement of the reference's ISC and PSC scorers): the C oracle must reproduce
them on CPU, the CUDA path on the GPU -- integer levels bit-exact, and the float level bit-exact
too because the default scoring mode keeps the reference's summation order."""
import os

import numpy as np
import pytest

from greyjack_b200 import instances as inst

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    "nqueens16": lambda: inst.nqueens(16, seed=45),
    "tsp40": lambda: inst.tsp(40, seed=7),
    "cvrp24": lambda: inst.cvrp(24, 4, seed=2),
    "vrptw30": lambda: inst.vrptw(30, 4, n_depots=2, seed=3, service_variant=False),
    "vrpsvc30": lambda: inst.vrptw(30, 4, n_depots=2, seed=3, service_variant=True),
}


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_golden(name, oracle):
    spec, g = CASES[name](), _load(name)
    op = oracle.OracleProblem(spec)
    assert np.array_equal(op.score_plain(g["samples"]), g["plain"])
    got = op.score_incremental_csr(g["base"], g["offsets"], g["var_ids"], g["values"])
    assert np.array_equal(got, g["incremental"])
    assert set(g["kinds"].tolist()) == {0, 1, 2, 3, 4, 5}          # every move of mover.rs is covered


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_reproduces_golden(name):
    from greyjack_b200 import Problem
    spec, g = CASES[name](), _load(name)
    gp = Problem(spec)
    assert np.array_equal(gp.request_score_plain(g["samples"]), g["plain"])
    got = gp.request_score_incremental_csr(g["base"], g["offsets"], g["var_ids"], g["values"])
    assert np.array_equal(got, g["incremental"])
    gp.close()


def test_known_answers_of_the_reference_docs():
    """Closed forms verified by hand while surveying against the fixtures'
    generator functions, so the second opinion itself is anchored."""
    import importlib.util
    spec_ = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mg)
    s = inst.nqueens(8)
    assert mg.nqueens_isc(s, [0, 4, 7, 5, 2, 6, 1, 3]) == [0.0]          # an 8-queens solution
    assert mg.nqueens_isc(s, list(range(8))) == [7.0]                     # main diagonal: N-1
    assert mg.nqueens_isc(s, [3] * 8) == [7.0]                            # one row: N-1
    assert [mg.rint(x) for x in (4.4, 4.6, 4.5, 2.5)] == [4, 5, 5, 3]     # gj_integer.rs:157-181
