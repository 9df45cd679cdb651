"""TSPLIB reader (host side): examples/tsp/src/persistence/domain_builder.rs:90-211."""
import numpy as np
import pytest

from greyjack_b200 import instances as inst

SQUARE = "\n".join([
    "NAME : square4",
    "COMMENT : unit test",
    "TYPE : TSP",
    "DIMENSION : 4",
    "EDGE_WEIGHT_TYPE : EUC_2D",
    "NODE_COORD_SECTION",
    "1 0.0 0.0",
    "2   3.0   0.0",
    "3 3.0 4.0  third",
    "4 0.0 4.0",
    "EOF",
    "",
])


def test_read_euc_2d():
    meta, xy, m = inst.read_tsplib(SQUARE)
    assert meta == {"dataset_name": "square4", "distance_type": "EUC_2D"}
    assert m is None and xy.shape == (4, 2) and np.array_equal(xy[2], [3.0, 4.0])


def test_spec_scores_the_perimeter(oracle):
    spec = inst.tsp_from_tsplib(SQUARE)
    assert spec.n_vars == 3 and spec.n_locations == 4 and spec.name == "square4"
    assert spec.distance_matrix[0, 2] == 5.0
    # nearest neighbour from the depot: 1 (3.0), then 2 (4.0), then 3 (3.0)
    assert np.array_equal(spec.initial, [1.0, 2.0, 3.0])
    sc = oracle.OracleProblem(spec).score_plain(spec.initial[None, :])[0]
    assert np.array_equal(sc, [0.0, 14.0])


def test_explicit_matrix_and_errors():
    text = ("NAME : m3\nEDGE_WEIGHT_TYPE : EXPLICIT\nNODE_COORD_SECTION\n1 0 0\n2 1 0\n3 0 1\nEOF\n"
            "0 2 9 \n2 0 4 \n9 4 0 \nEOF\n")
    spec = inst.tsp_from_tsplib(text, greedy=False)
    assert np.array_equal(spec.distance_matrix, [[0, 2, 9], [2, 0, 4], [9, 4, 0]])
    with pytest.raises(ValueError):
        inst.read_tsplib("NAME : x\nEDGE_WEIGHT_TYPE : EUC_2D\n")
    with pytest.raises(ValueError):
        inst.tsp_from_tsplib("NAME : x\nEDGE_WEIGHT_TYPE : EUC_2D\nNODE_COORD_SECTION\n1 0 0\nEOF\n")
