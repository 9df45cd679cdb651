"""greyjack_b200 -- Python harness over the C-ABI CUDA engine (libgreyjack_b200.so).

The product is the shared library (greyjack-solver-rust_b200/csrc, include/greyjack_b200.h);
this package only mirrors the reference's host-side names for tests, bench.py and demos."""
from . import instances  # noqa: F401

from ._lib import GjError, LIB_PATH, load  # noqa: F401
from .problem import PinnedArray, Problem, deltas_to_csr, pinned_copy  # noqa: F401
from .agents import (GeneticAlgorithm, Islands, LateAcceptance, ScoreLimit, ScoreNoImprovement,  # noqa: F401
                     SimulatedAnnealing, StepsLimit, TabuSearch, TimeSpentLimit)
from .solver import Solver  # noqa: F401,E402
from . import wire  # noqa: F401,E402
from .program import ConstraintProgram, TermSpec, nqueens_program, tsp_program, vrp_program  # noqa: F401,E402
