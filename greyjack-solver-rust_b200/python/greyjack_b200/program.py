"""Constraint programs: the reference's constraint registry (PlainScoreCalculator::{new, add_constraint,
remove_constraint, set_constraint_weights, get_score}, plain_score_calculator.rs:20-94) with the
constraints written as data -- terms of four relational primitives -- instead of Polars closures, so
that they can run on the device (csrc/gj_program.cu).  The examples' constraints are provided as
ready-made programs: nqueens_program, tsp_program, vrp_program."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List

import numpy as np

from . import _lib
from . import instances as inst
from .problem import Problem, _ptr

DISTINCT_DEFICIT, GATHER_FOLD, SEGMENT_OVER_CAP, MAXPLUS_LATENESS = 0, 1, 2, 3
TW_ISC_FILE, TW_ISC_SERVICE, TW_PSC = 0, 1, 2


@dataclass
class TermSpec:
    op: int
    value_offset: int = 0
    value_stride: int = 1
    seg_offset: int = 0
    seg_stride: int = 0
    key_value_coef: int = 1
    key_index_coef: int = 0
    variant: int = 0
    scale: float = 1.0


class ConstraintProgram:
    """PlainScoreCalculator for constraints expressed as terms."""

    def __init__(self, problem: Problem, levels: int = None):
        self._L = _lib.load()
        self.problem = problem
        self.levels = int(levels if levels is not None else problem.levels)
        h = C.c_void_p()
        _lib.check(self._L.gj_program_create(problem.handle, C.c_int32(self.levels), C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self._L.gj_program_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_constraint(self, name: str, level: int, terms: List[TermSpec]):
        c = _lib.Constraint()
        c.name = name.encode()
        c.level = level
        c.n_terms = len(terms)
        for i, t in enumerate(terms):
            for f, _ in _lib.Term._fields_:
                setattr(c.terms[i], f, getattr(t, f))
        _lib.check(self._L.gj_program_add_constraint(self.handle, C.byref(c)))

    def remove_constraint(self, name: str):
        _lib.check(self._L.gj_program_remove_constraint(self.handle, name.encode()))

    def set_constraint_weights(self, weights: Dict[str, float]):
        names = (C.c_char_p * len(weights))(*[k.encode() for k in weights])
        vals = (C.c_double * len(weights))(*[float(v) for v in weights.values()])
        _lib.check(self._L.gj_program_set_constraint_weights(self.handle, names, vals, C.c_int32(len(weights))))

    @property
    def n_constraints(self) -> int:
        return int(self._L.gj_program_n_constraints(self.handle))

    def get_score(self, samples) -> np.ndarray:
        x = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, self.problem.n_vars)
        out = np.empty((x.shape[0], self.levels), dtype=np.float64)
        _lib.check(self._L.gj_program_get_score(self.handle, _ptr(x), C.c_int64(x.shape[0]), _ptr(out)))
        return out


def nqueens_program(problem: Problem) -> ConstraintProgram:
    """examples/nqueens/src/score/plain_score_calculator.rs:37-59: all_different =
    (rows - n_unique(row)) + (rows - n_unique(column + row)) + (rows - n_unique(column - row))"""
    g = ConstraintProgram(problem, 1)
    g.add_constraint("all_different", 0, [
        TermSpec(DISTINCT_DEFICIT, key_value_coef=1, key_index_coef=0),
        TermSpec(DISTINCT_DEFICIT, key_value_coef=1, key_index_coef=1),
        TermSpec(DISTINCT_DEFICIT, key_value_coef=-1, key_index_coef=1)])
    return g


def tsp_program(problem: Problem) -> ConstraintProgram:
    """examples/tsp/src/score/plain_score_calculator.rs:34-43 (no_duplicating_stops, hard) and :70-84
    (minimize_distance, soft)"""
    g = ConstraintProgram(problem, 2)
    g.add_constraint("no_duplicating_stops", 0, [TermSpec(DISTINCT_DEFICIT)])
    g.add_constraint("minimize_distance", 1, [TermSpec(GATHER_FOLD)])
    return g


def vrp_program(problem: Problem) -> ConstraintProgram:
    """examples/vrp/src/score/plain_score_calculator.rs: no_duplicating_stops (:51-68, x1000, hard),
    capacity (:95-107, hard), minimize_distance (:142-167, soft), late_arrival_penalty (:191-230, medium;
    only registered for time-windowed plans)"""
    g = ConstraintProgram(problem, 3)
    seg = dict(value_offset=1, value_stride=2, seg_offset=0, seg_stride=2)
    g.add_constraint("no_duplicating_stops", 0, [TermSpec(DISTINCT_DEFICIT, scale=1000.0, **seg)])
    g.add_constraint("capacity", 0, [TermSpec(SEGMENT_OVER_CAP, **seg)])
    g.add_constraint("minimize_distance", 2, [TermSpec(GATHER_FOLD, **seg)])
    if problem.spec.time_windowed:
        g.add_constraint("late_arrival_penalty", 1, [TermSpec(MAXPLUS_LATENESS, variant=TW_PSC, **seg)])
    return g
