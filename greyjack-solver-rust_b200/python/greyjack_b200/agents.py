"""Agent builders with the reference's constructor argument lists, and the device-resident
island group they build.

    TabuSearch::new       greyjack/src/agents/tabu_search.rs:32-40
    LateAcceptance::new   greyjack/src/agents/late_acceptance.rs:31-38
    GeneticAlgorithm::new greyjack/src/agents/genetic_algorithm.rs:34-44

Termination strategies (agents/termination_strategies/*) stay host-side objects."""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _lib
from .problem import Problem, _ptr

TS, LA, GA, SA = 0, 1, 2, 3
SCORING_FULL, SCORING_DELTA, SCORING_DELTA_UNFUSED, SCORING_DELTA_F64 = 0, 1, 2, 3     # GJ_SCORING_* (include/greyjack_b200.h)


# ---- termination strategies (TerminationStrategiesVariants::{StL, TSL, SNI, ScL}) ------------
@dataclass
class StepsLimit:
    """steps_limit.rs"""
    steps_count_limit: int
    steps_made: int = 0

    def update(self, best_score=None, steps=1):
        self.steps_made += steps

    def is_accomplish(self):
        return self.steps_made > self.steps_count_limit

    def get_accomplish_rate(self):
        return self.steps_made / max(1, self.steps_count_limit)


@dataclass
class TimeSpentLimit:
    """time_spent_limit.rs (milliseconds)"""
    time_seconds_limit_ms: int
    start: Optional[float] = None

    def update(self, best_score=None, steps=1):
        if self.start is None:
            self.start = time.time()

    def is_accomplish(self):
        return self.start is not None and (time.time() - self.start) * 1000.0 >= self.time_seconds_limit_ms

    def get_accomplish_rate(self):
        return 0.0 if self.start is None else (time.time() - self.start) * 1000.0 / self.time_seconds_limit_ms


@dataclass
class ScoreNoImprovement:
    """score_no_improvement.rs (milliseconds without a better agent_top score)"""
    time_seconds_limit_ms: int
    last: Optional[tuple] = None
    since: Optional[float] = None

    def update(self, best_score=None, steps=1):
        now = time.time()
        key = None if best_score is None else tuple(best_score)
        if self.last is None or (key is not None and key < self.last):
            self.last, self.since = key, now

    def is_accomplish(self):
        return self.since is not None and (time.time() - self.since) * 1000.0 >= self.time_seconds_limit_ms

    def get_accomplish_rate(self):
        return 0.0 if self.since is None else (time.time() - self.since) * 1000.0 / self.time_seconds_limit_ms


@dataclass
class ScoreLimit:
    """score_limit.rs: stop once agent_top <= target"""
    target: tuple
    reached: bool = False

    def update(self, best_score=None, steps=1):
        if best_score is not None and tuple(best_score) <= tuple(self.target):
            self.reached = True

    def is_accomplish(self):
        return self.reached

    def get_accomplish_rate(self):
        return 1.0 if self.reached else 0.0


# ---- builders -----------------------------------------------------------------------------------
class _Builder:
    agent = TS

    def _params(self, n_islands, seed) -> _lib.AgentParams:
        p = _lib.AgentParams()
        p.agent = self.agent
        p.n_islands = n_islands
        p.seed = seed
        p.tabu_entity_rate = float(self.tabu_entity_rate)
        p.has_mutation_rate_multiplier = int(self.mutation_rate_multiplier is not None)
        p.mutation_rate_multiplier = float(self.mutation_rate_multiplier or 0.0)
        p.has_move_probas = int(self.move_probas is not None)
        if self.move_probas is not None:
            assert len(self.move_probas) == 6, "Optional move probas vector length is not equal to available moves count"
            for i, v in enumerate(self.move_probas):
                p.move_probas[i] = float(v)
        p.migration_frequency = int(self.migration_frequency)
        p.reference_noop_moves = int(getattr(self, "reference_noop_moves", True))
        mode = getattr(self, "scoring", "full")
        p.scoring_mode = {"full": SCORING_FULL, "delta": SCORING_DELTA, "delta_unfused": SCORING_DELTA_UNFUSED,
                          "delta_f64": SCORING_DELTA_F64}[mode]
        p.chain_steps_per_launch = int(getattr(self, "chain_steps_per_launch", 0))
        return p

    def build_agent(self, problem: Problem, n_islands: int = 1, seed: int = 0, initial=None) -> "Islands":
        return Islands(problem, self, n_islands=n_islands, seed=seed, initial=initial)


class TabuSearch(_Builder):
    agent = TS

    def __init__(self, neighbours_count, tabu_entity_rate, compare_to_global, mutation_rate_multiplier,
                 move_probas, migration_frequency, termination_strategy=None, reference_noop_moves=True,
                 scoring="full"):
        self.scoring = scoring
        self.neighbours_count = neighbours_count
        self.tabu_entity_rate = tabu_entity_rate
        self.compare_to_global = compare_to_global
        self.mutation_rate_multiplier = mutation_rate_multiplier
        self.move_probas = move_probas
        self.migration_frequency = migration_frequency
        self.termination_strategy = termination_strategy
        self.reference_noop_moves = reference_noop_moves

    def _params(self, n_islands, seed):
        p = super()._params(n_islands, seed)
        p.neighbours_count = int(self.neighbours_count)
        p.compare_to_global = int(bool(self.compare_to_global))
        return p


class LateAcceptance(_Builder):
    agent = LA

    def __init__(self, late_acceptance_size, tabu_entity_rate, mutation_rate_multiplier, move_probas,
                 migration_frequency, termination_strategy=None, reference_noop_moves=True, scoring="full",
                 chain_steps_per_launch=0):
        self.scoring = scoring
        self.chain_steps_per_launch = chain_steps_per_launch
        self.late_acceptance_size = late_acceptance_size
        self.tabu_entity_rate = tabu_entity_rate
        self.mutation_rate_multiplier = mutation_rate_multiplier
        self.move_probas = move_probas
        self.migration_frequency = migration_frequency
        self.termination_strategy = termination_strategy
        self.reference_noop_moves = reference_noop_moves

    def _params(self, n_islands, seed):
        p = super()._params(n_islands, seed)
        p.late_acceptance_size = int(self.late_acceptance_size)
        return p


class SimulatedAnnealing(_Builder):
    """agents/simulated_annealing.rs:31-39"""
    agent = SA

    def __init__(self, initial_temperature, cooling_rate, tabu_entity_rate, mutation_rate_multiplier, move_probas,
                 migration_frequency, termination_strategy=None, reference_noop_moves=True, scoring="delta",
                 chain_steps_per_launch=0):
        self.initial_temperature = list(initial_temperature)
        self.cooling_rate = cooling_rate
        self.tabu_entity_rate = tabu_entity_rate
        self.mutation_rate_multiplier = mutation_rate_multiplier
        self.move_probas = move_probas
        self.migration_frequency = migration_frequency
        self.termination_strategy = termination_strategy
        self.reference_noop_moves = reference_noop_moves
        self.scoring = scoring
        self.chain_steps_per_launch = chain_steps_per_launch

    def _params(self, n_islands, seed):
        p = super()._params(n_islands, seed)
        p.has_cooling_rate = int(self.cooling_rate is not None)
        p.cooling_rate = float(self.cooling_rate or 0.0)
        for i in range(3):
            p.initial_temperature[i] = float(self.initial_temperature[i]) if i < len(self.initial_temperature) else 1.0
        return p


class GeneticAlgorithm(_Builder):
    agent = GA

    def __init__(self, population_size, crossover_probability, p_best_rate, tabu_entity_rate,
                 mutation_rate_multiplier, move_probas, migration_rate, migration_frequency,
                 termination_strategy=None):
        self.population_size = population_size
        self.crossover_probability = crossover_probability
        self.p_best_rate = p_best_rate
        self.tabu_entity_rate = tabu_entity_rate
        self.mutation_rate_multiplier = mutation_rate_multiplier
        self.move_probas = move_probas
        self.migration_rate = migration_rate
        self.migration_frequency = migration_frequency
        self.termination_strategy = termination_strategy

    def _params(self, n_islands, seed):
        p = super()._params(n_islands, seed)
        p.population_size = int(self.population_size)
        p.crossover_probability = float(self.crossover_probability)
        p.p_best_rate = float(self.p_best_rate)
        p.migration_rate = float(self.migration_rate)
        return p


# ---- island group ---------------------------------------------------------------------------------
class Islands:
    """A group of agents of one kind resident on one GPU (gj_islands handle)."""

    def __init__(self, problem: Problem, builder: _Builder, n_islands=1, seed=0, initial=None):
        self._L = _lib.load()
        self.problem = problem
        self.builder = builder
        self.n_islands = n_islands
        self.params = builder._params(n_islands, seed)
        init = None
        if initial is not None:
            init = np.ascontiguousarray(initial, dtype=np.float64).reshape(n_islands, problem.n_vars)
        h = C.c_void_p()
        _lib.check(self._L.gj_islands_create(problem.handle, C.byref(self.params), _ptr(init), C.byref(h)))
        self.handle = h
        self.K = int(builder.neighbours_count) if builder.agent == TS else (
            1 if builder.agent in (LA, SA) else 2 * ((int(builder.population_size) + 1) // 2))

    def close(self):
        if getattr(self, "handle", None):
            self._L.gj_islands_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, n_steps=1, stream=0):
        _lib.check(self._L.gj_islands_step(self.handle, C.c_int64(n_steps), C.c_void_p(stream)))

    def stats(self):
        c, s, a = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self._L.gj_islands_stats(self.handle, C.byref(c), C.byref(s), C.byref(a)))
        return {"candidates": c.value, "steps": s.value, "accepted": a.value}

    def set_accomplish_rate(self, rate: float):
        """termination_strategy.get_accomplish_rate() for SimulatedAnnealing without a cooling rate."""
        _lib.check(self._L.gj_islands_set_accomplish_rate(self.handle, C.c_double(rate)))

    def trace_aux(self, island=0):
        out = np.zeros(5, dtype=np.float64)
        _lib.check(self._L.gj_islands_trace_aux(self.handle, C.c_int32(island), _ptr(out)))
        return {"random": out[0], "accept_proba": out[1], "temperature": out[2:5].copy()}

    @property
    def step_path(self) -> str:
        """Kernel path of a step: fused / fused_lean / chain / vrp_chain / delta / vrp_delta / full / ga."""
        return self._L.gj_islands_step_path(self.handle).decode()

    def set_profiling(self, on: bool):
        _lib.check(self._L.gj_islands_set_profiling(self.handle, C.c_int32(int(on))))

    def profile_read(self):
        """(summed ms of the scoring kernel, launches) since profiling was switched on."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        _lib.check(self._L.gj_islands_profile_read(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def _ind(self, fn, island):
        v = np.empty(self.problem.n_vars, dtype=np.float64)
        s = np.empty(self.problem.levels, dtype=np.float64)
        _lib.check(fn(self.handle, C.c_int32(island), _ptr(v), _ptr(s)))
        return v, s

    def best(self, island=-1):
        return self._ind(self._L.gj_islands_best, island)

    def current(self, island=0):
        return self._ind(self._L.gj_islands_current, island)

    def set_external_ring(self, on: bool, island_base: int = 0):
        _lib.check(self._L.gj_islands_set_external_ring(self.handle, C.c_int32(int(on)), C.c_int32(island_base)))

    def migrant_bytes(self) -> int:
        return int(self._L.gj_islands_migrant_bytes(self.handle))

    def export_migrants(self, d_buffer: int, stream=0):
        _lib.check(self._L.gj_islands_export_migrants(self.handle, C.c_void_p(d_buffer), C.c_void_p(stream)))

    def import_migrants(self, d_buffer: int, stream=0):
        _lib.check(self._L.gj_islands_import_migrants(self.handle, C.c_void_p(d_buffer), C.c_void_p(stream)))

    def ga_population(self, island=0):
        """(rows [pop][n_vars], scores [pop][levels], order [pop]) of a GeneticAlgorithm island."""
        pop, n, lv = int(self.builder.population_size), self.problem.n_vars, self.problem.levels
        rows = np.zeros((pop, n)); scores = np.zeros((pop, lv)); order = np.zeros(pop, dtype=np.int32)
        _lib.check(self._L.gj_islands_ga_population(self.handle, C.c_int32(island), _ptr(rows), _ptr(scores), _ptr(order)))
        return rows, scores, order

    def ga_trace_generation(self, island=0):
        """One GA generation of every island, the decisions of `island` exposed (gj_ga_trace)."""
        pop, n, lv = int(self.builder.population_size), self.problem.n_vars, self.problem.levels
        half = (pop + 1) // 2
        nc = 2 * half
        out = {"order_before": np.zeros(pop, dtype=np.int32), "pairs": np.zeros((half, 8)),
               "desc": np.zeros((nc, 20), dtype=np.int32), "cand_rows": np.zeros((nc, n)),
               "cand_scores": np.zeros((nc, lv)), "replace": np.zeros((pop, 3)), "src": np.zeros(pop, dtype=np.int32)}
        t = _lib.GaTrace()
        t.order_before = out["order_before"].ctypes.data; t.pairs = out["pairs"].ctypes.data
        t.move_desc = out["desc"].ctypes.data; t.cand_rows = out["cand_rows"].ctypes.data
        t.cand_scores = out["cand_scores"].ctypes.data; t.replace = out["replace"].ctypes.data
        t.src = out["src"].ctypes.data
        _lib.check(self._L.gj_islands_ga_trace_generation(self.handle, C.c_int32(island), C.byref(t)))
        return out

    def global_top_bytes(self) -> int:
        return int(self._L.gj_islands_global_top_bytes(self.handle))

    def export_global_top(self, d_buffer: int, stream=0):
        _lib.check(self._L.gj_islands_export_global_top(self.handle, C.c_void_p(d_buffer), C.c_void_p(stream)))

    def import_global_top(self, d_records: int, count: int, stream=0):
        _lib.check(self._L.gj_islands_import_global_top(self.handle, C.c_void_p(d_records), C.c_int32(count), C.c_void_p(stream)))

    def trace_tabu(self, island=0, group=0):
        """The tabu deque of one island / semantic group, newest id first -> (ids, size)."""
        cap = max(1, self.problem.n_vars)
        ids = np.zeros(cap, dtype=np.int32)
        fill, size = C.c_int32(0), C.c_int32(0)
        _lib.check(self._L.gj_islands_trace_tabu(self.handle, C.c_int32(island), C.c_int32(group), _ptr(ids),
                                                 C.c_int32(cap), C.byref(fill), C.byref(size)))
        return ids[:fill.value].copy(), size.value

    def trace_step(self, island=0):
        """One TS/LA step with everything exposed (see gj_islands_trace_step)."""
        K, n, lv = self.K, self.problem.n_vars, self.problem.levels
        cap = K * (n + 16)
        offs = np.zeros(K + 1, dtype=np.uint64)
        ids = np.zeros(cap, dtype=np.uint64)
        vals = np.zeros(cap, dtype=np.float64)
        kinds = np.zeros(K, dtype=np.int32)
        desc = np.zeros((K, 20), dtype=np.int32)
        scores = np.zeros((K, lv), dtype=np.float64)
        sel, acc = C.c_int64(-1), C.c_int32(0)
        _lib.check(self._L.gj_islands_trace_step(
            self.handle, C.c_int32(island), _ptr(offs), _ptr(ids), _ptr(vals), C.c_int64(cap),
            _ptr(kinds), _ptr(desc), _ptr(scores), C.byref(sel), C.byref(acc)))
        deltas = [[(int(ids[k]), float(vals[k])) for k in range(int(offs[j]), int(offs[j + 1]))]
                  for j in range(K)]
        return {"deltas": deltas, "kinds": kinds, "desc": desc, "scores": scores,
                "selected": sel.value, "accepted": bool(acc.value)}
