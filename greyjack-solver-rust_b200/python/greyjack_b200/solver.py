"""Host loop around device-resident islands: the part of Solver::solve (solver/solver.rs:25-143)
and Agent::solve (agents/base/agent_base.rs:124-188) that is NOT data-parallel -- termination
strategies, observers, logging, and (for SimulatedAnnealing without a cooling rate) the accomplish
rate that drives the temperature (agent_base.rs:537-552).  Everything else of a step runs inside
gj_islands_step on the GPU.

    best_vars, best_score = Solver.solve(problem, TabuSearch(...), n_jobs=148,
                                         termination_strategy=StepsLimit(1000), observers=[obs])
"""
from __future__ import annotations

import copy
import time
from typing import Callable, Iterable, Optional

import numpy as np

from .agents import SA, Islands, _Builder
from .problem import Problem

# SolverLoggingLevels (solver/solver.rs)
SILENT, INFO, TRACE = 0, 1, 2


def _is_number_vector(x) -> bool:
    return not isinstance(x, str) and len(x) > 0 and all(isinstance(v, (int, float, np.floating, np.integer)) for v in x)


class Solver:
    @staticmethod
    def solve(problem: Problem, agent_builder: _Builder, n_jobs: int = 1,
              termination_strategy=None, observers: Optional[Iterable] = None,
              logging_level: int = SILENT, initial_solution=None, seed: int = 0,
              steps_per_call: Optional[int] = None, log: Callable[[str], None] = print):
        """n_jobs = islands on this GPU (the reference: one agent per rayon worker, solver.rs:58-64).
        observers: objects with update(dict) (ObserverTrait::update), called whenever the global best
        improves with {"score": [...], "variable_values": [...], "step": n}.
        initial_solution: a variable vector, or the reference's solution Value / JSON text
        (InitialSolutionVariants::CotwinValuesVector, solver.rs:108-119; see wire.py) -- every island
        starts from it.
        Returns (variable_values, score) of the global best individual; wire.solution_to_value turns
        that into the Value the reference's Solver::solve returns."""
        term = copy.deepcopy(termination_strategy if termination_strategy is not None
                             else getattr(agent_builder, "termination_strategy", None))
        if term is None:
            raise ValueError("a termination strategy is required")
        init = None
        if initial_solution is not None:
            if isinstance(initial_solution, (str, list, tuple)) and not _is_number_vector(initial_solution):
                from . import wire
                initial_solution, _ = wire.solution_from_value(problem.spec, initial_solution)
            init = np.tile(np.asarray(initial_solution, dtype=np.float64), (n_jobs, 1))
        islands = Islands(problem, agent_builder, n_islands=n_jobs, seed=seed, initial=init)
        chunk = int(steps_per_call or max(1, int(getattr(agent_builder, "migration_frequency", 1))))
        observers = list(observers or [])
        best_seen = None
        steps = 0
        t0 = time.time()
        try:
            while True:
                if agent_builder.agent == SA and agent_builder.cooling_rate is None:
                    islands.set_accomplish_rate(min(1.0, max(0.0, term.get_accomplish_rate())))
                islands.step(chunk)
                steps += chunk
                vars_, score = islands.best(-1)
                term.update(score, steps=chunk)
                key = tuple(score)
                if best_seen is None or key < best_seen:
                    best_seen = key
                    for ob in observers:
                        ob.update({"score": list(score), "variable_values": vars_.tolist(), "step": steps})
                    if logging_level >= INFO:
                        log(f"[gj] step {steps}  best {list(score)}  {time.time() - t0:.2f}s")
                elif logging_level >= TRACE:
                    log(f"[gj] step {steps}  best {list(score)}")
                if term.is_accomplish():
                    break
            return islands.best(-1)
        finally:
            islands.close()
