#!/usr/bin/env python
"""Writes the inputs of the reference-side dumper (oracle/reference_dump/README.md): instance files the
reference's own DomainBuilders read (tests/golden/reference_inputs/*.tsp / *.vrp / *.json) and, per
example, a JSON with candidate samples, a base + delta lists and a candidate for the mover.  The
reference (run by anyone with cargo) turns each <example>.json into reference_outputs/<example>.json;
tests/test_reference_dump.py compares the oracle and the CUDA path with those outputs.

usage: python tests/golden/make_reference_inputs.py      (idempotent: fixed seeds)"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "greyjack-solver-rust_b200", "python"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from greyjack_b200 import instances as inst  # noqa: E402
from helpers import permutation_samples, random_moves, random_samples  # noqa: E402
from oracle import gj_oracle  # noqa: E402

OUT = os.path.join(HERE, "reference_inputs")
REL = "tests/golden/reference_inputs"


def specs():
    """(example, ProblemSpec, instance entry of the JSON, files to write)"""
    nq = inst.nqueens(24, seed=45)
    yield "nqueens", nq, {"n_queens": 24, "seed": 45}, {}
    t = inst.tsp(40, seed=7, greedy=False)
    text = inst.write_tsplib(t, "tsp40")
    yield "tsp", inst.tsp_from_tsplib(text, greedy=False), {"path": f"{REL}/tsp40.tsp"}, {"tsp40.tsp": text}
    v = inst.vrptw(30, 4, n_depots=2, seed=11, service_variant=False, greedy=False)
    text = inst.write_vrp(v, "vrptw30-k4")
    yield "vrp", inst.vrp_from_file(text, greedy=False), {"path": f"{REL}/vrptw30-k4.vrp"}, {"vrptw30-k4.vrp": text}
    s = inst.vrptw(30, 4, n_depots=2, seed=12, service_variant=True, greedy=False)
    doc = inst.write_vrp_service_json(s, "vrpsvc30")
    yield "vrp_service", inst.vrp_service_from_json(doc, greedy=False), {"path": f"{REL}/vrpsvc30.json"}, \
        {"vrpsvc30.json": json.dumps(doc)}


def build(example, spec, instance):
    rng = np.random.default_rng({"nqueens": 1, "tsp": 2, "vrp": 3, "vrp_service": 4}[example])
    op = gj_oracle.OracleProblem(spec)
    samples = np.concatenate([random_samples(spec, 24, rng), permutation_samples(spec, 24, rng)])
    base = permutation_samples(spec, 1, rng)[0]
    deltas, kinds = random_moves(op, spec, base, 96, rng)
    return {"example": example, "instance": instance,
            "samples": samples.tolist(), "base": base.tolist(),
            "deltas": [[[int(c), float(v)] for c, v in d] for d in deltas], "delta_kinds": [int(k) for k in kinds],
            "mover": {"candidate": base.tolist(), "n_moves": 16, "tabu_entity_rate": 0.0}}


def main():
    os.makedirs(OUT, exist_ok=True)
    for example, spec, instance, files in specs():
        for name, text in files.items():
            with open(os.path.join(OUT, name), "w") as f:
                f.write(text)
        with open(os.path.join(OUT, f"{example}.json"), "w") as f:
            json.dump(build(example, spec, instance), f)
        print("wrote", example)


if __name__ == "__main__":
    main()
