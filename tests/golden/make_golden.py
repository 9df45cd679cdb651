#!/usr/bin/env python
"""Generates the golden fixtures in tests/golden/*.npz.

The reference ships no golden vectors for its scorers and cannot be built
here (Rust + polars 0.46.0), so the fixtures are produced by a SECOND, independent restatement of
the reference's scoring arithmetic, written in plain Python / pandas straight from the reference
sources, and cross-checked against the C oracle before they are written:

  * ISC scorers: line-by-line Python of examples/{nqueens,tsp,vrp,vrp_service}/src/score/
    incremental_score_calculator.rs (HashSet sizes, route bucketing, sequential f64 folds --
    Python floats are IEEE f64, so the sums are bit-identical to Rust's when done in the same order);
  * PSC scorers: the Polars queries of examples/*/src/score/plain_score_calculator.rs restated
    with pandas (group_by / nunique, inner joins, 3-key sort, partition walks).

Run from the repo root:  python tests/golden/make_golden.py
Fixtures hold inputs AND expected outputs, so neither /root/reference nor this script is needed
when the tests run (CPU: oracle vs fixtures; GPU: CUDA vs fixtures)."""
import math
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "greyjack-solver-rust_b200", "python"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from greyjack_b200 import instances as inst  # noqa: E402


# ---- decoding: variables/gj_integer.rs:66-83, utils/math_utils.rs:6-8 -------------------------
def rint(x):
    f, c = math.floor(x), math.ceil(x)
    return f if abs(x - f) < abs(c - x) else c


def decode(spec, i, x):
    if spec.frozen is not None and spec.frozen[i]:
        return int(spec.initial[i])
    lb, ub = spec.lower_bounds[i], spec.upper_bounds[i]
    return int(rint(min(max(x, lb), ub)))


def decode_row(spec, row):
    return [decode(spec, i, float(x)) for i, x in enumerate(row)]


def apply_deltas(spec, base_dec, deltas):
    cand = list(base_dec)
    for var, val in deltas:
        cand[var] = decode(spec, var, float(val))
    return cand


# ---- ISC restatements ------------------------------------------------------------------------------
def nqueens_isc(spec, rows):
    n = len(rows)
    cols = list(range(n)) if spec.column_id is None else [int(c) for c in spec.column_id]
    u_rows = len(set(rows))
    u_desc = len(set(c + r for c, r in zip(cols, rows)))
    u_asc = len(set(c - r for c, r in zip(cols, rows)))
    return [float(n - u_rows) + float(n - u_desc) + float(n - u_asc)]


def tsp_isc(spec, stops):
    D = spec.distance_matrix
    n = len(stops)
    hard = float(n - len(set(stops)))
    dist = 0.0
    dist += D[0][stops[0]]
    dist += D[stops[n - 1]][0]
    fold = 0.0
    for i in range(1, n):
        fold = fold + D[stops[i - 1]][stops[i]]
    dist += fold
    return [hard, dist]


def vrp_isc(spec, flat, service_rule):
    D = spec.distance_matrix
    veh, cus = flat[0::2], flat[1::2]
    K = spec.n_vehicles
    n = len(cus)
    unique_pen = 1000.0 * float(n - len(set(cus)))
    demand = [0] * K
    for v, c in zip(veh, cus):
        demand[v] += int(spec.demand[c])
    cap_pen = sum(abs(int(spec.vehicle_capacity[v]) - demand[v]) for v in range(K)
                  if int(spec.vehicle_capacity[v]) - demand[v] < 0)
    routes = [[] for _ in range(K)]
    for v, c in zip(veh, cus):
        routes[v].append(c)
    dists, lates = [0.0] * K, [0.0] * K
    for v in range(K):
        r = routes[v]
        if not r:
            continue
        depot = int(spec.vehicle_depot[v])
        cur = 0.0
        cur += D[depot][r[0]]
        cur += D[r[-1]][depot]
        fold = 0.0
        for i in range(1, len(r)):
            fold = fold + D[r[i - 1]][r[i]]
        cur += fold
        dists[v] = cur
        if spec.time_windowed:
            t = int(spec.work_day_start[v])
            pen = 0.0
            for c in r:
                ws, we, sv = int(spec.tw_start[c]), int(spec.tw_end[c]), int(spec.service_time[c])
                t = max(t, ws)
                if service_rule:                       # vrp_service ISC :121-122
                    if t > we + sv:
                        pen += float(t - (we + sv))
                else:                                  # vrp ISC :120-121
                    if t + sv > we:
                        pen += float((t + sv) - we)
                t += sv
            if t > int(spec.work_day_end[v]):
                pen += float(t - int(spec.work_day_end[v]))
            lates[v] = pen
    sum_d = 0.0
    for d in dists:
        sum_d += d
    sum_l = 0.0
    for x in lates:
        sum_l += x
    return [unique_pen + float(cap_pen), sum_l, sum_d]


def isc(spec, cand):
    if spec.kind == inst.NQUEENS:
        return nqueens_isc(spec, cand)
    if spec.kind == inst.TSP:
        return tsp_isc(spec, cand)
    return vrp_isc(spec, cand, spec.kind == inst.VRP_SERVICE)


# ---- PSC restatements (pandas in place of polars) --------------------------------------------------
def nqueens_psc(spec, samples):
    n = spec.n_vars
    S = len(samples)
    cols = np.arange(n) if spec.column_id is None else np.asarray(spec.column_id)
    df = pd.DataFrame({"sample_id": np.repeat(np.arange(S), n), "row_id": np.concatenate(samples),
                       "column_id": np.tile(cols, S)})
    df["desc_id"] = df["column_id"] + df["row_id"]
    df["asc_id"] = df["column_id"] - df["row_id"]
    g = df.groupby("sample_id")
    out = (g["row_id"].size() - g["row_id"].nunique()) + (g["desc_id"].size() - g["desc_id"].nunique()) + \
          (g["asc_id"].size() - g["asc_id"].nunique())
    return [[float(v)] for v in out.sort_index().to_numpy()]


def tsp_psc(spec, samples):
    n = spec.n_vars
    S = len(samples)
    df = pd.DataFrame({"sample_id": np.repeat(np.arange(S), n), "location_vec_id": np.concatenate(samples)})
    g = df.groupby("sample_id")["location_vec_id"]
    hard = (g.size() - g.nunique()).sort_index().to_numpy().astype(float)
    out = []
    for s in range(S):                                   # sort by sample_id + partition_by(sample_id)
        stops = df[df["sample_id"] == s]["location_vec_id"].tolist()
        out.append([float(hard[s]), tsp_isc(spec, stops)[1]])
    return out


def vrp_psc(spec, samples):
    n = spec.n_vars // 2
    S = len(samples)
    K = spec.n_vehicles
    stops = pd.DataFrame({
        "sample_id": np.repeat(np.arange(S), n),
        "vehicle_id": np.concatenate([np.asarray(s)[0::2] for s in samples]),
        "customer_id": np.concatenate([np.asarray(s)[1::2] for s in samples])})
    vehicles = pd.DataFrame({"vehicle_id": np.arange(K), "depot_vec_id": np.asarray(spec.vehicle_depot),
                             "capacity": np.asarray(spec.vehicle_capacity).astype(np.int64),
                             "work_day_start": np.asarray(spec.work_day_start).astype(np.int64),
                             "work_day_end": np.asarray(spec.work_day_end).astype(np.int64)})
    L = spec.n_locations
    customers = pd.DataFrame({"customer_id": np.arange(L), "demand": np.asarray(spec.demand).astype(np.int64),
                              "time_window_start": np.asarray(spec.tw_start).astype(np.int64),
                              "time_window_end": np.asarray(spec.tw_end).astype(np.int64),
                              "service_time": np.asarray(spec.service_time).astype(np.int64)})
    stops["index"] = np.arange(len(stops))
    common = stops.merge(vehicles, on="vehicle_id", how="inner").merge(customers, on="customer_id", how="inner")
    common = common.sort_values(["sample_id", "vehicle_id", "index"], kind="stable")
    g = stops.groupby("sample_id")["customer_id"]
    dup = (g.size() - g.nunique()).sort_index().to_numpy().astype(float) * 1000.0
    trip = common.groupby(["sample_id", "vehicle_id"], as_index=False).agg(sum_trip_demand=("demand", "sum"))
    trip = trip.merge(vehicles, on="vehicle_id", how="inner")
    trip["diff"] = trip["capacity"] - trip["sum_trip_demand"]
    bad = trip[trip["diff"] < 0]
    cap = np.zeros(S)
    for sid, sub in bad.groupby("sample_id"):
        cap[sid] = float(sub["diff"].abs().sum())
    D = spec.distance_matrix
    out = []
    for s in range(S):
        sdf = common[common["sample_id"] == s]
        dist_total, late_total = 0.0, 0.0
        for v, vdf in sdf.groupby("vehicle_id", sort=True):
            ids = vdf["customer_id"].tolist()
            depot = int(vdf["depot_vec_id"].iloc[0])
            cur = 0.0
            cur += D[depot][ids[0]]
            cur += D[ids[-1]][depot]
            fold = 0.0
            for i in range(1, len(ids)):
                fold = fold + D[ids[i - 1]][ids[i]]
            cur += fold
            dist_total = dist_total + cur
            if spec.time_windowed:                         # PSC :207-216 walks 0..len-1
                t = int(vdf["work_day_start"].iloc[0])
                ws, we, sv = vdf["time_window_start"].tolist(), vdf["time_window_end"].tolist(), vdf["service_time"].tolist()
                pen = 0.0
                for i in range(len(ids) - 1):
                    t = max(t, ws[i])
                    if t > we[i] + sv[i]:
                        pen += float(t - (we[i] + sv[i]))
                    t += sv[i]
                if t > int(vdf["work_day_end"].iloc[0]):
                    pen += float(t - int(vdf["work_day_end"].iloc[0]))
                late_total = late_total + pen
        out.append([dup[s] + cap[s], late_total, dist_total])
    return out


def psc(spec, samples_dec):
    if spec.kind == inst.NQUEENS:
        return nqueens_psc(spec, samples_dec)
    if spec.kind == inst.TSP:
        return tsp_psc(spec, samples_dec)
    return vrp_psc(spec, samples_dec)


# ---- fixtures ------------------------------------------------------------------------------------------
CASES = {
    "nqueens16": lambda: inst.nqueens(16, seed=45),
    "tsp40": lambda: inst.tsp(40, seed=7),
    "cvrp24": lambda: inst.cvrp(24, 4, seed=2),
    "vrptw30": lambda: inst.vrptw(30, 4, n_depots=2, seed=3, service_variant=False),
    "vrpsvc30": lambda: inst.vrptw(30, 4, n_depots=2, seed=3, service_variant=True),
}


def main():
    from helpers import random_moves, random_samples
    from oracle import gj_oracle
    for name, mk in CASES.items():
        spec = mk()
        op = gj_oracle.OracleProblem(spec)
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        # plain: 24 wild candidates (fractional / out-of-range values exercise the decoder)
        samples = random_samples(spec, 24, rng, wild=True)
        dec = [decode_row(spec, row) for row in samples]
        plain = np.array(psc(spec, dec), dtype=np.float64)
        # incremental: 40 moves of every kind from the instance's start vector
        base = np.asarray(spec.initial, dtype=np.float64)
        deltas, kinds = random_moves(op, spec, base, 40, rng)
        base_dec = decode_row(spec, base)
        incr = np.array([isc(spec, apply_deltas(spec, base_dec, d)) for d in deltas], dtype=np.float64)
        # second opinion vs the C oracle: integer levels and (same summation order) floats bit-exact
        assert np.array_equal(op.score_plain(samples), plain), name
        assert np.array_equal(op.score_incremental(base, deltas), incr), name
        offs, ids, vals = gj_oracle.deltas_to_csr(deltas)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), samples=samples, plain=plain, base=base,
                            offsets=offs, var_ids=ids, values=vals, incremental=incr,
                            kinds=np.asarray(kinds, dtype=np.int32))
        print(f"{name}: {len(samples)} plain + {len(deltas)} incremental vectors, levels={spec.levels}")


if __name__ == "__main__":
    main()
