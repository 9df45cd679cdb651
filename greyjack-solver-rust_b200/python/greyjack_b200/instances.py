"""Synthetic problem instances and the plain-numpy problem
description (`ProblemSpec`) that the C-ABI binding and the oracle binding both consume.

The reference ships no data (examples/data is git-ignored), so every instance here is
generated: coordinates U[0,1000)^2 from SplitMix64(seed), distance matrix = the examples'
`round(sqrt(dx^2+dy^2), 3)` truncation (examples/tsp/src/domain/location.rs:38-50),
variable bounds / semantic groups exactly as the examples' cotwin builders declare them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

NQUEENS, TSP, VRP, VRP_SERVICE = 0, 1, 2, 3
KIND_NAMES = {NQUEENS: "nqueens", TSP: "tsp", VRP: "vrp", VRP_SERVICE: "vrp_service"}
LEVELS = {NQUEENS: 1, TSP: 2, VRP: 3, VRP_SERVICE: 3}

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n outputs of SplitMix64 started at `seed` (vectorised, wraps mod 2^64)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, n: int) -> np.ndarray:
    return (splitmix64(seed, n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def randint(seed: int, n: int, lo: int, hi_inclusive: int) -> np.ndarray:
    span = np.uint64(hi_inclusive - lo + 1)
    return (splitmix64(seed, n) % span).astype(np.int64) + lo


def round3(d: np.ndarray) -> np.ndarray:
    """greyjack/src/utils/math_utils.rs:10-13 with precision 3 (truncation toward -inf)."""
    fl = np.floor(d)
    return fl + np.floor((d - fl) * 1000.0) / 1000.0


def distance_matrix(xy: np.ndarray) -> np.ndarray:
    """examples/tsp/src/domain/location.rs:38-50; xy is [n,2] (latitude, longitude)."""
    dlat = xy[None, :, 0] - xy[:, None, 0]
    dlon = xy[None, :, 1] - xy[:, None, 1]
    a = dlat * dlat
    b = dlon * dlon
    return round3(round3(np.sqrt(a + b)))


@dataclass
class ProblemSpec:
    kind: int
    n_vars: int
    lower_bounds: np.ndarray
    upper_bounds: np.ndarray
    frozen: Optional[np.ndarray] = None          # uint8 [n_vars]
    initial: Optional[np.ndarray] = None         # f64 [n_vars]; NaN = None
    groups: Dict[str, np.ndarray] = field(default_factory=dict)  # semantic group -> var ids
    column_id: Optional[np.ndarray] = None       # int64 [n_vars] (nqueens)
    n_locations: int = 0
    distance_matrix: Optional[np.ndarray] = None  # f64 [L,L]
    coords: Optional[np.ndarray] = None          # f64 [L,2]
    n_depots: int = 0
    n_vehicles: int = 0
    vehicle_depot: Optional[np.ndarray] = None   # int64 [K]
    vehicle_capacity: Optional[np.ndarray] = None  # uint64 [K]
    work_day_start: Optional[np.ndarray] = None
    work_day_end: Optional[np.ndarray] = None
    demand: Optional[np.ndarray] = None          # uint64 [L]
    tw_start: Optional[np.ndarray] = None
    tw_end: Optional[np.ndarray] = None
    service_time: Optional[np.ndarray] = None
    time_windowed: bool = False
    weights: np.ndarray = field(default_factory=lambda: np.ones(4, dtype=np.float64))
    score_precision: Optional[list] = None       # Solver::solve arg 5
    name: str = ""

    @property
    def levels(self) -> int:
        return LEVELS[self.kind]


def nqueens(n: int, seed: int = 45) -> ProblemSpec:
    """examples/nqueens: row_id[i] in [0, n-1], column_id[i] = i, base = seeded permutation
    (domain_builder.rs: DomainBuilder::new(n, seed)); one semantic group "common"."""
    perm = np.arange(n, dtype=np.int64)
    r = splitmix64(seed, n)
    for i in range(n - 1, 0, -1):  # Fisher-Yates
        j = int(r[i] % np.uint64(i + 1))
        perm[i], perm[j] = perm[j], perm[i]
    return ProblemSpec(
        kind=NQUEENS, n_vars=n,
        lower_bounds=np.zeros(n), upper_bounds=np.full(n, float(n - 1)),
        initial=perm.astype(np.float64),
        groups={"common": np.arange(n, dtype=np.int32)},
        column_id=np.arange(n, dtype=np.int64),
        score_precision=None, name=f"nqueens-{n}")


def tsp(n_cities: int, seed: int = 1, greedy: bool = True, with_matrix: bool = True) -> ProblemSpec:
    """examples/tsp: n_stops = n_cities-1 variables in [1, n_cities-1], depot = location 0
    (persistence/cotwin_builder.rs:49-77).  with_matrix=False leaves distance_matrix None (the
    device builds it from the coordinates: Problem(spec, use_coords=True)) -- for instances whose
    matrix is gigabytes."""
    xy = uniform01(seed, 2 * n_cities).reshape(n_cities, 2) * 1000.0
    D = distance_matrix(xy) if with_matrix else None
    greedy = greedy and with_matrix
    n = n_cities - 1
    init = np.arange(1, n_cities, dtype=np.float64)
    spec = ProblemSpec(
        kind=TSP, n_vars=n,
        lower_bounds=np.ones(n), upper_bounds=np.full(n, float(n_cities - 1)),
        initial=init, groups={"common": np.arange(n, dtype=np.int32)},
        n_locations=n_cities, distance_matrix=D, coords=xy, n_depots=1,
        score_precision=[3, 3], name=f"tsp-{n_cities}")
    if greedy:
        spec.initial = tsp_greedy(D)
    return spec


def tsp_greedy(D: np.ndarray) -> np.ndarray:
    """Nearest-neighbour init, examples/tsp/src/persistence/cotwin_builder.rs:87-117
    (ties -> lowest id)."""
    L = D.shape[0]
    used = np.zeros(L, dtype=bool)
    used[0] = True
    out = np.empty(L - 1, dtype=np.float64)
    prev = 0
    for i in range(L - 1):
        row = np.where(used, np.inf, D[prev])
        best = int(np.argmin(row))
        used[best] = True
        out[i] = best
        prev = best
    return out


def _vrp_common(kind, n_customers, n_vehicles, n_depots, seed, time_windowed, greedy):
    L = n_depots + n_customers
    xy = uniform01(seed, 2 * L).reshape(L, 2) * 1000.0
    D = distance_matrix(xy)
    demand = np.zeros(L, dtype=np.uint64)
    demand[n_depots:] = randint(seed + 101, n_customers, 1, 100).astype(np.uint64)
    total = int(demand.sum())
    cap = int(np.ceil(1.1 * total / n_vehicles))
    tw_start = np.zeros(L, dtype=np.uint64)
    tw_end = np.zeros(L, dtype=np.uint64)
    service = np.zeros(L, dtype=np.uint64)
    day_start = np.zeros(n_vehicles, dtype=np.uint64)
    day_end = np.zeros(n_vehicles, dtype=np.uint64)
    if time_windowed:
        st = randint(seed + 202, n_customers, 0, 43200)
        width = randint(seed + 303, n_customers, 3600, 14400)
        tw_start[n_depots:] = st.astype(np.uint64)
        tw_end[n_depots:] = (st + width).astype(np.uint64)
        service[n_depots:] = randint(seed + 404, n_customers, 300, 900).astype(np.uint64)
        day_end[:] = 86400
    n = 2 * n_customers
    lb = np.empty(n); ub = np.empty(n)
    lb[0::2] = 0.0; ub[0::2] = float(n_vehicles - 1)        # vehicle_id
    lb[1::2] = float(n_depots); ub[1::2] = float(L - 1)     # customer_id
    veh_ids = np.arange(0, n, 2, dtype=np.int32)
    cus_ids = np.arange(1, n, 2, dtype=np.int32)
    groups = {"vehicle_assignment": veh_ids, "customer_assignment": cus_ids}
    if kind == VRP:
        # examples/vrp/src/persistence/cotwin_builder.rs:127-133: both variables also sit
        # in "common"; vrp_service keeps the two groups disjoint (:128-132 there).
        groups["common"] = np.arange(n, dtype=np.int32)
    spec = ProblemSpec(
        kind=kind, n_vars=n, lower_bounds=lb, upper_bounds=ub, groups=groups,
        n_locations=L, distance_matrix=D, coords=xy, n_depots=n_depots,
        n_vehicles=n_vehicles,
        vehicle_depot=(np.arange(n_vehicles) % n_depots).astype(np.int64),
        vehicle_capacity=np.full(n_vehicles, cap, dtype=np.uint64),
        work_day_start=day_start, work_day_end=day_end,
        demand=demand, tw_start=tw_start, tw_end=tw_end, service_time=service,
        time_windowed=time_windowed, score_precision=[0, 0, 3],
        name=f"{KIND_NAMES[kind]}-{n_customers}x{n_vehicles}" + ("-tw" if time_windowed else ""))
    spec.initial = vrp_greedy(spec) if greedy else vrp_round_robin(spec)
    return spec


def cvrp(n_customers: int, n_vehicles: int, seed: int = 2, greedy: bool = True) -> ProblemSpec:
    """Config C3: 1 depot, demand U{1..100}, capacity = ceil(1.1*sum/k); examples/vrp."""
    return _vrp_common(VRP, n_customers, n_vehicles, 1, seed, False, greedy)


def vrptw(n_stops: int, n_vehicles: int, n_depots: int = 5, seed: int = 3,
          service_variant: bool = True, greedy: bool = True) -> ProblemSpec:
    """Config C4: time-window VRP; service_variant picks the vrp_service lateness rule
    (SURVEY.md Q3)."""
    kind = VRP_SERVICE if service_variant else VRP
    return _vrp_common(kind, n_stops, n_vehicles, n_depots, seed, True, greedy)


def vrp_round_robin(spec: ProblemSpec) -> np.ndarray:
    n_stops = spec.n_vars // 2
    out = np.empty(spec.n_vars)
    out[0::2] = np.arange(n_stops) % spec.n_vehicles
    out[1::2] = np.arange(n_stops) + spec.n_depots
    return out


def vrp_greedy(spec: ProblemSpec) -> np.ndarray:
    """examples/vrp/src/persistence/cotwin_builder.rs:153-255: fill vehicles one by one with
    the nearest unassigned customer until the next one no longer fits.  Customers left over
    (reference: None -> random sample) are appended round-robin here so that the start
    vector is deterministic."""
    L, nd, K = spec.n_locations, spec.n_depots, spec.n_vehicles
    D = spec.distance_matrix
    used = np.zeros(L, dtype=bool)
    used[:nd] = True
    veh, cus = [], []
    remaining = L - nd
    for k in range(K):
        if remaining <= 0:
            break
        prev = int(spec.vehicle_depot[k])
        cap = int(spec.vehicle_capacity[k])
        collected = 0
        while collected < cap and remaining > 0:
            row = np.where(used, np.inf, D[prev])
            best = int(np.argmin(row))
            dem = int(spec.demand[best])
            if collected + dem <= cap:
                collected += dem
                used[best] = True
                remaining -= 1
                veh.append(k); cus.append(best)
                prev = best
            else:
                break
    left = np.nonzero(~used)[0]
    for j, c in enumerate(left):
        veh.append(j % K); cus.append(int(c))
    out = np.empty(spec.n_vars)
    out[0::2] = veh
    out[1::2] = cus
    return out


# ---- TSPLIB files (host side only) -------------------------------------
def read_tsplib(text: str):
    """DomainBuilder::read_tsp_file (examples/tsp/src/persistence/domain_builder.rs:90-192): header
    lines up to NODE_COORD_SECTION (NAME, EDGE_WEIGHT_TYPE kept: last space-separated token), then
    `id x y [name]` lines up to EOF; for a type other than EUC_2D a distance matrix follows (one row
    per line, up to the next EOF).  Returns (metadata, coords f64 [L, 2], matrix or None)."""
    lines = iter(text.splitlines())
    meta = {}
    for line in lines:
        if "NODE_COORD_SECTION" in line:
            break
        if "NAME" in line:
            meta["dataset_name"] = line.split(" ")[-1].strip()
        if "EDGE_WEIGHT_TYPE" in line:
            meta["distance_type"] = line.split(" ")[-1].strip()
    else:
        raise ValueError("no NODE_COORD_SECTION")
    if "distance_type" not in meta:
        raise ValueError("no EDGE_WEIGHT_TYPE")
    xy = []
    for line in lines:
        if "EOF" in line:
            break
        parts = line.split()
        if len(parts) < 3:
            raise ValueError(f"bad location line {line!r}")
        xy.append((float(parts[1]), float(parts[2])))
    coords = np.array(xy, dtype=np.float64).reshape(-1, 2)
    matrix = None
    if "EUC_2D" not in meta["distance_type"]:
        rows = []
        for line in lines:
            if "EOF" in line:
                break
            parts = line.split(" ")[:-1]          # the reference drops the last token of every row
            rows.append([float(x) for x in parts if x != ""])
        if rows:
            matrix = np.array(rows, dtype=np.float64)
    return meta, coords, matrix


def tsp_from_tsplib(text: str, greedy: bool = True) -> ProblemSpec:
    """The TSP model of examples/tsp over a TSPLIB instance: depot = first location, one stop
    variable per remaining location (cotwin_builder.rs:49-77), distances by the example's formula
    (location.rs:38-50: Euclid truncated to 3 decimals) unless the file carries a matrix."""
    meta, xy, matrix = read_tsplib(text)
    L = xy.shape[0]
    if L < 2:
        raise ValueError("need at least two locations")
    D = matrix if matrix is not None else distance_matrix(xy)
    if D.shape != (L, L):
        raise ValueError(f"distance matrix is {D.shape}, expected {(L, L)}")
    n = L - 1
    spec = ProblemSpec(
        kind=TSP, n_vars=n, lower_bounds=np.ones(n), upper_bounds=np.full(n, float(L - 1)),
        initial=np.arange(1, L, dtype=np.float64), groups={"common": np.arange(n, dtype=np.int32)},
        n_locations=L, distance_matrix=D, coords=xy, n_depots=1, score_precision=[3, 3],
        name=meta.get("dataset_name", "tsplib"))
    if greedy:
        spec.initial = tsp_greedy(D)
    return spec


# ---- VRP files, domain round trips, frozen replanning (host side only) --------
def read_vrp(text: str):
    """DomainBuilder::read_vrp_file (examples/vrp/src/persistence/domain_builder.rs:145-331): header up to
    NODE_COORD_SECTION (vehicles count = the `-kNN` suffix of NAME, CAPACITY, EDGE_WEIGHT_TYPE, TYPE: last
    space-separated token of the line), `id x y [name]` lines up to DEMAND_SECTION / EOF, an explicit
    matrix (rows up to EOF) unless the type is EUC_2D, demand lines `id demand [tw_start tw_end service]`
    up to DEPOT_SECTION / EOF, depot ids up to -1 / EOF.
    Returns (metadata, coords [L, 2], ids, matrix or None, demand_info [[...]], depot_ids)."""
    lines = iter(text.splitlines())
    meta = {}
    for line in lines:
        if "NODE_COORD_SECTION" in line:
            break
        last = line.split(" ")[-1].strip()
        if "NAME" in line:
            meta["dataset_name"] = last
            meta["vehicles_count"] = last.split("-")[-1].replace("k", "")
        if "TYPE" in line:
            meta["task_type"] = last
        if "EDGE_WEIGHT_TYPE" in line:
            meta["distance_type"] = last
        if "CAPACITY" in line:
            meta["vehicles_capacity"] = last
    else:
        raise ValueError("no NODE_COORD_SECTION")
    for key in ("vehicles_count", "vehicles_capacity", "distance_type"):
        if key not in meta:
            raise ValueError(f"header lacks {key}")
    xy, ids = [], []
    for line in lines:
        if "EOF" in line or "DEMAND_SECTION" in line:
            break
        parts = line.split()
        if len(parts) < 3:
            raise ValueError(f"bad customer line {line!r}")
        ids.append(int(parts[0]))
        xy.append((float(parts[1]), float(parts[2])))
    matrix = None
    if "EUC_2D" not in meta["distance_type"]:
        rows = []
        for line in lines:
            if "EOF" in line:
                break
            parts = line.split(" ")[:-1]          # the reference drops the last token of every row
            rows.append([float(x) for x in parts if x != ""])
        if rows:
            matrix = np.array(rows, dtype=np.float64)
    demand_info = []
    depot_ids = []
    in_depots = False
    for line in lines:
        if "DEMAND_SECTION" in line:              # (the header line itself, when a matrix came first)
            continue
        if "DEPOT_SECTION" in line:
            in_depots = True
            continue
        if "EOF" in line:
            if in_depots:
                break
            continue
        s = line.strip()
        if not s:
            continue
        if in_depots:
            if "-1" in s:
                break
            depot_ids.append(int(s))
        else:
            demand_info.append([int(x) for x in s.split(" ") if x != ""])
    return meta, np.array(xy, dtype=np.float64).reshape(-1, 2), ids, matrix, demand_info, depot_ids


def vrp_from_file(text: str, service_variant: bool = False, greedy: bool = True) -> ProblemSpec:
    """DomainBuilder::build_domain_from_scratch (:18-91) + CotwinBuilder::build_cotwin of examples/vrp
    (service_variant: examples/vrp_service): depots = the first len(DEPOT_SECTION) locations, vehicle i
    starts at depot i % n_depots with the depot's window as its work day, every other location is one
    planning stop (vehicle_id in [0, K-1], customer_id in [n_depots, L-1]); distances by the example's
    formula unless the file carries a matrix (both truncated to 3 decimals)."""
    meta, xy, ids, matrix, demand_info, depot_ids = read_vrp(text)
    L = xy.shape[0]
    if len(demand_info) != L:
        raise ValueError("Customers or demands have been readed incorrect")
    n_depots = len(depot_ids)
    if n_depots < 1 or n_depots >= L:
        raise ValueError("DEPOT_SECTION must name at least one depot and leave customers")
    K = int(meta["vehicles_count"])
    cap = int(meta["vehicles_capacity"])
    demand = np.zeros(L, dtype=np.uint64)
    tw_start = np.zeros(L, dtype=np.uint64); tw_end = np.zeros(L, dtype=np.uint64); service = np.zeros(L, dtype=np.uint64)
    time_windowed = False
    for i, row in enumerate(demand_info):
        if row[0] != ids[i]:
            raise ValueError("Invalid customer to demand mapping")
        demand[i] = row[1]
        if len(row) == 5:
            time_windowed = True
            tw_start[i], tw_end[i], service[i] = row[2], row[3], row[4]
    if matrix is not None:
        D = np.floor(matrix) + np.floor((matrix - np.floor(matrix)) * 1000.0) / 1000.0      # round(dm, 3), math_utils.rs:10-13
    else:
        D = distance_matrix(xy)
    if D.shape != (L, L):
        raise ValueError(f"distance matrix is {D.shape}, expected {(L, L)}")
    kind = VRP_SERVICE if service_variant else VRP
    n_stops = L - n_depots
    n = 2 * n_stops
    lb = np.empty(n); ub = np.empty(n)
    lb[0::2] = 0.0; ub[0::2] = float(K - 1)
    lb[1::2] = float(n_depots); ub[1::2] = float(L - 1)
    groups = {"vehicle_assignment": np.arange(0, n, 2, dtype=np.int32), "customer_assignment": np.arange(1, n, 2, dtype=np.int32)}
    if kind == VRP:
        groups["common"] = np.arange(n, dtype=np.int32)
    depot_of = (np.arange(K) % n_depots).astype(np.int64)
    spec = ProblemSpec(
        kind=kind, n_vars=n, lower_bounds=lb, upper_bounds=ub, groups=groups, n_locations=L, distance_matrix=D,
        coords=xy, n_depots=n_depots, n_vehicles=K, vehicle_depot=depot_of,
        vehicle_capacity=np.full(K, cap, dtype=np.uint64),
        work_day_start=tw_start[depot_of].copy(), work_day_end=tw_end[depot_of].copy(),
        demand=demand, tw_start=tw_start, tw_end=tw_end, service_time=service, time_windowed=time_windowed,
        score_precision=[0, 0, 3], name=meta["dataset_name"])
    spec.initial = vrp_greedy(spec) if greedy else np.full(n, np.nan)
    return spec


def vrp_routes_from_solution(spec: ProblemSpec, values) -> list:
    """DomainBuilder::build_from_solution (examples/vrp .../domain_builder.rs:93-141): walk the planning
    stops in entity order and append every stop's customer to its vehicle -> one customer list per
    vehicle (visiting order = entity order, which is what the scorers assume)."""
    v = np.asarray(values, dtype=np.float64)
    routes = [[] for _ in range(spec.n_vehicles)]
    for s in range(spec.n_vars // 2):
        routes[int(v[2 * s])].append(int(v[2 * s + 1]))
    return routes


def tsp_path_from_solution(spec: ProblemSpec, values) -> list:
    """examples/tsp DomainBuilder::build_from_solution (:58-77): the vehicle's trip path as location ids"""
    return [int(x) for x in np.asarray(values, dtype=np.float64)]


def vrp_replanning_spec(spec: ProblemSpec, routes, frozen_vehicles=(), drop_vehicles=()) -> ProblemSpec:
    """The replanning scenario of examples/vrp/src/main.rs:120-141: an existing plan (routes per vehicle)
    becomes the start of a new solve -- optionally with vehicles removed and the customers of some
    vehicles pinned.  CotwinBuilder::build_planning_stops with is_already_initialized
    (cotwin_builder.rs:108-119): planning stop i takes the i-th (vehicle, customer) pair of the plan in
    vehicle-major route order as its initial value, frozen when the customer is pinned.  Customers of
    dropped vehicles lose their assignment (initial None -> sampled)."""
    import copy
    out = copy.deepcopy(spec)
    keep = [k for k in range(spec.n_vehicles) if k not in set(drop_vehicles)]
    K = len(keep)
    if K < 1:
        raise ValueError("no vehicle left")
    n_stops = spec.n_vars // 2
    out.n_vehicles = K
    out.vehicle_depot = np.asarray(spec.vehicle_depot)[keep].copy()
    out.vehicle_capacity = np.asarray(spec.vehicle_capacity)[keep].copy()
    out.work_day_start = np.asarray(spec.work_day_start)[keep].copy()
    out.work_day_end = np.asarray(spec.work_day_end)[keep].copy()
    out.upper_bounds = spec.upper_bounds.copy()
    out.upper_bounds[0::2] = float(K - 1)
    initial = np.full(spec.n_vars, np.nan)
    frozen = np.zeros(spec.n_vars, dtype=np.uint8)
    i = 0
    for new_k, old_k in enumerate(keep):
        for c in routes[old_k]:
            initial[2 * i], initial[2 * i + 1] = float(new_k), float(c)
            if old_k in set(frozen_vehicles):
                frozen[2 * i] = frozen[2 * i + 1] = 1
            i += 1
    if i > n_stops:
        raise ValueError("the plan holds more stops than the problem")
    out.initial = initial
    out.frozen = frozen
    return out


def write_vrp(spec: ProblemSpec, name: str = None) -> str:
    """A ProblemSpec of the VRP family as a file read_vrp / the reference's DomainBuilder accepts (EUC_2D);
    used for golden round trips and for handing synthetic instances to the reference-side dumper."""
    name = name or f"synthetic-n{spec.n_locations}-k{spec.n_vehicles}"
    if not name.split("-")[-1].startswith("k"):
        name += f"-k{spec.n_vehicles}"
    out = [f"NAME : {name}", "COMMENT : greyjack-b200 synthetic", "TYPE : CVRP", f"DIMENSION : {spec.n_locations}",
           "EDGE_WEIGHT_TYPE : EUC_2D", f"CAPACITY : {int(spec.vehicle_capacity[0])}", "NODE_COORD_SECTION"]
    for i in range(spec.n_locations):
        out.append(f"{i + 1} {float(spec.coords[i, 0])!r} {float(spec.coords[i, 1])!r}")
    out.append("DEMAND_SECTION")
    for i in range(spec.n_locations):
        if spec.time_windowed:
            out.append(f"{i + 1} {int(spec.demand[i])} {int(spec.tw_start[i])} {int(spec.tw_end[i])} {int(spec.service_time[i])}")
        else:
            out.append(f"{i + 1} {int(spec.demand[i])}")
    out.append("DEPOT_SECTION")
    for d in range(spec.n_depots):
        out.append(f"{d + 1}")
    out += ["-1", "EOF", ""]
    return "\n".join(out)


def write_tsplib(spec: ProblemSpec, name: str = None) -> str:
    """A TSP ProblemSpec as an EUC_2D TSPLIB file (read_tsplib / the reference's DomainBuilder)"""
    out = [f"NAME : {name or spec.name}", "TYPE : TSP", f"DIMENSION : {spec.n_locations}", "EDGE_WEIGHT_TYPE : EUC_2D",
           "NODE_COORD_SECTION"]
    for i in range(spec.n_locations):
        out.append(f"{i + 1} {float(spec.coords[i, 0])!r} {float(spec.coords[i, 1])!r}")
    out += ["EOF", ""]
    return "\n".join(out)


def write_vrp_service_json(spec: ProblemSpec, name: str = None) -> dict:
    """The JSON domain examples/vrp_service reads (persistence/domain_builder.rs:20-62): metadata strings,
    customers_dict {n_customers, "0": {...}, ...}, depot_dict {n_depots, "0": id, ...}."""
    cust = {"n_customers": int(spec.n_locations)}
    for i in range(spec.n_locations):
        cust[str(i)] = {"id": i + 1, "name": str(i + 1), "latitude": float(spec.coords[i, 0]),
                        "longitude": float(spec.coords[i, 1]), "demand": float(spec.demand[i]),
                        "time_window_start": int(spec.tw_start[i]), "time_window_end": int(spec.tw_end[i]),
                        "service_time": int(spec.service_time[i])}
    depots = {"n_depots": int(spec.n_depots)}
    for d in range(spec.n_depots):
        depots[str(d)] = d + 1
    return {"metadata": {"dataset_name": name or spec.name, "distance_type": "EUC_2D", "task_type": "CVRPTW",
                         "time_window_task_type": "true" if spec.time_windowed else "false",
                         "vehicles_capacity": str(int(spec.vehicle_capacity[0])), "vehicles_count": str(int(spec.n_vehicles))},
            "customers_dict": cust, "depot_dict": depots}


def vrp_service_from_json(doc: dict, greedy: bool = True) -> ProblemSpec:
    """examples/vrp_service DomainBuilder::build_domain_from_scratch (:20-110) + its CotwinBuilder: the
    same model as vrp_from_file with the service lateness rule and two disjoint semantic groups."""
    meta = doc["metadata"]
    n = int(doc["customers_dict"]["n_customers"])
    rows = [doc["customers_dict"][str(i)] for i in range(n)]
    tw = str(meta["time_window_task_type"]).replace('"', "") == "true"
    lines = [f"NAME : {meta['dataset_name']}-k{int(meta['vehicles_count'])}", "EDGE_WEIGHT_TYPE : EUC_2D",
             f"CAPACITY : {int(meta['vehicles_capacity'])}", "NODE_COORD_SECTION"]
    lines += [f"{r['id']} {float(r['latitude'])!r} {float(r['longitude'])!r}" for r in rows]
    lines.append("DEMAND_SECTION")
    for r in rows:
        lines.append(f"{r['id']} {int(r['demand'])}" + (f" {int(r['time_window_start'])} {int(r['time_window_end'])} {int(r['service_time'])}" if tw else ""))
    lines.append("DEPOT_SECTION")
    lines += [str(int(doc["depot_dict"][str(d)])) for d in range(int(doc["depot_dict"]["n_depots"]))]
    lines += ["-1", "EOF", ""]
    spec = vrp_from_file("\n".join(lines), service_variant=True, greedy=greedy)
    spec.name = str(meta["dataset_name"])
    return spec
