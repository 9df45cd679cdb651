"""ctypes binding of the CPU ORACLE (oracle/gj_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never by the product package.  Builds libgj_oracle.so on demand with gcc."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgj_oracle.so")


def _host_signature() -> str:
    """CPU model + ISA flags of THIS host: the oracle is built with -march=native, so
    a library that travelled from another machine (the build container -> the GPU box) is rebuilt."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        keep = [ln for ln in txt.splitlines() if ln.startswith(("model name", "flags"))][:2]
        return hashlib.sha1("\n".join(keep).encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gj_oracle.c")
    hdr = os.path.join(_HERE, "gj_oracle.h")
    stamp = _SO + ".host"
    sig = _host_signature()
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(f) > os.path.getmtime(_SO) for f in (src, hdr))
    if not stale:
        try:
            with open(stamp) as f:
                stale = f.read().strip() != sig
        except OSError:
            stale = True
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libgj_oracle.so"])
        with open(stamp, "w") as f:
            f.write(sig)
    return _SO


class _Problem(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_vars", C.c_int32),
        ("lower_bounds", C.c_void_p), ("upper_bounds", C.c_void_p),
        ("frozen", C.c_void_p), ("initial", C.c_void_p),
        ("column_id", C.c_void_p),
        ("n_locations", C.c_int32), ("distance_matrix", C.c_void_p),
        ("n_vehicles", C.c_int32), ("vehicle_depot", C.c_void_p),
        ("vehicle_capacity", C.c_void_p), ("work_day_start", C.c_void_p),
        ("work_day_end", C.c_void_p), ("demand", C.c_void_p), ("tw_start", C.c_void_p),
        ("tw_end", C.c_void_p), ("service_time", C.c_void_p),
        ("time_windowed", C.c_int32), ("weights", C.c_double * 4),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.gjo_rint.restype = C.c_double; L.gjo_rint.argtypes = [C.c_double]
        L.gjo_round.restype = C.c_double; L.gjo_round.argtypes = [C.c_double, C.c_uint64]
        L.gjo_fix_integer.restype = C.c_double
        L.gjo_fix_integer.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_double]
        L.gjo_fix_float.restype = C.c_double
        L.gjo_fix_float.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_double]
        L.gjo_inverse_transform_integer.restype = C.c_int64
        L.gjo_inverse_transform_integer.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_double]
        L.gjo_score_cmp.restype = C.c_int
        L.gjo_score_le.restype = C.c_int
        L.gjo_fitness.restype = C.c_double
        L.gjo_ts_select.restype = C.c_int64
        L.gjo_bench_ts.restype = C.c_int64
        L.gjo_bench_plain.restype = C.c_int64
        L.gjo_bench_la.restype = C.c_int64
        L.gjo_bench_ga.restype = C.c_int64
        L.gjo_ga_select.restype = C.c_int64
        L.gjo_ga_select.argtypes = [C.c_double, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleProblem:
    """Keeps numpy arrays alive behind a gjo_problem struct."""

    def __init__(self, spec):
        self.spec = spec
        k = {}
        k["lb"] = np.ascontiguousarray(spec.lower_bounds, dtype=np.float64)
        k["ub"] = np.ascontiguousarray(spec.upper_bounds, dtype=np.float64)
        k["frozen"] = None if spec.frozen is None else np.ascontiguousarray(spec.frozen, dtype=np.uint8)
        k["initial"] = None if spec.initial is None else np.ascontiguousarray(spec.initial, dtype=np.float64)
        k["column_id"] = None if spec.column_id is None else np.ascontiguousarray(spec.column_id, dtype=np.int64)
        k["D"] = None if spec.distance_matrix is None else np.ascontiguousarray(spec.distance_matrix, dtype=np.float64)
        for name, dt in (("vehicle_depot", np.int64), ("vehicle_capacity", np.uint64),
                         ("work_day_start", np.uint64), ("work_day_end", np.uint64),
                         ("demand", np.uint64), ("tw_start", np.uint64), ("tw_end", np.uint64),
                         ("service_time", np.uint64)):
            v = getattr(spec, name)
            k[name] = None if v is None else np.ascontiguousarray(v, dtype=dt)
        self._keep = k
        p = _Problem()
        p.kind = spec.kind; p.n_vars = spec.n_vars
        p.lower_bounds = _ptr(k["lb"]); p.upper_bounds = _ptr(k["ub"])
        p.frozen = _ptr(k["frozen"]); p.initial = _ptr(k["initial"])
        p.column_id = _ptr(k["column_id"])
        p.n_locations = spec.n_locations; p.distance_matrix = _ptr(k["D"])
        p.n_vehicles = spec.n_vehicles
        for name in ("vehicle_depot", "vehicle_capacity", "work_day_start", "work_day_end",
                     "demand", "tw_start", "tw_end", "service_time"):
            setattr(p, name, _ptr(k[name]))
        p.time_windowed = int(bool(spec.time_windowed))
        for i in range(4):
            p.weights[i] = float(spec.weights[i])
        self.c = p
        self.levels = spec.levels

    # -- scorers ------------------------------------------------------------------
    def score_plain(self, samples) -> np.ndarray:
        x = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, self.spec.n_vars)
        out = np.empty((x.shape[0], self.levels), dtype=np.float64)
        rc = lib().gjo_score_plain(C.byref(self.c), _ptr(x), C.c_int64(x.shape[0]), _ptr(out))
        assert rc == 0
        return out

    def score_incremental(self, base, deltas) -> np.ndarray:
        """deltas: list of lists of (var_id, value) -- Vec<Vec<(usize, f64)>>."""
        offsets, ids, vals = deltas_to_csr(deltas)
        return self.score_incremental_csr(base, offsets, ids, vals)

    def score_incremental_csr(self, base, offsets, ids, vals) -> np.ndarray:
        b = np.ascontiguousarray(base, dtype=np.float64)
        S = len(offsets) - 1
        out = np.empty((S, self.levels), dtype=np.float64)
        rc = lib().gjo_score_incremental(C.byref(self.c), _ptr(b), _ptr(offsets), _ptr(ids),
                                         _ptr(vals), C.c_int64(S), _ptr(out))
        assert rc == 0
        return out

    # -- mover --------------------------------------------------------------------
    def _move(self, fn, cand, group_ids, args, incremental):
        cand = np.ascontiguousarray(cand, dtype=np.float64)
        g = np.ascontiguousarray(group_ids, dtype=np.int32)
        n = self.spec.n_vars
        cols = np.zeros(2 * n + 16, dtype=np.int32)
        vals = np.zeros(2 * n + 16, dtype=np.float64)
        out = np.zeros(n, dtype=np.float64)
        k = fn(_ptr(cand), C.c_int(n), _ptr(g), C.c_int(len(g)), *args, C.c_int(int(incremental)),
               _ptr(cols), _ptr(vals), _ptr(out))
        if k < 0:
            return None
        if incremental:
            return cols[:k].copy(), vals[:k].copy()
        return cols[:k].copy(), out

    def move_change(self, cand, group_ids, chosen, new_values, incremental=True):
        ch = np.ascontiguousarray(chosen, dtype=np.int32)
        nv = np.ascontiguousarray(new_values, dtype=np.float64)
        return self._move(lib().gjo_move_change, cand, group_ids, (_ptr(ch), C.c_int(len(ch)), _ptr(nv)), incremental)

    def move_swap(self, cand, group_ids, chosen, incremental=True):
        ch = np.ascontiguousarray(chosen, dtype=np.int32)
        return self._move(lib().gjo_move_swap, cand, group_ids, (_ptr(ch), C.c_int(len(ch))), incremental)

    def move_swap_edges(self, cand, group_ids, chosen, incremental=True):
        ch = np.ascontiguousarray(chosen, dtype=np.int32)
        return self._move(lib().gjo_move_swap_edges, cand, group_ids, (_ptr(ch), C.c_int(len(ch))), incremental)

    def move_scramble(self, cand, group_ids, start, perm, incremental=True):
        pm = np.ascontiguousarray(perm, dtype=np.int32)
        return self._move(lib().gjo_move_scramble, cand, group_ids,
                          (C.c_int(start), C.c_int(len(pm)), _ptr(pm)), incremental)

    def move_insertion(self, cand, group_ids, get_out, put_in, incremental=True):
        return self._move(lib().gjo_move_insertion, cand, group_ids,
                          (C.c_int(get_out), C.c_int(put_in)), incremental)

    def move_inverse(self, cand, group_ids, a, b, incremental=True):
        return self._move(lib().gjo_move_inverse, cand, group_ids, (C.c_int(a), C.c_int(b)), incremental)

    def fix_deltas(self, cols, vals):
        c = np.ascontiguousarray(cols, dtype=np.int32)
        v = np.array(vals, dtype=np.float64)
        lib().gjo_fix_deltas(C.byref(self.c), _ptr(c), _ptr(v), C.c_int(len(c)))
        return v

    def fix_variables(self, cand, cols):
        c = np.ascontiguousarray(cols, dtype=np.int32)
        v = np.array(cand, dtype=np.float64)
        lib().gjo_fix_variables(C.byref(self.c), _ptr(v), _ptr(c), C.c_int(len(c)))
        return v

    # -- baselines ----------------------------------------------------------------
    def bench_ts(self, base, n_moves, n_steps, n_threads, seed, move_probas, precision):
        b = np.ascontiguousarray(base, dtype=np.float64)
        mp = np.ascontiguousarray(move_probas, dtype=np.float64)
        pr = None if precision is None else np.ascontiguousarray(precision, dtype=np.int64)
        secs = C.c_double(0.0)
        best = np.zeros(3, dtype=np.float64)
        n = lib().gjo_bench_ts(C.byref(self.c), _ptr(b), C.c_int(n_moves), C.c_int(n_steps),
                               C.c_int(n_threads), C.c_uint64(seed), _ptr(mp), _ptr(pr),
                               C.byref(secs), _ptr(best))
        return int(n), secs.value, best[: self.levels].copy()

    def _groups_csr(self):
        offs, ids = [0], []
        for g in self.spec.groups.values():
            ids.extend(int(x) for x in g)
            offs.append(len(ids))
        return np.asarray(offs, dtype=np.int64), np.asarray(ids, dtype=np.int32)

    def bench_la(self, base, late_size, n_steps, n_threads, seed, move_probas, precision):
        """One LateAcceptance agent per thread (late_acceptance_base.rs:116-241) -> (candidates, seconds, best)."""
        b = np.ascontiguousarray(base, dtype=np.float64)
        mp = np.ascontiguousarray(move_probas, dtype=np.float64)
        pr = None if precision is None else np.ascontiguousarray(precision, dtype=np.int64)
        offs, ids = self._groups_csr()
        secs = C.c_double(0.0)
        best = np.zeros(3, dtype=np.float64)
        n = lib().gjo_bench_la(C.byref(self.c), _ptr(b), _ptr(offs), _ptr(ids), C.c_int(len(offs) - 1),
                               C.c_int(late_size), C.c_int(n_steps), C.c_int(n_threads), C.c_uint64(seed),
                               _ptr(mp), _ptr(pr), C.byref(secs), _ptr(best))
        return int(n), secs.value, best[: self.levels].copy()

    def bench_ga(self, pop, crossover_probability, p_best_rate, n_generations, n_threads, seed, move_probas, precision):
        """One GeneticAlgorithm agent per thread (genetic_algorithm_base.rs:141-213) -> (candidates, seconds, best)."""
        mp = np.ascontiguousarray(move_probas, dtype=np.float64)
        pr = None if precision is None else np.ascontiguousarray(precision, dtype=np.int64)
        offs, ids = self._groups_csr()
        secs = C.c_double(0.0)
        best = np.zeros(3, dtype=np.float64)
        n = lib().gjo_bench_ga(C.byref(self.c), _ptr(offs), _ptr(ids), C.c_int(len(offs) - 1), C.c_int(pop),
                               C.c_double(crossover_probability), C.c_double(p_best_rate), C.c_int(n_generations),
                               C.c_int(n_threads), C.c_uint64(seed), _ptr(mp), _ptr(pr), C.byref(secs), _ptr(best))
        return int(n), secs.value, best[: self.levels].copy()

    def bench_plain(self, samples, n_threads, repeats=1):
        x = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, self.spec.n_vars)
        out = np.empty((x.shape[0], self.levels), dtype=np.float64)
        secs = C.c_double(0.0)
        n = lib().gjo_bench_plain(C.byref(self.c), _ptr(x), C.c_int64(x.shape[0]),
                                  C.c_int(n_threads), C.c_int(repeats), C.byref(secs), _ptr(out))
        return int(n), secs.value, out


def deltas_to_csr(deltas):
    offsets = np.zeros(len(deltas) + 1, dtype=np.uint64)
    tot = 0
    for i, d in enumerate(deltas):
        tot += len(d)
        offsets[i + 1] = tot
    ids = np.zeros(max(tot, 1), dtype=np.uint64)
    vals = np.zeros(max(tot, 1), dtype=np.float64)
    j = 0
    for d in deltas:
        for (vid, val) in d:
            ids[j] = vid; vals[j] = val; j += 1
    return offsets, ids, vals


# -- free helpers --------------------------------------------------------------------
def rint(x): return lib().gjo_rint(C.c_double(x))
def round_(x, p): return lib().gjo_round(C.c_double(x), C.c_uint64(p))
def fix_integer(v, lb, ub, frozen=False, initial=0.0):
    return lib().gjo_fix_integer(C.c_double(v), C.c_double(lb), C.c_double(ub), C.c_int(int(frozen)), C.c_double(initial))
def fix_float(v, lb, ub, frozen=False, initial=0.0):
    return lib().gjo_fix_float(C.c_double(v), C.c_double(lb), C.c_double(ub), C.c_int(int(frozen)), C.c_double(initial))
def inverse_transform_integer(v, lb, ub, frozen=False, initial=0.0):
    return lib().gjo_inverse_transform_integer(C.c_double(v), C.c_double(lb), C.c_double(ub), C.c_int(int(frozen)), C.c_double(initial))


def score_cmp(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    return lib().gjo_score_cmp(_ptr(a), _ptr(b), C.c_int(len(a)))


def score_le(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    return bool(lib().gjo_score_le(_ptr(a), _ptr(b), C.c_int(len(a))))


def score_round(s, precision):
    s = np.array(s, dtype=np.float64)
    if precision is None:
        return s
    pr = np.ascontiguousarray(precision, dtype=np.int64)
    flat = s.reshape(-1, len(pr))
    for row in flat:
        lib().gjo_score_round(_ptr(row), _ptr(pr), C.c_int(len(pr)))
    return flat.reshape(s.shape)


def fitness(s):
    s = np.ascontiguousarray(s, dtype=np.float64)
    return lib().gjo_fitness(_ptr(s), C.c_int(len(s)))


def sort_scores(scores):
    s = np.array(scores, dtype=np.float64)
    if s.ndim == 1:
        s = s.reshape(-1, 1)
    lib().gjo_sort_scores(_ptr(s), C.c_int(s.shape[0]), C.c_int(s.shape[1]))
    return s


def ts_select(scores, current):
    s = np.ascontiguousarray(scores, dtype=np.float64)
    cur = np.ascontiguousarray(current, dtype=np.float64)
    acc = C.c_int(0)
    idx = lib().gjo_ts_select(_ptr(s), C.c_int64(s.shape[0]), C.c_int(s.shape[1]), _ptr(cur), C.byref(acc))
    return int(idx), bool(acc.value)


def la_accept(cand, current, late, late_size):
    """late: list of score vectors, front first.  Returns (accepted, new_late)."""
    levels = len(cand)
    buf = np.zeros((late_size + 2, levels), dtype=np.float64)
    for i, s in enumerate(late):
        buf[i] = s
    n = C.c_int(len(late))
    c = np.ascontiguousarray(cand, dtype=np.float64)
    cur = np.ascontiguousarray(current, dtype=np.float64)
    acc = lib().gjo_la_accept(_ptr(c), _ptr(cur), _ptr(buf), C.byref(n), C.c_int(late_size), C.c_int(levels))
    return bool(acc), [buf[i].copy() for i in range(n.value)]


def sa_accept(cand, current, temperature, cooling_rate, inverted_accomplish_rate, u):
    """Returns (accepted, new_temperature, accept_proba)."""
    c = np.ascontiguousarray(cand, dtype=np.float64)
    cur = np.ascontiguousarray(current, dtype=np.float64)
    t = np.array(temperature, dtype=np.float64)
    proba = C.c_double(0.0)
    acc = lib().gjo_sa_accept(_ptr(c), _ptr(cur), C.c_int(len(c)), _ptr(t),
                              C.c_int(int(cooling_rate is not None)), C.c_double(cooling_rate or 0.0),
                              C.c_double(inverted_accomplish_rate), C.c_double(u), C.byref(proba))
    return bool(acc), t, proba.value


def ga_replace(cand_scores, pop_scores, worst_ids):
    cs = np.ascontiguousarray(cand_scores, dtype=np.float64)
    ps = np.ascontiguousarray(pop_scores, dtype=np.float64)
    w = np.ascontiguousarray(worst_ids, dtype=np.int64)
    out = np.zeros(len(w), dtype=np.int64)
    lib().gjo_ga_replace(_ptr(cs), _ptr(ps), _ptr(w), C.c_int64(len(w)), C.c_int(cs.shape[1]), _ptr(out))
    return out


def ga_select(p_best_proba, id_draw, pop, worst=False):
    """select_p_best / select_p_worst with explicit draws -> (index into the sorted population or -1, last_top_id)."""
    lt = C.c_int64(0)
    r = lib().gjo_ga_select(C.c_double(p_best_proba), C.c_int64(int(id_draw)), C.c_int64(pop), C.c_int(int(worst)),
                            C.cast(C.byref(lt), C.c_void_p))
    return int(r), int(lt.value)


def ga_cross(c1, c2, weight):
    """GeneticAlgorithmBase::cross for all-discrete problems -> (child1, child2)."""
    a = np.ascontiguousarray(c1, dtype=np.float64); b = np.ascontiguousarray(c2, dtype=np.float64)
    o1 = np.empty_like(a); o2 = np.empty_like(a)
    lib().gjo_ga_cross(_ptr(a), _ptr(b), C.c_int(len(a)), C.c_double(weight), None, _ptr(o1), _ptr(o2))
    return o1, o2


def distance_matrix(xy):
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    n = xy.shape[0]
    D = np.empty((n, n), dtype=np.float64)
    lib().gjo_distance_matrix(_ptr(xy), C.c_int(n), _ptr(D))
    return D


def tsp_greedy_init(D):
    D = np.ascontiguousarray(D, dtype=np.float64)
    out = np.empty(D.shape[0] - 1, dtype=np.float64)
    lib().gjo_tsp_greedy_init(_ptr(D), C.c_int(D.shape[0]), _ptr(out))
    return out
