"""Host-side mirror of the reference's scoring interface over the C ABI:

    OOPScoreRequester::request_score_plain        -> Problem.request_score_plain
    OOPScoreRequester::request_score_incremental  -> Problem.request_score_incremental
    {Plain,Incremental}ScoreCalculator::set_constraint_weights -> set_constraint_weights

(greyjack/src/score_calculation/score_requesters/oop_score_requester.rs:336-355, 443-463).
Everything is computed by the CUDA library; numpy only carries the buffers."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .instances import ProblemSpec


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def deltas_to_csr(deltas):
    """Vec<Vec<(usize, f64)>> -> (offsets u64 [S+1], var_ids u64, values f64)."""
    offsets = np.zeros(len(deltas) + 1, dtype=np.uint64)
    np.cumsum([len(d) for d in deltas], out=offsets[1:])
    tot = int(offsets[-1])
    ids = np.zeros(max(tot, 1), dtype=np.uint64)
    vals = np.zeros(max(tot, 1), dtype=np.float64)
    k = 0
    for d in deltas:
        for vid, val in d:
            ids[k] = vid
            vals[k] = val
            k += 1
    return offsets, ids, vals


class PinnedArray:
    """numpy view over page-locked host memory from gj_host_alloc (copies to / from it are
    asynchronous DMA).  Keep the object alive while the array is in use."""

    def __init__(self, shape, dtype):
        L = _lib.load()
        self._L = L
        dt = np.dtype(dtype)
        n = int(np.prod(shape))
        self.ptr = C.c_void_p()
        _lib.check(L.gj_host_alloc(C.c_size_t(max(1, n * dt.itemsize)), C.byref(self.ptr)))
        buf = (C.c_char * (n * dt.itemsize)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)

    def __del__(self):
        try:
            if self.ptr:
                self.array = None
                self._L.gj_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_copy(a):
    """Copy of `a` in page-locked host memory; returns (PinnedArray, ndarray view)."""
    a = np.ascontiguousarray(a)
    p = PinnedArray(a.shape, a.dtype)
    p.array[...] = a
    return p, p.array


class Problem:
    """Device-resident Cotwin + score requester."""

    def __init__(self, spec: ProblemSpec, device: int = 0, use_coords: bool = False):
        L = _lib.load()
        self.spec = spec
        self.levels = spec.levels
        self.n_vars = spec.n_vars
        keep = {}
        d = _lib.ProblemDesc()
        d.kind = spec.kind
        d.n_vars = spec.n_vars
        keep["lb"] = _c(spec.lower_bounds, np.float64); d.lower_bounds = _ptr(keep["lb"])
        keep["ub"] = _c(spec.upper_bounds, np.float64); d.upper_bounds = _ptr(keep["ub"])
        keep["frozen"] = _c(spec.frozen, np.uint8); d.frozen = _ptr(keep["frozen"])
        keep["initial"] = _c(spec.initial, np.float64); d.initial = _ptr(keep["initial"])
        names = list(spec.groups.keys())
        self.group_names = names
        if names:
            offs = np.zeros(len(names) + 1, dtype=np.int64)
            np.cumsum([len(spec.groups[k]) for k in names], out=offs[1:])
            ids = np.concatenate([np.asarray(spec.groups[k], dtype=np.int32) for k in names])
            keep["goffs"], keep["gids"] = offs, np.ascontiguousarray(ids)
            d.n_groups = len(names)
            d.group_offsets = _ptr(keep["goffs"]); d.group_var_ids = _ptr(keep["gids"])
        keep["col"] = _c(spec.column_id, np.int64); d.column_id = _ptr(keep["col"])
        d.n_locations = spec.n_locations
        if use_coords:
            keep["xy"] = _c(spec.coords, np.float64); d.coords = _ptr(keep["xy"])
        else:
            keep["D"] = _c(spec.distance_matrix, np.float64); d.distance_matrix = _ptr(keep["D"])
        d.n_vehicles = spec.n_vehicles
        for name, dt in (("vehicle_depot", np.int64), ("vehicle_capacity", np.uint64),
                         ("work_day_start", np.uint64), ("work_day_end", np.uint64),
                         ("demand", np.uint64), ("tw_start", np.uint64), ("tw_end", np.uint64),
                         ("service_time", np.uint64)):
            keep[name] = _c(getattr(spec, name), dt)
            setattr(d, name, _ptr(keep[name]))
        d.time_windowed = int(bool(spec.time_windowed))
        for i in range(4):
            d.weights[i] = float(spec.weights[i])
        for i in range(3):
            d.score_precision[i] = -1
        if spec.score_precision is not None:
            for i, pr in enumerate(spec.score_precision):
                d.score_precision[i] = int(pr)
        self._keep = keep
        self.device = device
        h = C.c_void_p()
        _lib.check(L.gj_problem_create(C.byref(d), C.c_int32(device), C.byref(h)))
        self.handle = h
        self._L = L

    def close(self):
        if getattr(self, "handle", None):
            self._L.gj_problem_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference-facing calls (host buffers) ------------------------------------------
    def request_score_plain(self, samples) -> np.ndarray:
        x = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, self.n_vars)
        out = np.empty((x.shape[0], self.levels), dtype=np.float64)
        _lib.check(self._L.gj_score_plain(self.handle, _ptr(x), C.c_int64(x.shape[0]), _ptr(out)))
        return out

    def request_score_incremental(self, sample, deltas) -> np.ndarray:
        offsets, ids, vals = deltas_to_csr(deltas)
        return self.request_score_incremental_csr(sample, offsets, ids, vals)

    def request_score_incremental_csr(self, sample, offsets, ids, vals, out=None) -> np.ndarray:
        b = np.ascontiguousarray(sample, dtype=np.float64)
        S = len(offsets) - 1
        if out is None:
            out = np.empty((S, self.levels), dtype=np.float64)
        _lib.check(self._L.gj_score_incremental(self.handle, _ptr(b), _ptr(offsets), _ptr(ids),
                                                _ptr(vals), C.c_int64(S), _ptr(out)))
        return out

    def request_score_incremental_packed(self, sample, offsets, ids_u32, vals_i32, out=None) -> np.ndarray:
        """Packed wire format (gj_score_incremental_packed): u32 ids, i32 already-decoded values."""
        b = np.ascontiguousarray(sample, dtype=np.float64)
        S = len(offsets) - 1
        if out is None:
            out = np.empty((S, self.levels), dtype=np.float64)
        _lib.check(self._L.gj_score_incremental_packed(self.handle, _ptr(b), _ptr(offsets), _ptr(ids_u32),
                                                       _ptr(vals_i32), C.c_int64(S), _ptr(out)))
        return out

    def set_constraint_weights(self, weights):
        w = np.ascontiguousarray(weights, dtype=np.float64)
        _lib.check(self._L.gj_problem_set_constraint_weights(self.handle, _ptr(w), C.c_int32(len(w))))

    def set_exact_sums(self, on: bool):
        """True (default): reference summation order, bit-exact float level; False: tree sums."""
        _lib.check(self._L.gj_problem_set_exact_sums(self.handle, C.c_int32(int(on))))

    def distance_matrix(self) -> np.ndarray:
        n = self.spec.n_locations
        out = np.empty((n, n), dtype=np.float64)
        _lib.check(self._L.gj_problem_get_distance_matrix(self.handle, _ptr(out)))
        return out

    # -- device-resident variants (raw device pointers, e.g. torch .data_ptr()) -------------
    def score_plain_device(self, d_samples: int, S: int, d_scores: int, stream: int = 0):
        _lib.check(self._L.gj_score_plain_device(self.handle, C.c_void_p(d_samples), C.c_int64(S),
                                                 C.c_void_p(d_scores), C.c_void_p(stream)))

    def score_plain_i32_device(self, d_samples: int, row_stride: int, S: int, d_scores: int, stream: int = 0):
        _lib.check(self._L.gj_score_plain_i32_device(self.handle, C.c_void_p(d_samples),
                                                     C.c_int64(row_stride), C.c_int64(S),
                                                     C.c_void_p(d_scores), C.c_void_p(stream)))

    def score_incremental_device(self, d_base: int, d_offsets: int, d_ids: int, d_vals: int, S: int,
                                 d_scores: int, stream: int = 0):
        _lib.check(self._L.gj_score_incremental_device(
            self.handle, C.c_void_p(d_base), C.c_void_p(d_offsets), C.c_void_p(d_ids),
            C.c_void_p(d_vals), C.c_int64(S), C.c_void_p(d_scores), C.c_void_p(stream)))
