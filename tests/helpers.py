"""Shared generators for the parity tests: random candidates and random move lists built
with the ORACLE's mover (explicit random choices), so GPU and oracle see identical sets."""
import numpy as np

from greyjack_b200 import instances as inst

SOFT_RTOL = 1e-12   # stated float tolerance for tree-reduced distance sums (TSP soft level)


def random_samples(spec, S, rng, wild=True):
    """S candidate vectors; `wild` adds fractional / out-of-bound values for the decoder."""
    lo, hi = spec.lower_bounds, spec.upper_bounds
    x = rng.integers(lo.astype(np.int64), hi.astype(np.int64) + 1, size=(S, spec.n_vars)).astype(np.float64)
    if wild and S > 2:
        k = max(1, S // 4)
        x[:k] += rng.uniform(-0.5, 0.5, size=(k, spec.n_vars))        # rint paths (incl. near ties)
        x[k:2 * k] += rng.integers(-3, 4, size=(k, spec.n_vars)) * (hi - lo + 1)  # clamp paths
        x[2 * k, :] = x[2 * k, :] + 0.5                               # exact ties -> ceil
    return x


def permutation_samples(spec, S, rng):
    """Feasible-looking candidates: TSP permutations / VRP permutations with random vehicles."""
    out = np.empty((S, spec.n_vars))
    for j in range(S):
        if spec.kind == inst.TSP:
            out[j] = rng.permutation(spec.n_vars) + 1
        elif spec.kind == inst.NQUEENS:
            out[j] = rng.permutation(spec.n_vars)
        else:
            n = spec.n_vars // 2
            out[j, 0::2] = rng.integers(0, spec.n_vehicles, size=n)
            out[j, 1::2] = rng.permutation(n) + spec.n_depots
    return out


def random_moves(op, spec, base, K, rng, kinds=(0, 1, 2, 3, 4, 5), max_k=4, incremental=True):
    """K random moves from `base` through the oracle mover.  Returns (deltas, kinds) where
    deltas[j] is a list of (var_id, value) after fix_deltas (tabu_search_base.rs:124-132)."""
    names = list(spec.groups.keys())
    deltas, mk = [], []
    while len(deltas) < K:
        kind = int(rng.choice(kinds))
        g = np.asarray(spec.groups[names[int(rng.integers(len(names)))]], dtype=np.int32)
        glen = len(g)
        res = None
        if kind == 0:
            k = int(rng.integers(1, max_k + 1))
            if glen < k:
                continue
            ch = rng.choice(glen, size=k, replace=False)
            cols = g[ch]
            nv = spec.lower_bounds[cols] + rng.random(k) * (spec.upper_bounds[cols] - spec.lower_bounds[cols])
            res = op.move_change(base, g, ch, nv, incremental)
        elif kind == 1:
            k = int(rng.integers(2, max_k + 1))
            if glen < k:
                continue
            res = op.move_swap(base, g, rng.choice(glen, size=k, replace=False), incremental)
        elif kind == 2:
            k = int(rng.integers(2, max_k + 1))
            if glen < 3:
                continue
            k = min(k, glen - 1)
            res = op.move_swap_edges(base, g, rng.choice(glen - 1, size=k, replace=False), incremental)
        elif kind == 3:
            cnt = int(rng.integers(3, 7))
            if glen <= cnt:
                continue
            start = int(rng.integers(0, glen - cnt))
            res = op.move_scramble(base, g, start, rng.permutation(cnt), incremental)
        elif kind == 4:
            a, b = rng.choice(glen, size=2, replace=False)
            res = op.move_insertion(base, g, int(a), int(b), incremental)
        else:
            a, b = rng.choice(glen, size=2, replace=False)
            res = op.move_inverse(base, g, int(a), int(b), incremental)
        if res is None:
            continue
        cols, vals = res
        if incremental:
            vals = op.fix_deltas(cols, vals)
            deltas.append([(int(c), float(v)) for c, v in zip(cols, vals)])
        else:
            deltas.append(op.fix_variables(vals, cols))
        mk.append(kind)
    return deltas, mk


def assert_scores_match(got, want, spec, soft_exact=False):
    """Integer levels bit-exact; the float (distance) level within SOFT_RTOL unless the
    kernel keeps the reference's summation order (VRP), where it must be bit-exact too."""
    got = np.asarray(got); want = np.asarray(want)
    assert got.shape == want.shape
    L = spec.levels
    int_levels = list(range(L - 1)) if L > 1 else [0]
    for l in int_levels:
        assert np.array_equal(got[:, l], want[:, l]), f"integer level {l} differs"
    if L > 1:
        if soft_exact:
            assert np.array_equal(got[:, L - 1], want[:, L - 1]), "soft level not bit-exact"
        else:
            np.testing.assert_allclose(got[:, L - 1], want[:, L - 1], rtol=SOFT_RTOL, atol=0.0)
