"""TEST INFRASTRUCTURE: the numeric bar for f / g values (north_star: 1e-12 RELATIVE).

    |a - b| <= RTOL * |b| + FLOOR_ULPS * eps * T(x)

The first term is the acceptance criterion as stated.  The second is the explicit absolute floor:
every objective here is a SUM whose result can be arbitrarily smaller than its terms (Ackley at its
minimum: -20 - e + 20 + e = 4.4e-16), and no two correctly rounded evaluation orders agree to a
relative 1e-12 of such a result -- they agree to a few ulp of the LARGEST MAGNITUDE THE SUM PASSES
THROUGH.  T(x) is that magnitude, per point, from the definition of the functor (sum of the absolute
values of the terms of the final summation); FLOOR_ULPS = 4.  For |f| >~ T * 1e-3 the relative
term dominates and the bar is the plain 1e-12.

`record()` keeps the worst observed relative error and ulp distance per functor; conftest.py prints
the table at the end of the session (and the GPU suite writes it to gpurun_out/ when that exists).
"""
import numpy as np

RTOL = 1e-12
FLOOR_ULPS = 4.0
EPS = np.finfo(np.float64).eps          # 2.22e-16 = ulp(1.0)

WORST = {}      # functor -> dict(rel=, ulp=, n=, floor_hits=)


def term_scale(functor, x):
    """T(x) per point; x is (n, d)."""
    x = np.asarray(x, np.float64)
    n, d = x.shape
    if functor == "rastrigin":           # 10 d + sum(x^2 - 10 cos(2 pi x))
        return 10.0 * d + np.sum(x * x, 1) + 10.0 * d
    if functor == "ackley":              # -20 exp(..) - exp(..) + 20 + e
        return np.full(n, 2.0 * (20.0 + np.e))
    if functor == "styblinski_tang":     # 0.5 sum(x^4 - 16 x^2 + 5 x)
        return 0.5 * np.sum(x ** 4 + 16.0 * x * x + 5.0 * np.abs(x), 1)
    if functor == "eggholder":           # -(x1+47) sin(..) - x0 sin(..)
        return np.abs(x[:, 1] + 47.0) + np.abs(x[:, 0])
    if functor == "rosenbrock":
        return np.sum(100.0 * (np.abs(x[:, 1:]) + x[:, :-1] ** 2) ** 2 + (1.0 + np.abs(x[:, :-1])) ** 2, 1)
    if functor == "cattle_feed":
        return np.abs(x) @ np.array([24.55, 26.75, 39.0, 40.5])
    if functor == "sphere":
        return np.sum(x * x, 1)
    raise KeyError(functor)


def ulp_distance(a, b):
    """|a - b| in units of the spacing of doubles at max(|a|, |b|) (finite entries)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    sp = np.spacing(np.maximum(np.abs(a), np.abs(b)))
    return np.abs(a - b) / sp


def record(functor, a, b, floor):
    fin = np.isfinite(b)
    if not fin.any():
        return
    a, b, floor = a[fin], b[fin], np.broadcast_to(floor, fin.shape)[fin]
    err = np.abs(a - b)
    nz = np.abs(b) > 0
    rel = float(np.max(err[nz] / np.abs(b[nz]))) if nz.any() else 0.0
    ulp = float(np.max(ulp_distance(a, b)))
    w = WORST.setdefault(functor, dict(rel=0.0, ulp=0.0, n=0, floor_hits=0, worst_over_floor=0.0))
    w["rel"] = max(w["rel"], rel)
    w["ulp"] = max(w["ulp"], ulp)
    w["n"] += int(fin.sum())
    over = err > RTOL * np.abs(b)                 # points that needed the floor
    w["floor_hits"] += int(over.sum())
    if over.any():
        w["worst_over_floor"] = max(w["worst_over_floor"], float(np.max(err[over] / floor[over])))


def assert_close(a, b, functor, x):
    """a: values under test, b: reference values, x: (n, d) points they were evaluated at."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isinf(a), np.isinf(b)) and np.array_equal(np.isnan(a), np.isnan(b))
    fin = np.isfinite(b)
    if not fin.any():
        return
    floor = FLOOR_ULPS * EPS * term_scale(functor, x)
    record(functor, a, b, floor)
    bad = fin & (np.abs(a - b) > RTOL * np.abs(b) + floor)
    if bad.any():
        i = int(np.flatnonzero(bad)[0])
        raise AssertionError("%s: %d of %d values outside 1e-12 relative + %g-ulp floor; first: x=%r got %r want %r"
                             % (functor, int(bad.sum()), int(fin.sum()), FLOOR_ULPS, x[i].tolist(), a[i], b[i]))


def report():
    lines = ["worst observed f error per functor (rtol %g, floor %g ulp of the largest term)" % (RTOL, FLOOR_ULPS)]
    for k in sorted(WORST):
        w = WORST[k]
        lines.append("  %-16s values %9d  max rel err %.3e  max ulp distance %10.1f  needed the floor: %d (worst = %.2f of it)"
                     % (k, w["n"], w["rel"], w["ulp"], w["floor_hits"], w["worst_over_floor"]))
    return "\n".join(lines)
