"""Registered objectives and constraint sets.

Each entry pairs the numpy callable a user of the reference passes to ``shgo(func, ...)`` (the
reference evaluates it one vertex at a time, shgo/_shgo_lib/_vertex.py:338-352) with the id of the
device functor (shgo_b200/csrc/functors.cuh) that evaluates it in batch on the GPU.  Passing one
of these callables (or ``scipy.optimize.rosen``) to :func:`shgo_b200.shgo` selects the fused
generate->evaluate kernels; any other callable takes the generic paths described in
:mod:`shgo_b200.complex`.
"""
import numpy as np

# ids must match include/shgo_b200.h
F_ROSENBROCK, F_RASTRIGIN, F_ACKLEY, F_STYBLINSKI_TANG, F_EGGHOLDER, F_CATTLE_FEED, F_SPHERE = range(7)
G_NONE, G_CATTLE_FEED, G_SUM_LE_6 = range(3)


def _register(fid):
    def deco(fn):
        fn.shgo_b200_functor = fid
        return fn
    return deco


def _register_g(gid, pos):
    def deco(fn):
        fn.shgo_b200_constraint = (gid, pos)
        return fn
    return deco


@_register(F_ROSENBROCK)
def rosenbrock(x):
    """scipy.optimize.rosen"""
    x = np.asarray(x)
    return np.sum(100.0 * (x[1:] - x[:-1] ** 2.0) ** 2.0 + (1 - x[:-1]) ** 2.0, axis=0)


@_register(F_RASTRIGIN)
def rastrigin(x):
    return 10 * len(x) + np.sum(x ** 2 - 10 * np.cos(2 * np.pi * x))


@_register(F_ACKLEY)
def ackley(x):
    d = len(x)
    return (-20 * np.exp(-0.2 * np.sqrt(np.sum(x ** 2) / d))
            - np.exp(np.sum(np.cos(2 * np.pi * x)) / d) + 20 + np.e)


@_register(F_STYBLINSKI_TANG)
def styblinski_tang(x):
    return 0.5 * np.sum(x ** 4 - 16 * x ** 2 + 5 * x)


@_register(F_EGGHOLDER)
def eggholder(x):
    """reference shgo/_shgo.py:329-333"""
    return (-(x[1] + 47.0) * np.sin(np.sqrt(abs(x[0] / 2.0 + (x[1] + 47.0))))
            - x[0] * np.sin(np.sqrt(abs(x[0] - (x[1] + 47.0)))))


@_register(F_CATTLE_FEED)
def cattle_feed(x):
    """reference shgo/_shgo.py:408-409"""
    return 24.55 * x[0] + 26.75 * x[1] + 39 * x[2] + 40.50 * x[3]


@_register_g(G_CATTLE_FEED, 0)
def cattle_feed_g1(x):
    """reference shgo/_shgo.py:411-412"""
    return 2.3 * x[0] + 5.6 * x[1] + 11.1 * x[2] + 1.3 * x[3] - 5


@_register_g(G_CATTLE_FEED, 1)
def cattle_feed_g2(x):
    """reference shgo/_shgo.py:414-418"""
    return (12 * x[0] + 11.9 * x[1] + 41.8 * x[2] + 52.1 * x[3] - 21
            - 1.645 * np.sqrt(0.28 * x[0] ** 2 + 0.19 * x[1] ** 2
                              + 20.5 * x[2] ** 2 + 0.62 * x[3] ** 2))


def cattle_feed_h1(x):
    """reference shgo/_shgo.py:420-421 (equality; ignored by the sampling stage)"""
    return x[0] + x[1] + x[2] + x[3] - 1


@_register(F_SPHERE)
def sphere(x):
    r = x[0] ** 2
    for i in range(1, len(x)):
        r = r + x[i] ** 2
    return r


@_register_g(G_SUM_LE_6, 0)
def sum_le_6(x):
    """reference shgo/tests/test__shgo.py:44-45"""
    return -(np.sum(x, axis=0) - 6.0)


_GSET_SIZE = {G_CATTLE_FEED: 2, G_SUM_LE_6: 1}


def _unwrap(func):
    """SHGO wraps func in scipy's _FunctionWrapper(f, args); look through it."""
    inner = getattr(func, "f", None)
    args = getattr(func, "args", None)
    if inner is not None and callable(inner) and (args is None or len(args) == 0):
        return inner
    return func


def functor_id(func):
    """Device functor id of an objective callable, or None."""
    func = _unwrap(func)
    fid = getattr(func, "shgo_b200_functor", None)
    if fid is not None:
        return fid
    try:
        from scipy.optimize import rosen
        if func is rosen:
            return F_ROSENBROCK
    except Exception:  # pragma: no cover
        pass
    return None


MAX_FUSED_DIM = 32      # kMaxDim of csrc/common.cuh


def supports(fid, gid, d):
    """Does the device functor `fid` with constraint set `gid` exist for dimension d?  Mirrors the
    dim_ok() rules of csrc/functors.cuh; DeviceComplex sends anything else down the generic route
    (the user's own callable, evaluated per vertex like the reference does)."""
    if fid is None or gid is None or not 1 <= d <= MAX_FUSED_DIM:
        return False
    if fid == F_ROSENBROCK:
        return d >= 2 and gid == G_NONE
    if fid in (F_RASTRIGIN, F_ACKLEY, F_STYBLINSKI_TANG):
        return gid == G_NONE
    if fid == F_EGGHOLDER:
        return d == 2 and gid == G_NONE
    if fid == F_CATTLE_FEED:
        return d == 4 and gid in (G_NONE, G_CATTLE_FEED)
    if fid == F_SPHERE:
        return gid == G_NONE or (gid == G_SUM_LE_6 and d < 8)
    return False


def constraint_set_id(g_cons, g_args=None):
    """Device constraint-set id of a tuple of `ineq` callables: G_NONE for no constraints, the id
    when the tuple is exactly one registered set in order, else None."""
    if not g_cons:
        return G_NONE
    if g_args is not None and any(len(a) for a in g_args):
        return None
    tags = [getattr(g, "shgo_b200_constraint", None) for g in g_cons]
    if any(t is None for t in tags):
        return None
    gid = tags[0][0]
    if [t for t in tags] == [(gid, i) for i in range(_GSET_SIZE[gid])]:
        return gid
    return None


# objective functor compiled together with each constraint set (csrc/eval.cu dispatch table): lets a
# caller that only wants the feasibility flags of a registered set run the fused kernel
PAIRED_OBJECTIVE = {G_CATTLE_FEED: F_CATTLE_FEED, G_SUM_LE_6: F_SPHERE}
