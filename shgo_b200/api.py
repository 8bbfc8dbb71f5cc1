"""``shgo_b200.shgo`` / ``SHGO``: the reference's public API on the device complex.

Nothing of the reference's driver is re-implemented: its ``SHGO`` class (option parsing, stopping
criteria, iteration loop, scipy local minimisation, result assembly -- shgo/_shgo.py:490-1399) is
used as is, from whichever driver package is importable (the reference package ``shgo`` if present,
else scipy's vendored copy ``scipy.optimize._shgo``, a near-identical file).  What changes is the
object behind ``SHGO.HC``: the driver module binds ``Complex`` by name at import
(shgo/_shgo.py:17), so :func:`install` rebinds that name to :class:`DeviceComplex`; with
``sampling_method='sobol'`` the points additionally come from the device generator through the
driver's official custom-sampling seam (``sampling_method`` callable f(n, d), _shgo.py:226-228,
1434-1444), bit-identical to scipy's engine.

    from shgo_b200 import shgo
    res = shgo(func, bounds, constraints=cons, n=2**20, iters=1, sampling_method='sobol')
"""
import contextlib
import importlib
import threading
import types

import numpy as np

from . import complex as _cx
from . import hotpath as hp

_DRIVERS = ("shgo._shgo", "scipy.optimize._shgo")
_COMPLEX_MODULES = {"shgo._shgo": "shgo._shgo_lib._complex",
                    "scipy.optimize._shgo": "scipy.optimize._shgo_lib._complex"}


def driver_module(name=None):
    """The host driver module whose SHGO class is used (first importable of `_DRIVERS`)."""
    last = None
    for mod in ((name,) if name else _DRIVERS):
        try:
            return importlib.import_module(mod)
        except ImportError as e:  # pragma: no cover
            last = e
    raise ImportError("no SHGO driver found (tried %s): %s" % (", ".join(_DRIVERS), last))


_LOCK = threading.RLock()     # install / uninstall edit module globals of the driver


def _device_lcb(saved):
    """construct_lcb_simplicial (_shgo.py:1246-1270) served from the device when the complex is a
    DeviceComplex: no neighbour objects are created; any other complex gets the driver's own."""
    def construct_lcb_simplicial(self, v_min):
        if isinstance(self.HC, _cx.DeviceComplex):
            cbounds = self.HC.local_bounds(v_min, self.bounds)
            if self.disp:
                import logging
                logging.info(f'cbounds found for v_min.x_a = {v_min.x_a}')
                logging.info(f'cbounds = {cbounds}')
            return cbounds
        return saved(self, v_min)

    construct_lcb_simplicial.__doc__ = saved.__doc__
    return construct_lcb_simplicial


# pools of at least this many vertices are filtered / ranked on the device; below,
# the driver's own loop over the handful of materialised vertices is used unchanged
DEVICE_POOL_RANK_MIN = 2048


def _device_minimizers(saved):
    """SHGO.minimizers (_shgo.py:1109-1157) with the LMC exclusion and the f-ordering done on the
    device for large pools; produces the same attributes as the driver's loop + sort_min_pool.  Among
    exactly equal f values the order is the cache order (np.argsort promises none)."""
    def minimizers(self):
        HC = self.HC
        if not (isinstance(HC, _cx.DeviceComplex) and len(HC._pool_cache_order) >= DEVICE_POOL_RANK_MIN):
            return saved(self)
        xl = list(self.LMC.xl_maps) if len(self.LMC.xl_maps) > 0 else []
        ranked = HC.ranked_pool(xl)
        kept = set(ranked.tolist())
        by_index = HC.V._by_index
        in_cache_order = [by_index[i] for i in HC._pool_cache_order.tolist() if i in kept]
        self.X_min = np.array([v.x_a for v in in_cache_order])
        self.X_min_cache = {tuple(v.x_a): v.x for v in in_cache_order}
        pool = np.empty(len(ranked), dtype=object)
        pool[:] = [by_index[i] for i in ranked.tolist()]
        self.minimizer_pool = pool
        self.minimizer_pool_F = np.array([v.f for v in pool], dtype=float)
        return self.X_min

    minimizers.__doc__ = saved.__doc__
    return minimizers


def _device_subspace(saved):
    """SHGO.sampling_subspace (_shgo.py:1452-1463) on the device when the constraints allow it"""
    def sampling_subspace(self):
        C = None
        if isinstance(getattr(self, "HC", None), _cx.DeviceComplex):
            C = _cx.device_sampling_subspace(self.C, self.g_cons, self.g_args)
        if C is None:
            return saved(self)
        self.C = C
        if self.C.size == 0:
            self.res.message = ('No sampling point found within the '
                                + 'feasible set. Increasing sampling '
                                + 'size.')
            if self.disp:
                import logging
                logging.info(self.res.message)

    sampling_subspace.__doc__ = saved.__doc__
    return sampling_subspace


_PATCHED_METHODS = (("construct_lcb_simplicial", "_shgo_b200_saved_lcb", _device_lcb),
                    ("minimizers", "_shgo_b200_saved_minimizers", _device_minimizers),
                    ("sampling_subspace", "_shgo_b200_saved_subspace", _device_subspace))


def install(driver=None):
    """Rebind ``Complex`` in the driver module to DeviceComplex (what INTEGRATION.md asks a
    reference maintainer to do in one line).  Returns the driver module.  This is a process-wide
    patch of the driver module: every later ``shgo()`` of that driver runs on the device until
    :func:`uninstall`.  ``shgo_b200.shgo`` / ``shgo_b200.SHGO`` scope it to their own call."""
    with _LOCK:
        m = driver_module(driver)
        cm = importlib.import_module(_COMPLEX_MODULES[m.__name__])
        _cx._ORIGINAL_COMPLEX.setdefault(cm.__name__, cm.Complex)
        if m.Complex is not _cx.DeviceComplex:
            m._shgo_b200_saved_complex = m.Complex
            m.Complex = _cx.DeviceComplex
        _cx.DeviceComplex.scipy_constraint_rule = m.__name__.startswith("scipy.")
        for name, slot, make in _PATCHED_METHODS:
            if not hasattr(m.SHGO, slot):
                saved = getattr(m.SHGO, name)
                setattr(m.SHGO, slot, saved)
                setattr(m.SHGO, name, make(saved))
        return m


def uninstall(driver=None):
    with _LOCK:
        m = driver_module(driver)
        saved = getattr(m, "_shgo_b200_saved_complex", None)
        if saved is not None:
            m.Complex = saved
            del m._shgo_b200_saved_complex
        for name, slot, _ in _PATCHED_METHODS:
            if hasattr(m.SHGO, slot):
                setattr(m.SHGO, name, getattr(m.SHGO, slot))
                delattr(m.SHGO, slot)
        _cx.DeviceComplex.scipy_constraint_rule = False
        return m


@contextlib.contextmanager
def _installed(driver=None):
    """install() for the duration of the block only (held under a lock: the patch is a module
    global of the driver, two threads must not interleave install / uninstall)."""
    with _LOCK:
        m = driver_module(driver)
        already = m.Complex is _cx.DeviceComplex
        install(driver)
        try:
            yield m
        finally:
            if not already:
                uninstall(driver)


class DeviceSobolSampler:
    """``sampling_method`` callable f(n, d) backed by the device generator.

    Mirrors the reference's engine semantics (qmc.Sobol(d, scramble=False, seed=0), _shgo.py:703-711):
    the state persists across calls, every call returns the NEXT n points of the sequence in [0,1)^d
    (the driver scales them to the bounds itself, _shgo.py:1446-1449).  The scaled points of indices
    [0, total) are also what DeviceComplex regenerates inside the fused evaluate kernel."""

    def __init__(self, dim, bounds, device=None):
        self.dim = dim
        self.unit = hp.SobolTables([(0.0, 1.0)] * dim, device)
        self.tab = hp.SobolTables(bounds, device)
        self.total = 0
        self.host = None          # scaled points [0, total) as the driver will see them

    def __call__(self, n, d):
        assert d == self.dim
        u = hp.sobol_points(self.unit, self.total, n, "aos").cpu().numpy()
        if self.total == 0:
            self.host = hp.sobol_points(self.tab, 0, n, "aos").cpu().numpy()
        else:
            self.host = None      # several draws: the complex uploads the points it is given
        self.total += n
        return u


def _device_sampler(n, bounds, sampling_method):
    """(n, sampling_method, sampler) with 'sobol' replaced by the device generator."""
    if sampling_method != "sobol":
        return n, sampling_method, None
    # the driver rounds n up to a power of two for its own Sobol engine (_shgo.py:696-697);
    # a callable sampler bypasses that, so do it here
    nn = 100 if n is None else n
    n = int(2 ** np.ceil(np.log2(nn)))
    b = np.array(bounds, float) if not hasattr(bounds, "lb") else np.stack([bounds.lb, bounds.ub], 1)
    b = b.copy()
    b[~np.isfinite(b[:, 0]), 0] = -1e50
    b[~np.isfinite(b[:, 1]), 1] = 1e50
    sampler = DeviceSobolSampler(b.shape[0], b)
    return n, sampler, sampler


# Set to False to keep options['local_iter'] calls on the closed-form lattice (fast; the first local
# search may then start from a different pool vertex of the same generation than the reference's).
EXACT_CACHE_ORDER = True


def _needs_cache_order(options, sampling_method):
    """options['local_iter'] starts its local searches from X_min[0], the pool vertex the reference's
    cache created first (_shgo.py:1177); on the simplicial complex that order has a closed form only
    across generations, so such a call gets its graph from the reference's generator."""
    return (EXACT_CACHE_ORDER and bool(options and options.get("local_iter"))
            and sampling_method == "simplicial")


def SHGO(func, bounds, args=(), constraints=None, n=None, iters=None, callback=None,
         minimizer_kwargs=None, options=None, sampling_method="simplicial", workers=1, driver=None):
    """The driver's SHGO object (same constructor as shgo/_shgo.py:491) with ``HC`` on the device.
    Use it like the reference's: ``s.iterate_all()``, ``s.iterate()``, ``s.res`` ..."""
    # The patch of the driver module lasts for the construction only (HC is built inside
    # SHGO.__init__, _shgo.py:680-684): a later plain `scipy.optimize.shgo` / reference `shgo` call
    # in the same process is not rerouted.  What the object needs afterwards is bound to it.
    with _installed(driver) as m:
        n, sampling_method, sampler = _device_sampler(n, bounds, sampling_method)
        # picked up by the DeviceComplex the driver constructs
        _cx.set_pending_sampler(sampler, _needs_cache_order(options, sampling_method))
        try:
            obj = m.SHGO(func, bounds, args=args, constraints=constraints, n=n, iters=iters, callback=callback,
                         minimizer_kwargs=minimizer_kwargs, options=options, sampling_method=sampling_method,
                         workers=workers)
        finally:
            _cx.set_pending_sampler(None)
        saved = {name: getattr(m.SHGO, slot, getattr(m.SHGO, name)) for name, slot, _ in _PATCHED_METHODS}
    for name, _, make in _PATCHED_METHODS:
        setattr(obj, name, types.MethodType(make(saved[name]), obj))
    obj.HC.scipy_constraint_rule = m.__name__.startswith("scipy.")
    return obj


def shgo(func, bounds, args=(), constraints=None, n=100, iters=1, callback=None,
         minimizer_kwargs=None, options=None, sampling_method="simplicial", *, workers=1, driver=None):
    """Drop-in for the reference's ``shgo()`` (shgo/_shgo.py:22-487): same arguments, same
    OptimizeResult fields.  It IS the driver's own ``shgo()``, run while ``Complex`` is bound to
    DeviceComplex; only ``sampling_method='sobol'`` is rerouted to the device generator."""
    with _installed(driver) as m:
        n, sampling_method, sampler = _device_sampler(n, bounds, sampling_method)
        _cx.set_pending_sampler(sampler, _needs_cache_order(options, sampling_method))
        try:
            return m.shgo(func, bounds, args=args, constraints=constraints, n=n, iters=iters,
                          callback=callback, minimizer_kwargs=minimizer_kwargs, options=options,
                          sampling_method=sampling_method, workers=workers)
        finally:
            _cx.set_pending_sampler(None)
