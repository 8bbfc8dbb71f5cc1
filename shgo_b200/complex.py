"""DeviceComplex -- the CUDA-backed replacement for the reference's ``Complex`` at the ``SHGO.HC`` seam.

The reference constructs ``self.HC = Complex(dim, domain, sfield, sfield_args, symmetry, constraints,
workers)`` in exactly one place (shgo/_shgo.py:680-684) and afterwards touches only a small duck-typed
surface:

    HC.refine(n) / HC.refine_all()            grow the hypercube complex        (_shgo.py:1016,1019)
    HC.vf_to_vv(points, simplices)            ingest qhull's vertex-face mesh   (_shgo.py:1090)
    HC.V.process_pools()                      evaluate g, f; classify           (_shgo.py:1044,1099)
    HC.V.nfev, HC.V.size()                                                      (_shgo.py:834,1017)
    HC.V.cache (ordered mapping coordinate tuple -> vertex), HC.V[x]            (_shgo.py:867,1029,1115)
    vertex .x .x_a .f .nn .minimiser() .star()                                  (_shgo.py:868,1030,1124,1264)

DeviceComplex offers that surface and keeps everything that scales with the number of sample points
on the GPU: coordinates, f, feasibility, adjacency (explicit CSR or the implicit lattice rule), the
star test, pool compaction and the arg-min.  Python vertex objects are materialised only for what
the driver actually looks at: the minimiser pool, the arg-min vertex and, on demand, their stars.

Graph sources
    'lattice'    whole generations of the simplicial complex (refine_all, or refine(2^d+1) from
                 scratch): closed form, vertices generated on the device, adjacency implicit.
    'mesh'       Sobol / custom sampling: qhull's simplices (host) -> device CSR (first-three-slot
                 edges, exactly what Complex.vf_to_vv builds, _complex.py:1053-1074).
    'hostgraph'  partial generations (refine(n) stopping mid-generator, _complex.py:474-512) and
                 `symmetry`: sequential by construction, no closed form.  The
                 driver-side Python ``Complex`` is run field-less as a pure graph generator, like
                 qhull for the Sobol path; its vertices/edges are uploaded as points + CSR and
                 evaluated / classified on the device like a mesh.

Objective evaluation
    registered functor (shgo_b200.functors, scipy.optimize.rosen)  -> fused device kernels
    callable marked with @torch_vectorized                         -> called once on the (d, n) CUDA
                                                                      tensor of coordinates
    any other Python callable                                      -> it is user code: called per
                                                                      vertex on the host exactly like
                                                                      VertexCacheField.compute_sfield
                                                                      (_vertex.py:338-352); the values
                                                                      are uploaded and everything else
                                                                      (mask, star test, pool) stays on
                                                                      the device.
There is no CPU implementation of the classification: without the CUDA library these classes raise.
"""
import collections
import importlib
import threading

import numpy as np
import torch

from . import functors as F
from . import hotpath as hp
from . import lattice as lat


_PENDING = threading.local()
# meshes with at least this many vertices are ALSO kept in a Z-order (Morton) numbering of the points,
# used by the star test only: qhull numbers the vertices like the (Sobol-ordered) points, so neighbour
# ids are scattered over the whole range and the gathers of f miss every cache
MORTON_MIN_VERTICES = 1 << 15
LATTICE_MAX_DIM = 12     # closed-form star / local-bounds kernels (csrc/lattice.cu, bounds.cu); above: generator replay


def set_pending_sampler(sampler, exact_cache_order=False):
    """Hand the device Sobol sampler of an upcoming driver call to the DeviceComplex that call will
    construct (the driver builds `HC` itself, _shgo.py:680-684).  exact_cache_order: the call uses
    options['local_iter'], whose result depends on the order in which the reference's cache created
    the vertices of one generation -- that complex takes its graph from the reference's generator."""
    _PENDING.sampler = sampler
    _PENDING.exact_cache_order = bool(exact_cache_order)


def torch_vectorized(fn):
    """Mark an objective / constraint as torch-vectorised: fn(X, *args) with X a (d, n) float64
    CUDA tensor (X[i] is the i-th coordinate of every point) returning an (n,) tensor."""
    fn.shgo_b200_torch = True
    return fn


def _is_torch_vectorized(fn):
    return bool(getattr(F._unwrap(fn), "shgo_b200_torch", False))


def _split_wrapper(fn):
    """(callable, args) of a possibly _FunctionWrapper-wrapped objective."""
    inner = getattr(fn, "f", None)
    if inner is not None and callable(inner) and hasattr(fn, "args"):
        return inner, tuple(fn.args or ())
    return fn, ()


class DeviceVertex:
    """Host-side handle of one vertex (the reference's VertexScalarField, _vertex.py:75-163)."""

    __slots__ = ("x", "x_a", "f", "index", "feasible", "_cache", "_is_min", "__weakref__")

    def __init__(self, cache, index, x_a, f, is_min, feasible=True):
        self._cache = cache
        self.index = int(index)
        self.x_a = np.array(x_a, dtype=np.float64)
        self.x = tuple(self.x_a.tolist())
        self.f = float(f)
        self._is_min = bool(is_min)
        self.feasible = bool(feasible)

    def __hash__(self):
        return hash(self.x)

    def __eq__(self, other):
        return self is other

    def __repr__(self):
        return "DeviceVertex(%d, x=%r, f=%r)" % (self.index, self.x, self.f)

    @property
    def nn(self):
        """Neighbour vertices, materialised on demand (a set, like the reference's).  Between a
        refinement and the next process_pools() the device arrays are being rebuilt and the set is
        empty; the only reader in that window is the dead `v_near` computation of
        SHGO.iterate_hypercube (_shgo.py:1027-1032, its consumer is commented out)."""
        return self._cache._neighbours(self)

    def minimiser(self):
        return self._is_min

    def maximiser(self):  # unused by shgo(); kept for interface parity (_vertex.py:153-162)
        return all(self.f > v.f for v in self.nn)

    def star(self):
        """st(v) = nn + {v}.  In the reference this INSERTS v into its own nn (_vertex.py:57-72),
        which makes it fail the strict star test at its next re-classification; the cache
        records that."""
        self._cache._mark_started(self)
        st = set(self.nn)
        st.add(self)
        return st

    def connect(self, v):
        raise NotImplementedError("the device complex is not edited edge by edge")

    disconnect = connect


class _NoPool:
    """Stands in for VertexCacheField._mapwrapper (scipy._lib._util.MapWrapper): the driver only
    ever calls its __exit__ (shgo/_shgo.py:801-802; unconditionally in scipy's copy).  `workers`
    parallelises per-vertex Python calls in the reference; batched device evaluation replaces it."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return None

    def close(self):
        pass

    terminate = join = close


class DeviceVertexCache:
    """The ``HC.V`` object: ordered, lazily materialised view of the device-resident vertices."""

    def __init__(self, complex_):
        self._c = complex_
        self._mapwrapper = _NoPool()
        self.cache = collections.OrderedDict()   # coordinate tuple -> DeviceVertex (materialised)
        self._by_index = {}
        self.nfev = 0
        self._started = set()                    # coordinates on which star() was called
        self._excluded = set()                   # started vertices re-classified since (never minima)
        self._persistent = {}                    # coordinate tuple -> last handle of pool members
        self._unlisted = {}                      # materialised stars (not part of `cache`)

    # -- reference interface ---------------------------------------------------------------
    def __iter__(self):
        return iter(list(self.cache.values()))

    def __getitem__(self, x):
        x = tuple(float(v) for v in x)
        try:
            return self.cache[x]
        except KeyError:
            pass
        if x in self._unlisted:
            return self._unlisted[x]
        if self._c._dirty:
            # graph rebuilt, f not yet recomputed: hand out the handle from the last classification
            v = self._persistent.get(x)
            if v is None:
                v = DeviceVertex(self, -1, x, float("nan"), False)
                self._persistent[x] = v
            return v
        v = self._c._materialise_by_coordinate(x)
        if v is None:
            raise KeyError(
                "%r is not a vertex of the device complex (inserting free-standing vertices is "
                "not supported)" % (x,))
        return v

    def size(self):
        return self._c._size()

    def process_pools(self):
        self._c._process()

    # -- internals -------------------------------------------------------------------------------
    def _reset_view(self):
        self.cache = collections.OrderedDict()
        self._unlisted = {}
        self._by_index = {}

    def _add(self, index, x_a, f, is_min, feasible=True, listed=True):
        """`listed` vertices appear in `cache` (what the driver iterates: pool + arg-min); stars
        materialised later on demand are reachable through V[x] and .nn only, so that reading
        v.nn inside the driver's `for x in HC.V.cache` loop (disp=True, _shgo.py:1136-1139) does not
        mutate the mapping being iterated."""
        index = int(index)
        v = self._by_index.get(index)
        if v is None:
            v = DeviceVertex(self, index, x_a, f, is_min, feasible)
            self._by_index[index] = v
            if listed:
                self.cache[v.x] = v
            else:
                self._unlisted[v.x] = v
            if is_min:
                self._persistent[v.x] = v
        return v

    def _neighbours(self, v):
        if self._c._dirty or v.index < 0 or self._by_index.get(v.index) is not v:
            return set()
        ids = self._c._neighbour_ids(v.index)
        out = set()
        missing = [i for i in ids.tolist() if i not in self._by_index]
        if missing:
            self._c._materialise(np.asarray(missing, np.int64), listed=False)
        for i in ids.tolist():
            out.add(self._by_index[i])
        return out

    def _mark_started(self, v):
        self._started.add(self._c._stable_key(v))
        # The reference's star() really inserts v into v.nn, and its generators walk nn sets
        # (`for v in vo.nn: v.connect(vc)`): the self-loop changes which edges later refinements
        # create.  The driver-side generator of the hostgraph route must see the same loop.
        ref = self._c._ref
        if ref is not None:
            rv = ref.V.cache.get(v.x)
            if rv is not None:
                rv.nn.add(rv)


class DeviceComplex:
    """Drop-in for ``shgo._shgo_lib._complex.Complex`` (constructor signature _complex.py:116-117)."""

    scipy_constraint_rule = False

    def __init__(self, dim, domain=None, sfield=None, sfield_args=(), symmetry=None,
                 constraints=None, workers=1):
        hp._require_cuda()
        self.dim = int(dim)
        self.domain = domain
        self.bounds = [(0.0, 1.0)] * self.dim if domain is None else [tuple(map(float, b)) for b in domain]
        self.symmetry = symmetry
        self.sfield = sfield
        self.sfield_args = tuple(sfield_args or ())
        self.workers = workers
        # `ineq` constraints only, like Complex.__init__ (_complex.py:136-152)
        self.min_cons = constraints
        self.g_cons, self.g_args = None, None
        if constraints is not None:
            cons = constraints if isinstance(constraints, (tuple, list)) else (constraints,)
            gs, ga = [], []
            for c in cons:
                # reference: cons['type'] == 'ineq' (_complex.py:144); scipy's vendored copy of the
                # driver tests `cons['type'] in ('ineq')`, a substring test that also admits 'eq' --
                # mirrored when that driver is the one being served (set by api.install)
                if c["type"] == "ineq" or (self.scipy_constraint_rule and c["type"] in ("ineq")):
                    gs.append(c["fun"])
                    ga.append(tuple(c.get("args", ())))
            self.g_cons, self.g_args = tuple(gs), tuple(ga)
        self.device = hp._dev()
        self.V = DeviceVertexCache(self)
        self.gen = 0
        self.mode = None            # 'lattice' | 'mesh' | 'hostgraph'
        self._dirty = False
        # evaluation route
        self._fid = F.functor_id(sfield) if not self.sfield_args else None
        self._gid = F.constraint_set_id(self.g_cons, self.g_args)
        if self._fid is not None and not F.supports(self._fid, self._gid, self.dim):
            # constraints not registered, or no device functor for this dimension: generic route
            self._fid = None
        self._torch = (self._fid is None and sfield is not None and _is_torch_vectorized(sfield)
                       and all(_is_torch_vectorized(g) for g in (self.g_cons or ())))
        # state
        self._lat_gen = -1
        self._f = self._feas = self._is_min = None
        self._points_host = None     # (n, d) for mesh / hostgraph
        self._row_ptr = self._col = None
        self._row_ptr_host = self._col_host = None
        self._n = 0
        self._host_values = {}       # host-callable route: coordinates -> (f, feasible)
        self._rank = None            # mesh: insertion rank of every vertex in the reference's cache
        self._mesh_calls = 0
        self._sampler = getattr(_PENDING, "sampler", None)   # consumed: one complex per driver call
        self._exact_cache_order = getattr(_PENDING, "exact_cache_order", False)
        _PENDING.sampler, _PENDING.exact_cache_order = None, False
        self._morton = None          # mesh: (order, row_ptr, col) of the Z-order numbering, see vf_to_vv
        self._ref = None             # driver-side Python Complex used as a graph generator
        self._ref_keys = []
        self._lcb, self._lcb_key, self._x_soa = {}, None, None
        self._epoch = 0              # classifications done so far
        self.argmin_index = -1
        self.pool_indices = np.zeros(0, np.int64)
        self._pool_cache_order = np.zeros(0, np.int64)

    # =====================================================================================
    # graph construction entry points (what SHGO.iterate_hypercube / iterate_delaunay call)
    # =====================================================================================
    def refine_all(self, centroids=True):
        """One whole generation (Complex.refine_all / triangulate, _complex.py:514-533,359)."""
        if (self.mode in (None, "lattice") and self.symmetry is None and not self._exact_cache_order
                and self._enter_lattice(self._lat_gen + 1)):
            return
        self._ensure_hostgraph()
        self._ref.refine_all()
        self._sync_hostgraph()

    def refine(self, n=1):
        """Add ~n vertices (Complex.refine, _complex.py:474-512).  Completing generation 0 from
        scratch is closed form; anything that stops mid-generator has none and is replayed on the
        driver-side Python generator (graph only), then evaluated / classified on the device."""
        if n is None:
            return self.refine_all()
        if (self.mode is None and self.symmetry is None and n == 2 ** self.dim + 1
                and not self._exact_cache_order and self._enter_lattice(0)):
            return
        self._ensure_hostgraph()
        self._ref.refine(n)
        self._sync_hostgraph()

    def vf_to_vv(self, vertices, simplices):
        """Vertex-face mesh -> vertex-vertex graph on the device (Complex.vf_to_vv,
        _complex.py:1053-1074): first-three-slot edges of every simplex, union with the edges of
        earlier calls (the reference's connect() only ever adds)."""
        if self.mode not in (None, "mesh"):
            raise RuntimeError("vf_to_vv on a complex that was grown by refine()")
        self.mode = "mesh"
        pts = np.ascontiguousarray(vertices, dtype=np.float64)
        if pts.ndim == 1:
            pts = pts.reshape(-1, self.dim)
        simp = np.ascontiguousarray(simplices)
        if simp.ndim != 2 or simp.shape[0] == 0:
            simp = np.zeros((0, 2), np.int32)
        simp = simp.astype(np.int32, copy=False)
        # The reference's vertices are keyed by coordinates and connect() only ever adds: a second
        # mesh is merged into the first.  When the new point array extends the old one (qhull in
        # incremental mode: Tri.points grows) indices persist and nothing has to be done; otherwise
        # (1-D meshes and non-incremental qhull hand over the fresh points only) the new points
        # are appended behind the old ones and the simplices re-indexed.
        if self._points_host is not None and self._n > 0:
            old = self._points_host
            if not (pts.shape[0] >= old.shape[0] and np.array_equal(pts[:old.shape[0]], old)):
                simp = (simp + old.shape[0]).astype(np.int32)
                pts = np.ascontiguousarray(np.concatenate([old, pts]))
        n = pts.shape[0]
        # vertices are keyed by coordinate tuple in the reference: coincident points are ONE vertex
        # (points of the unscrambled Sobol sequence are pairwise distinct: no need to look)
        canon = None if self._sampler_tables_if_points_are_sobol(pts) is not None else self._coincident_map(pts)
        if canon is not None:
            simp = canon[simp].astype(np.int32)
        simp_d = torch.from_numpy(np.ascontiguousarray(simp)).to(self.device)
        # insertion order of the reference's vertex cache = order of first appearance while
        # vf_to_vv walks the simplices; vertices of earlier calls keep their (earlier) rank
        fs = hp.first_seen(simp_d, n)
        rank = torch.where(fs >= 0, fs + (self._mesh_calls << 44), torch.full_like(fs, 2 ** 62))
        if self._rank is not None:
            old = torch.full_like(rank, 2 ** 62)
            old[: self._rank.numel()] = self._rank
            rank = torch.minimum(old, rank)
        self._rank = rank
        self._mesh_calls += 1
        if self._row_ptr is not None and self._n > 0:
            # carry the edges of earlier triangulations over as degenerate [a, b, b] triples
            old = hp.csr_edges3(self._row_ptr, self._col)
            new = hp.relabel3(simp_d, n)                      # first three slots, same numbering
            simp_d = torch.cat([old, new], 0)
        self._row_ptr, self._col = hp.csr_from_simplices(simp_d, n)
        self._row_ptr_host = self._col_host = None
        self._points_host = pts
        self._x_soa = None
        self._n = n
        self._dirty = True
        self._morton = None
        if n >= MORTON_MIN_VERTICES:
            # the same graph in a locality-preserving numbering (acceleration structure of the star
            # test; everything the driver sees stays in the mesh's own numbering)
            self._x_soa = torch.from_numpy(pts).to(self.device).t().contiguous()
            order, newid = hp.morton_order(self._x_soa, self.bounds)
            rp_m, col_m = hp.csr_from_simplices(hp.relabel3(simp_d, n, newid), n)
            self._morton = (order, rp_m, col_m)

    # =====================================================================================
    # evaluation + classification  (VertexCacheField.process_pools, _vertex.py:324-328)
    # =====================================================================================
    def _process(self):
        if self.mode is None or not self._dirty:
            return
        if self.mode == "lattice":
            self._process_lattice()
        else:
            self._process_explicit()
        self._dirty = False

    # ---- lattice ---------------------------------------------------------------------------
    def _enter_lattice(self, gen):
        """Switch to generation `gen` of the closed-form complex; False when the closed form does
        not describe what the reference builds for these bounds (fp64 midpoints that depend on the
        split direction make it create near-duplicate vertices, SURVEY.md section 7 hard part 6):
        the caller then replays the reference's own generator (hostgraph) so that the result stays
        identical to the reference's, duplicates included."""
        # limits of the closed-form kernels (lattice.cu / bounds.cu: d <= 12, < 2^32 vertices): beyond
        # them the reference's generator is replayed like for artefact bounds -- the reference still
        # works there (slowly), so must this
        if self.dim > LATTICE_MAX_DIM or lat.counts(self.dim, gen)[0] >= 2 ** 32:
            return False
        try:
            tab = lat.coordinate_tables(self.bounds, gen)
        except lat.LatticeArtefactError:
            return False
        self.mode = "lattice"
        self._tab_host = tab
        self._tab = torch.from_numpy(self._tab_host).to(self.device)
        self._lat_gen = gen
        self.gen = gen
        self._n, self._ngrid = lat.counts(self.dim, gen)
        self._dirty = True
        return True

    def _process_lattice(self):
        d, gen, n = self.dim, self._lat_gen, self._n
        if self._fid is not None:
            self._f, self._feas = hp.lattice_eval(self._fid, self._gid, d, gen, self._tab, 0, n)
        elif self._torch:
            x = hp.lattice_coords(d, gen, self._tab, 0, n)
            self._f, self._feas = self._evaluate_generic(x, None)
        else:
            pts = self._coordinates(np.arange(n, dtype=np.int64))
            self._f, self._feas = self._evaluate_generic(None, pts)
        self._is_min = hp.lattice_star(d, gen, self._f)
        self.V.nfev = int(self._feas.sum().item())
        self._finish_classification(present=None)

    # ---- explicit graphs (mesh, hostgraph) -------------------------------------------------
    def _process_explicit(self):
        n = self._n
        if n == 0:
            return
        pts = self._points_host
        if self._fid is not None:
            tab = self._sampler_tables_if_points_are_sobol(pts)
            if tab is not None:
                self._f, self._feas = hp.eval_sobol(self._fid, self._gid, tab, 0, n)
            else:
                x = torch.from_numpy(pts).to(self.device).t().contiguous()
                self._f, self._feas = hp.eval_points(self._fid, self._gid, x)
        elif self._torch:
            x = torch.from_numpy(pts).to(self.device).t().contiguous()
            self._f, self._feas = self._evaluate_generic(x, None)
        else:
            self._f, self._feas = self._evaluate_generic(None, pts)
        if self.mode == "mesh" and self._morton is not None:
            order, rp_m, col_m = self._morton
            flags_m = hp.star_csr(rp_m, col_m, hp.gather(self._f, order))
            self._is_min = hp.scatter_u8(flags_m, order, torch.empty(n, dtype=torch.uint8, device=self.device))
            nf = hp.count_nfev(self._row_ptr, self._feas)
        else:
            nf = torch.zeros(1, dtype=torch.int64, device=self.device)
            self._is_min = hp.star_csr(self._row_ptr, self._col, self._f, feasible=self._feas, nfev=nf)
        self.V.nfev = int(nf.item())
        present = hp.row_present(self._row_ptr)
        if self.mode == "hostgraph":
            # every vertex the generator created exists, also one without edges yet: it is
            # trivially a minimiser (all([]) is True, _vertex.py:148) and counts towards nfev;
            # in a mesh a point without edges is NOT a vertex
            self._is_min = torch.where(present.bool(), self._is_min, torch.ones_like(self._is_min))
            self.V.nfev = int(self._feas.sum().item())
            present = None
        self._finish_classification(present)
        if self.mode == "hostgraph":
            for v in self._ref.V.cache.values():   # "classified": edge changes from here on count
                v.check_min = False

    def _evaluate_generic(self, x_soa, pts_host):
        """torch-vectorised or host-callable objective + constraints; returns device (f, feasible)."""
        n = self._n
        func, fargs = _split_wrapper(self.sfield)
        fargs = fargs + self.sfield_args
        if x_soa is not None:
            g = None
            if self.g_cons:
                g = torch.stack([torch.as_tensor(gc(x_soa, *ga), dtype=torch.float64, device=self.device)
                                 .expand(n).contiguous() for gc, ga in zip(self.g_cons, self.g_args)])
            f = torch.as_tensor(func(x_soa, *fargs), dtype=torch.float64, device=self.device)
            f = f.expand(n).contiguous().clone()
            return hp.mask_infeasible(f, g)
        # user Python code: called per vertex, once per vertex that exists (results are cached by
        # coordinates like the reference's vertex cache), with the reference's semantics
        # (feasibility_check / compute_sfield, _vertex.py:330-352)
        if self.mode in ("lattice", "hostgraph"):
            present = np.arange(n)
        else:
            rp = self._row_ptr_cpu()
            present = np.flatnonzero(rp[1:] > rp[:-1])
        f_host = np.full(n, np.inf)
        feas_host = np.ones(n, np.uint8)
        known = self._host_values
        for i in present.tolist():
            xa = pts_host[i]
            key = xa.tobytes()
            hit = known.get(key)
            if hit is None:
                ok = True
                if self.g_cons:
                    for gc, ga in zip(self.g_cons, self.g_args):
                        # scalar g: the reference's `g(x) < 0.0`; scipy's copy of the driver also
                        # admits vector-valued constraints (`np.any(g(x) < 0)`)
                        if np.any(np.asarray(gc(xa, *ga)) < 0.0):
                            ok = False
                            break
                v = np.inf
                if ok:
                    try:
                        v = float(np.squeeze(func(xa, *fargs)))
                    except AttributeError:
                        v = np.inf
                    if np.isnan(v):
                        v = np.inf
                hit = known[key] = (v, ok)
            f_host[i] = hit[0]
            feas_host[i] = 1 if hit[1] else 0
        return (torch.from_numpy(f_host).to(self.device), torch.from_numpy(feas_host).to(self.device))

    def _row_ptr_cpu(self):
        if self._row_ptr_host is None:
            self._row_ptr_host = self._row_ptr.cpu().numpy()
        return self._row_ptr_host

    def _col_cpu(self):
        if self._col_host is None:
            self._col_host = self._col.cpu().numpy()
        return self._col_host

    # ---- shared tail: pool, arg-min, materialisation of what the driver will look at ----------
    def _finish_classification(self, present):
        V = self.V
        self._epoch += 1
        idx, cnt = hp.compact_pool(self._is_min)
        pool = idx[: int(cnt.item())].cpu().numpy()
        # the star() self-loop quirk: a vertex that was a local-search start is
        # never a minimiser again once it has been re-classified after an edge change
        if V._started:
            V._excluded |= self._started_and_changed()
            if V._excluded:
                keys = self._stable_keys_of(pool)
                pool = pool[np.array([k not in V._excluded for k in keys], bool)] if len(pool) else pool
        self.pool_indices = pool
        # exact ties of f: the reference's find_lowest_vertex keeps the first strictly lower vertex in
        # cache order (_shgo.py:863-880), i.e. the earliest inserted of the tied ones
        self.argmin_index, self.argmin_value = hp.argmin(
            self._f, present, self._rank if self.mode == "mesh" and self._rank is not None else None)
        V._reset_view()
        want = pool
        if self.argmin_index >= 0:
            want = np.union1d(pool, [self.argmin_index])
        want = np.asarray(want, np.int64)
        if self.mode == "lattice" and len(want) > 1:
            # list the vertices in the reference's cache order as far as it is known in closed form
            # (older generation first; generation 0 exactly): X_min / minimizer_pool inherit it
            first, sec = lat.insertion_key(want, self.dim, self._lat_gen)
            want = want[np.lexsort((sec, first))]
        if self.mode == "mesh" and self._rank is not None and len(want) > 1:
            # list the vertices in the reference's cache order: SHGO.minimizers builds X_min in
            # that order and minimise_pool starts from X_min[0] (_shgo.py:1146-1152,1177)
            r = self._rank[torch.from_numpy(want).to(self.device)].cpu().numpy()
            want = want[np.argsort(r, kind="stable")]
        pool_set = set(pool.tolist())
        self._pool_cache_order = np.array([i for i in want.tolist() if i in pool_set], np.int64)
        self._materialise(want, pool_set=pool_set)

    def ranked_pool(self, xl_maps=(), by="f", xref=None):
        """SURVEY.md section 8(f)-3 on the device: the minimiser pool without the vertices whose
        coordinates equal an earlier local minimum (`xl_maps`, SHGO.minimizers _shgo.py:1116-1122),
        ordered by f (sort_min_pool, _shgo.py:1213-1219) or by distance from `xref` (g_topograph,
        _shgo.py:1227-1243); equal keys keep the reference's cache order.  -> vertex ids."""
        self._process()
        ids = self._pool_cache_order
        if len(ids) == 0:
            return ids
        t = torch.from_numpy(ids).to(self.device)
        x = torch.from_numpy(np.ascontiguousarray(self._coordinates(ids).T)).to(self.device)
        local = hp.pool_rank(torch.arange(len(ids), device=self.device), self._f[t].contiguous(), x,
                             [np.asarray(v, np.float64) for v in xl_maps], by=by, xref=xref)
        return ids[local.cpu().numpy()]

    def _materialise(self, ids, pool_set=None, listed=True):
        if len(ids) == 0:
            return
        ids = np.asarray(ids, np.int64)
        t = torch.from_numpy(ids).to(self.device)
        f = self._f[t].cpu().numpy()
        feas = self._feas[t].cpu().numpy() if self._feas is not None else np.ones(len(ids), np.uint8)
        x = self._coordinates(ids)
        if pool_set is None:
            pool_set = set(self.pool_indices.tolist())
        for i, xi, fi, fe in zip(ids.tolist(), x, f, feas):
            self.V._add(i, xi, fi, i in pool_set, bool(fe), listed)

    def _coordinates(self, ids):
        if self.mode == "lattice":
            p = lat.decode(ids, self.dim, self._lat_gen)
            return np.stack([self._tab_host[i][p[:, i]] for i in range(self.dim)], 1)
        return self._points_host[ids]

    def _materialise_by_coordinate(self, x):
        if self.mode == "lattice":
            p = []
            for i, xi in enumerate(x):
                hit = np.flatnonzero(self._tab_host[i] == xi)
                if len(hit) == 0:
                    return None
                p.append(int(hit[0]))
            par = {q & 1 for q in p}
            if len(par) != 1:
                return None
            if par == {0}:
                vid = int(lat.vertex_ids(np.array([[q // 2 for q in p]]), self.dim, self._lat_gen, True)[0])
            else:
                vid = int(lat.vertex_ids(np.array([[q // 2 for q in p]]), self.dim, self._lat_gen, False)[0])
        elif self._points_host is not None:
            hit = np.flatnonzero(np.all(self._points_host == np.asarray(x), axis=1))
            if len(hit) == 0:
                return None
            vid = int(hit[0])
        else:
            return None
        if self._f is None:
            return None
        self._materialise(np.array([vid], np.int64), listed=False)
        return self.V._by_index.get(vid)

    # ---- local bounds (SHGO.construct_lcb_simplicial, _shgo.py:1246-1270) ----------------------
    def local_bounds(self, v_min, bounds=None):
        """Box of the local minimisation started at `v_min`: per axis the neighbour coordinates
        nearest below / above its own, the global bound where there is none -- in the reference's
        format [[lo, hi], ...].  The boxes of the whole pool are computed by one kernel launch the
        first time one of them is asked for; no neighbour is materialised on the host."""
        if self._dirty or self._f is None:
            raise RuntimeError("local_bounds() before process_pools()")
        index = int(v_min.index)
        if index < 0 or self.V._by_index.get(index) is not v_min:
            raise KeyError("%r is not a vertex of the current complex" % (v_min,))
        bkey = tuple(map(tuple, np.asarray(self.bounds if bounds is None else bounds, float).reshape(-1, 2)))
        if self._lcb_key != (self._epoch, bkey):
            self._lcb, self._lcb_key = {}, (self._epoch, bkey)
        hit = self._lcb.get(index)
        if hit is None:
            ids = np.union1d(self.pool_indices, [index]).astype(np.int64)
            ids = ids[[i not in self._lcb for i in ids.tolist()]]
            pool = torch.from_numpy(ids).to(self.device)
            if self.mode == "lattice":
                lo, hi = hp.lattice_local_bounds(self.dim, self._lat_gen, self._tab, pool, bkey)
            else:
                if self._x_soa is None or self._x_soa.shape[1] != self._n:
                    self._x_soa = torch.from_numpy(self._points_host).to(self.device).t().contiguous()
                lo, hi = hp.local_bounds_csr(self._row_ptr, self._col, self._x_soa, pool, bkey)
            lo, hi = lo.cpu().numpy(), hi.cpu().numpy()
            for k, i in enumerate(ids.tolist()):
                self._lcb[i] = (lo[k], hi[k])
            hit = self._lcb[index]
        return [[float(a), float(b)] for a, b in zip(hit[0], hit[1])]

    def _neighbour_ids(self, index):
        if self.mode == "lattice":
            return lat.neighbours(index, self.dim, self._lat_gen)
        rp, col = self._row_ptr_cpu(), self._col_cpu()
        return col[rp[index]:rp[index + 1]].astype(np.int64)

    def _size(self):
        if self.mode is None:
            return 0
        if self.mode == "lattice":
            return self._n
        if self.mode == "hostgraph":
            return self._n
        rp = self._row_ptr_cpu()
        return int(np.count_nonzero(rp[1:] > rp[:-1]))

    # ---- star() bookkeeping -----------------------------------------------------------------
    def _stable_key(self, v):
        """identity of a vertex across generations / iterations: its coordinates"""
        return v.x

    def _stable_keys_of(self, ids):
        return [tuple(c) for c in self._coordinates(np.asarray(ids, np.int64)).tolist()]

    def _started_and_changed(self):
        if self.mode == "hostgraph":
            changed = set()
            for key in self.V._started:
                rv = self._ref.V.cache.get(key)
                if rv is not None and getattr(rv, "check_min", True):
                    changed.add(key)
            return changed
        # whole-generation refinement and re-triangulation re-classify every vertex
        return set(self.V._started) if self.mode == "lattice" else set()

    # ---- hostgraph ----------------------------------------------------------------------------
    def _ensure_hostgraph(self):
        if self.mode == "hostgraph":
            return
        if self.mode == "mesh":
            raise RuntimeError("refine() on a complex that was built from a mesh")
        RefComplex = _driver_complex_class()
        # field-less in effect: the dummy field is never evaluated (process_pools is never called
        # on it); it only makes the generator create vertices that carry check_min flags
        # the driver's own `domain` object, not a normalised copy: with `symmetry` the reference's
        # triangulate() edits a copy of the bounds in place and behaves differently for an ndarray
        # (what the driver passes) than for a list of pairs
        self._ref = RefComplex(self.dim, domain=self.domain, sfield=_never_called, symmetry=self.symmetry)
        if self.mode == "lattice":
            # replay the whole generations already built, then continue step-wise
            self._ref.refine(None)
            for _ in range(self._lat_gen):
                self._ref.refine_all()
            for v in self._ref.V.cache.values():
                v.check_min = False
            # star() calls made so far (see DeviceVertexCache._mark_started): whole generations are
            # not affected by the self-loops (verified against the reference), what follows is
            for key in self.V._started:
                rv = self._ref.V.cache.get(key)
                if rv is not None:
                    rv.nn.add(rv)
        self.mode = "hostgraph"
        self._ref_keys = []

    def _sync_hostgraph(self):
        """vertices / edges of the Python generator -> points + CSR on the device"""
        cache = self._ref.V.cache
        keys = list(cache.keys())
        assert keys[:len(self._ref_keys)] == self._ref_keys  # insertion order is append-only
        self._ref_keys = keys
        n = len(keys)
        index = {k: i for i, k in enumerate(keys)}
        # self-loops (mirrored star() calls) are not edges of the graph that is classified: their
        # effect on minimiser() is what _started / _excluded track
        deg = np.fromiter((len(v.nn) - (v in v.nn) for v in cache.values()), np.int64, n)
        rp = np.zeros(n + 1, np.int64)
        np.cumsum(deg, out=rp[1:])
        col = np.empty(max(int(rp[-1]), 1), np.int32)
        for i, v in enumerate(cache.values()):
            if deg[i]:
                col[rp[i]:rp[i + 1]] = sorted(index[w.x] for w in v.nn if w is not v)
        col = col[: int(rp[-1])]
        self._points_host = np.array(keys, np.float64).reshape(n, self.dim)
        self._x_soa = None
        self._row_ptr_host, self._col_host = rp, col
        self._row_ptr = torch.from_numpy(rp).to(self.device)
        colp = np.zeros(max(len(col), 4), np.int32)
        colp[:len(col)] = col
        self._col = torch.from_numpy(colp).to(self.device)[:len(col)]
        self._n = n
        self._dirty = True

    # ---- helpers ---------------------------------------------------------------------------
    def _coincident_map(self, pts):
        """canonical index of every point (first occurrence of its coordinates), or None when all
        points are distinct"""
        n = pts.shape[0]
        if n < 2:
            return None
        order = np.lexsort(pts.T[::-1])
        sp = pts[order]
        same = np.all(sp[1:] == sp[:-1], axis=1)
        if not same.any():
            return None
        grp = np.concatenate([[0], np.cumsum(~same)])
        first = np.full(grp[-1] + 1, n, np.int64)
        np.minimum.at(first, grp, order)
        canon = np.empty(n, np.int64)
        canon[order] = first[grp]
        return canon

    def _sampler_tables_if_points_are_sobol(self, pts):
        """The fused generate->evaluate kernel applies when the mesh's points are exactly Sobol
        indices [0, n) of the registered device sampler: every row is compared (cheap next to the
        qhull run that produced the mesh), so a caller that edited an interior row gets the uploaded
        points evaluated, not the regenerated ones."""
        s = self._sampler
        if s is None or s.total != pts.shape[0] or s.host is None:
            return None
        if s.host.shape != pts.shape or not np.array_equal(pts, s.host):
            return None
        return s.tab


def device_sampling_subspace(C, g_cons, g_args, device=None):
    """SHGO.sampling_subspace (_shgo.py:1452-1463) on the device when every constraint can be
    evaluated there (a registered constraint set, or callables marked torch-vectorised); returns the
    filtered C-order array, or None when the constraints are plain host callables."""
    if not g_cons:
        return None
    C = np.ascontiguousarray(C, np.float64)
    n, d = C.shape
    gid = F.constraint_set_id(g_cons, g_args)
    torch_ok = all(_is_torch_vectorized(g) for g in g_cons)
    if not ((gid is not None and gid != F.G_NONE) or torch_ok) or n == 0:
        return None
    dev = hp._dev(device)
    x = torch.from_numpy(C).to(dev).t().contiguous()
    if torch_ok:
        g = torch.stack([torch.as_tensor(gc(x, *ga), dtype=torch.float64, device=dev).expand(n).contiguous()
                         for gc, ga in zip(g_cons, g_args)])
    else:
        # registered sets: the fused kernels give the feasibility flag of the whole set (finite g only,
        # so the two NaN rules coincide); turn it into a one-row "g" of +-1
        _, feas = hp.eval_points(F.PAIRED_OBJECTIVE[gid], gid, x)
        g = (feas.to(torch.float64) * 2.0 - 1.0).reshape(1, n)
    _, kept = hp.sampling_subspace(x, g)
    return kept.cpu().numpy()


def _never_called(x, *args):  # pragma: no cover
    raise RuntimeError("the graph generator must not evaluate the objective")


def _driver_complex_class():
    """The Python Complex of whichever driver package is importable (the reference `shgo`, else
    scipy's vendored copy); used ONLY as a sequential graph generator for partial generations."""
    for mod in ("shgo._shgo_lib._complex", "scipy.optimize._shgo_lib._complex"):
        try:
            m = importlib.import_module(mod)
            return _ORIGINAL_COMPLEX.get(mod, m.Complex)
        except ImportError:
            continue
    raise ImportError("no shgo driver package found (need `shgo` or scipy.optimize)")


_ORIGINAL_COMPLEX = {}
