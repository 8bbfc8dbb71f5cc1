"""CPU-side checks of the product's host logic (no GPU, no compute calls)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from shgo_b200 import _lib, build
    build.build_library()
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "shgo_b200.h")).read()
    declared = set(re.findall(r"SHGO_B200_API[^;]*?\b(shgo_b200_\w+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(L, name), name
    assert L.shgo_b200_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from shgo_b200 import _lib
    with pytest.raises(_lib.Shgob200Error):
        _lib.call("shgo_b200_device_info", 0, None, None, None, None)
    from shgo_b200 import hotpath
    with pytest.raises(RuntimeError):
        hotpath.SobolTables([(0, 1)] * 3)


def test_direction_numbers_match_scipy():
    from scipy.stats import qmc
    from shgo_b200 import _sobol
    for d in (1, 2, 5, 8, 10, 41, 333, 1024):
        ref = np.asarray(qmc.Sobol(d=d, scramble=False, seed=0)._sv, np.uint32)
        assert np.array_equal(_sobol.direction_numbers(d), ref), d


def test_functor_registry():
    from scipy.optimize import rosen
    from scipy._lib._util import _FunctionWrapper  # what SHGO wraps func in (_shgo.py:14,515)
    from shgo_b200 import functors as F
    assert F.functor_id(rosen) == F.F_ROSENBROCK
    assert F.functor_id(F.rastrigin) == F.F_RASTRIGIN
    assert F.functor_id(_FunctionWrapper(F.ackley, ())) == F.F_ACKLEY
    assert F.functor_id(_FunctionWrapper(F.ackley, (1,))) is None      # extra args: not the functor
    assert F.functor_id(lambda x: 0.0) is None
    assert F.constraint_set_id(None) == F.G_NONE
    assert F.constraint_set_id((F.cattle_feed_g1, F.cattle_feed_g2)) == F.G_CATTLE_FEED
    assert F.constraint_set_id((F.cattle_feed_g2, F.cattle_feed_g1)) is None
    assert F.constraint_set_id((F.cattle_feed_g1,)) is None
    assert F.constraint_set_id((F.sum_le_6,)) == F.G_SUM_LE_6
    x = np.array([0.3, -1.2, 2.5, 0.7])
    assert F.rosenbrock(x) == rosen(x)


def test_registered_numpy_callables_agree_with_oracle_definitions():
    """The product's numpy callables and the oracle's are the same functions."""
    from oracle import oracle as orc
    from shgo_b200 import functors as F
    rng = np.random.default_rng(0)
    for name, fn in [("rosenbrock", F.rosenbrock), ("rastrigin", F.rastrigin), ("ackley", F.ackley),
                     ("styblinski_tang", F.styblinski_tang), ("sphere", F.sphere)]:
        x = rng.uniform(-3, 3, 6)
        assert fn(x) == orc.NP_OBJECTIVES[name](x)
    x = rng.uniform(0, 1, 4)
    assert F.cattle_feed(x) == orc.np_cattle_feed(x)
    assert F.cattle_feed_g2(x) == orc.np_cattle_g2(x)
    x = rng.uniform(-512, 512, 2)
    assert F.eggholder(x) == orc.np_eggholder(x)


def test_near_tie_report():
    from shgo_b200.ties import explain_pool_difference, ulp_distance
    # path graph 0-1-2-3; evaluation b perturbs f[1] by one ulp across the tie with f[2]
    rp = np.array([0, 1, 3, 5, 6]); col = np.array([1, 0, 2, 1, 3, 2])
    fa = np.array([3.0, 1.0, 1.0, 2.0])
    fb = fa.copy(); fb[1] = np.nextafter(1.0, 0.0)
    ma = np.array([0, 0, 0, 0], bool)            # exact tie: neither 1 nor 2 is a strict minimum
    mb = np.array([0, 1, 0, 0], bool)
    ok, rep = explain_pool_difference(rp, col, fa, fb, ma, mb)
    assert ok and rep[0]["vertex"] == 1 and rep[0]["edges"] == [(1, 2)]
    fb[1] = 0.5                                   # a real difference is not a near-tie
    ok, rep = explain_pool_difference(rp, col, fa, fb, ma, mb)
    assert not ok
    assert ulp_distance(1.0, np.nextafter(1.0, 2.0)) == 1.0


def test_lattice_host_mirror_matches_oracle():
    """shgo_b200.lattice (coordinate tables, id <-> lattice index, neighbour enumeration used to
    materialise stars) against the oracle's closed form."""
    from oracle import oracle as orc
    from shgo_b200 import lattice
    for bounds, gen in [([(0, 1.0)] * 3, 2), ([(-5.12, 5.12)] * 4, 1), ([(-32.768, 32.768)] * 2, 3),
                        ([(0.0, 2.0)] * 5, 1), ([(-1.0, 6.0)], 3)]:
        d = len(bounds)
        assert np.array_equal(lattice.coordinate_tables(bounds, gen), orc.lattice_tables(bounds, gen))
        nv, ng = lattice.counts(d, gen)
        assert (nv, ng) == orc.lattice_nvert(d, gen)
        tab = lattice.coordinate_tables(bounds, gen)
        p = lattice.decode(np.arange(nv), d, gen)
        x = np.stack([tab[i][p[:, i]] for i in range(d)], 1)
        assert np.array_equal(x, orc.lattice_coords(bounds, gen))
        rp, col = orc.lattice_csr(d, gen)
        for v in range(0, nv, max(1, nv // 50)):
            assert sorted(lattice.neighbours(v, d, gen).tolist()) == col[rp[v]:rp[v + 1]].tolist()
        assert np.array_equal(lattice.vertex_ids(p[:ng] // 2, d, gen, True), np.arange(ng))
        assert np.array_equal(lattice.vertex_ids(p[ng:] // 2, d, gen, False), np.arange(ng, nv))
    with pytest.raises(lattice.LatticeArtefactError):
        lattice.coordinate_tables([(-3.3, 7.1)], 2)


def test_ctypes_signatures_mirror_the_header():
    """Every prototype of include/shgo_b200.h against the ctypes argument list the host side
    binds: same arity, and pointer / 64-bit / 32-bit / size_t in the same positions."""
    import ctypes
    from shgo_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "shgo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"SHGO_B200_API\s+([\w\s\*]+?)\b(shgo_b200_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) == len(_lib.SIGNATURES)

    def kind_of_c(param):
        p = " ".join(param.split())
        if p in ("void", ""):
            return None
        if "*" in p:
            return "ptr"
        t = p.rsplit(" ", 1)[0]
        return {"int": "i32", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "size_t": "size",
                "double": "f64"}[t.replace("const ", "")]

    def kind_of_ctypes(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
            return "ptr"
        return {ctypes.c_int: "i32", ctypes.c_uint32: "u32", ctypes.c_uint64: "u64", ctypes.c_int64: "i64",
                ctypes.c_size_t: "size", ctypes.c_double: "f64"}[t]

    for ret, name, params in protos:
        want = [k for k in (kind_of_c(p) for p in params.split(",")) if k]
        restype, argtypes = _lib.SIGNATURES[name]
        got = [kind_of_ctypes(t) for t in argtypes]
        # size_t and uint64_t are the same machine type; the binding may use either
        norm = lambda ks: ["u64" if k == "size" else k for k in ks]
        assert norm(got) == norm(want), (name, got, want)


@pytest.mark.parametrize("name", ["lattice_unit_1d", "lattice_unit_2d", "lattice_unit_3d", "lattice_rastrigin_4d",
                                  "lattice_rosen_5d", "lattice_unit_6d"])
def test_lattice_insertion_key_against_the_reference_cache_order(name):
    """The fixtures list the vertices in the reference's cache order, generation after generation:
    the closed-form key must give every vertex the generation it was created in and reproduce the
    order of generation 0 exactly (that is the part of X_min's order that has a closed form)."""
    from conftest import load_golden
    from shgo_b200 import lattice
    g = load_golden(name)
    bounds = [tuple(b) for b in g["bounds"]]
    d = len(bounds)
    born = {}
    for gen in range(int(g["gens"]) + 1):
        coords = g[f"coords_g{gen}"]
        for c in coords:
            born.setdefault(tuple(c), gen)
        tab = lattice.coordinate_tables(bounds, gen)
        P = np.stack([np.searchsorted(tab[i], coords[:, i]) for i in range(d)], 1)
        assert all(np.array_equal(tab[i][P[:, i]], coords[:, i]) for i in range(d))
        grid = np.all(P % 2 == 0, axis=1)
        ids = np.where(grid, lattice.vertex_ids(P // 2, d, gen, True), lattice.vertex_ids((P - 1) // 2, d, gen, False))
        first, sec = lattice.insertion_key(ids, d, gen)
        assert np.array_equal(first, [born[tuple(c)] for c in coords]), (name, gen)
        g0 = first == 0
        assert np.array_equal(np.argsort(sec[g0], kind="stable"), np.arange(int(g0.sum()))), (name, gen)


def test_functor_support_table_mirrors_the_device_dispatch():
    """functors.supports() must agree with the SHGO_CASE table of csrc/eval.cu and the dim_ok() rules
    of csrc/functors.cuh: anything else has to take the generic route instead of failing on the device."""
    from shgo_b200 import functors as F
    src = open(os.path.join(ROOT, "shgo_b200", "csrc", "eval.cu")).read()
    pairs = set(re.findall(r"SHGO_CASE\(SHGO_B200_F_(\w+), SHGO_B200_G_(\w+),", src))
    names_f = {"ROSENBROCK": F.F_ROSENBROCK, "RASTRIGIN": F.F_RASTRIGIN, "ACKLEY": F.F_ACKLEY,
               "STYBLINSKI_TANG": F.F_STYBLINSKI_TANG, "EGGHOLDER": F.F_EGGHOLDER,
               "CATTLE_FEED": F.F_CATTLE_FEED, "SPHERE": F.F_SPHERE}
    names_g = {"NONE": F.G_NONE, "CATTLE_FEED": F.G_CATTLE_FEED, "SUM_LE_6": F.G_SUM_LE_6}
    table = {(names_f[a], names_g[b]) for a, b in pairs}
    for fid in names_f.values():
        for gid in names_g.values():
            any_dim = any(F.supports(fid, gid, d) for d in range(1, 33))
            assert any_dim == ((fid, gid) in table), (fid, gid)
    assert not F.supports(F.F_ROSENBROCK, F.G_NONE, 1) and F.supports(F.F_ROSENBROCK, F.G_NONE, 2)
    assert F.supports(F.F_EGGHOLDER, F.G_NONE, 2) and not F.supports(F.F_EGGHOLDER, F.G_NONE, 3)
    assert F.supports(F.F_CATTLE_FEED, F.G_CATTLE_FEED, 4) and not F.supports(F.F_CATTLE_FEED, F.G_NONE, 5)
    assert F.supports(F.F_SPHERE, F.G_SUM_LE_6, 7) and not F.supports(F.F_SPHERE, F.G_SUM_LE_6, 8)
    assert not F.supports(F.F_RASTRIGIN, F.G_NONE, 33) and not F.supports(None, F.G_NONE, 3)
    assert not F.supports(F.F_RASTRIGIN, None, 3)


def test_cpu_stand_in_has_the_signatures_of_the_real_hot_path():
    """tests/fake_hotpath.py replaces shgo_b200.hotpath in the CPU tests of the host logic: a call the
    stand-in accepts must be one the real function accepts (same parameter names, same order)."""
    import inspect
    import fake_hotpath
    from shgo_b200 import hotpath
    src = inspect.getsource(fake_hotpath.install)
    names = re.findall(r'"(\w+)"', src[src.index("for name in"):src.index("obj = globals()[name]")])
    assert len(names) >= 18
    for n in names:
        real, fake = getattr(hotpath, n), getattr(fake_hotpath, n)
        if inspect.isclass(real):
            real, fake = real.__init__, fake.__init__
        assert list(inspect.signature(real).parameters) == list(inspect.signature(fake).parameters), n


def test_soa_leading_dimension_of_single_row_tensors():
    """torch reports stride 1 for the size-1 leading dimension of `(n, 1).t().contiguous()`; the C ABI
    wants ld >= n (round-1 GPU failure: every 1-D problem on non-Sobol points raised `ld < n`)."""
    import torch
    from shgo_b200 import hotpath
    x = torch.zeros(94, 1, dtype=torch.float64).t().contiguous()
    assert x.stride(0) == 1 and hotpath.soa_ld(x) == 94
    assert hotpath.soa_ld(torch.zeros(3, 7, dtype=torch.float64)) == 7
    assert hotpath.soa_ld(torch.zeros(3, 16, dtype=torch.float64)[:, :5]) == 16
    assert hotpath.soa_ld(torch.zeros(4, 0, dtype=torch.float64)) == 1
    with pytest.raises(ValueError):
        hotpath.soa_ld(torch.zeros(1, 8, dtype=torch.float64).expand(3, 8))


def test_abi_contract_rejects_what_the_library_rejects():
    """tests/abi_contract.py mirrors the SHGO_REQUIRE lines of csrc/*.cu: spot-check both directions."""
    import abi_contract as A
    with pytest.raises(A.AbiContractError, match="ld < n"):
        A.check("shgo_b200_eval_points", 0, 0, 1, 94, 1000, 1, 2000, 3000, 0)
    A.check("shgo_b200_eval_points", 0, 0, 1, 94, 1000, 94, 2000, 3000, 0)
    with pytest.raises(A.AbiContractError):
        A.check("shgo_b200_lattice_star_part", 13, 0, 1, 0, 10, 0, 1, 4096, 1, 1, 64, 0)
    with pytest.raises(A.AbiContractError, match="no contract"):
        A.check("shgo_b200_not_a_symbol")
    # every compute symbol the product marshals has a contract
    from shgo_b200 import _lib
    skip = ("version", "last_error", "device_info", "workspace", "ipc_", "ctx_", "fp64_peak",
            "iterate_sobol", "lattice_star", "blockmin_layout", "copy_async")
    for name in _lib.SIGNATURES:
        if not any(k in name for k in skip) or name == "shgo_b200_lattice_star_part":
            assert hasattr(A, name), name
