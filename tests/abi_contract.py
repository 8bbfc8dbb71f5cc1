"""TEST INFRASTRUCTURE: the argument contract of the C ABI (include/shgo_b200.h), checkable on the CPU.

`check(name, *args)` receives exactly what shgo_b200.hotpath marshals for `_lib.call(name, ...)` and
raises AbiContractError where the library's SHGO_REQUIRE lines (csrc/*.cu) would return
SHGO_B200_ERR_ARG.  tests/fake_hotpath.py dry-runs the product's own marshalling code through it
before it computes the answer with the oracle, so that a wrong `ld`, a null pointer or an
out-of-range size is caught by the `-m "not gpu"` suite and not first on a B200.
"""


class AbiContractError(AssertionError):
    pass


def _req(cond, msg, *a):
    if not cond:
        raise AbiContractError(msg % a if a else msg)


K_MAX_DIM = 32
SOBOL_BITS = 30


def _lattice_counts(d, gen):
    K = 1 << gen
    return float(K + 1) ** d, float(K) ** d


def _sobol_tables(sv, d, k0, n, lb, ub):
    _req(sv and lb and ub, "null table pointer")
    _req(1 <= d <= K_MAX_DIM, "dimension %d outside 1..%d for the fused path", d, K_MAX_DIM)
    _req(k0 + n <= (1 << SOBOL_BITS), "Sobol index range exceeds 2^30 points")


def shgo_b200_sobol_fill(sv, d, k0, n, lb, ub, x, ld, layout, stream):
    _sobol_tables(sv, d, k0, n, lb, ub)
    if n == 0:
        return
    _req(x, "null output")
    _req(layout in (0, 1), "bad layout")
    _req(layout == 1 or ld >= n, "ld < n")


def shgo_b200_eval_sobol(functor, gset, d, k0, n, sv, lb, ub, f, feasible, stream):
    _sobol_tables(sv, d, k0, n, lb, ub)
    if n == 0:
        return
    _req(f, "null output")


def shgo_b200_eval_points(functor, gset, d, n, x_soa, ld, f, feasible, stream):
    _req(1 <= d <= K_MAX_DIM, "dimension %d outside 1..%d", d, K_MAX_DIM)
    if n == 0:
        return
    _req(x_soa and f, "null pointer")
    _req(ld >= n, "ld < n")


def shgo_b200_mask_infeasible(f, n, g, m, feasible, stream):
    if n == 0:
        return
    _req(f, "null f")
    _req(m == 0 or g, "m > 0 but g is null")


def shgo_b200_csr_from_simplices(simplices, S, dp1, n, row_ptr, col, col_cap, nnz, tmp, tmp_bytes, stream):
    _req(row_ptr and nnz and tmp, "null pointer")
    _req(simplices or S == 0, "null simplices")
    _req(dp1 >= 2, "simplices need at least two vertex slots (dp1 = %d)", dp1)
    _req(1 <= n <= 0x7fffffff, "vertex count out of int32 range")
    pairs = 3 * S if dp1 >= 3 else S
    _req(col_cap >= 2 * pairs, "col capacity %d < %d", col_cap, 2 * pairs)
    _req(col or pairs == 0, "null col")


def shgo_b200_csr_edges3(row_ptr, col, n, out, cap, count, stream):
    _req(count, "null count")
    if n == 0:
        return
    _req(row_ptr and (out or cap == 0), "null pointer")


def shgo_b200_row_present(row_ptr, n, present, stream):
    if n == 0:
        return
    _req(row_ptr and present, "null pointer")


def shgo_b200_first_seen(simplices, S, dp1, n, first_seen, stream):
    _req(first_seen, "null output")
    _req(simplices or S == 0, "null simplices")
    _req(dp1 >= 2, "simplices need at least two vertex slots (dp1 = %d)", dp1)


def shgo_b200_morton_order(x_soa, ld, d, n, lb, ub, order, newid, tmp, tmp_bytes, stream):
    _req(d >= 1, "d must be positive (got %d)", d)
    _req(n <= 0x7fffffff, "vertex count out of int32 range")
    if n == 0:
        return
    _req(x_soa and lb and ub and order and tmp, "null pointer")
    _req(ld >= n, "ld < n")


def shgo_b200_relabel3(simplices, S, dp1, n, newid, out, stream):
    _req(dp1 >= 2, "simplices need at least two vertex slots (dp1 = %d)", dp1)
    if S == 0:
        return
    _req(simplices and out, "null pointer")


def _gs(src, idx, n, dst, stream):
    if n == 0:
        return
    _req(src and idx and dst, "null pointer")


shgo_b200_gather_f64 = shgo_b200_gather_u8 = shgo_b200_scatter_u8 = _gs


def shgo_b200_star_csr(row_ptr, col, f, row0, rows, nnz_hint, is_min, stream):
    if rows == 0:
        return
    _req(row_ptr and f and is_min, "null pointer")
    _req(col is None or col % 4 == 0, "col must be 4-byte aligned")


def shgo_b200_star_csr_nfev(row_ptr, col, f, feasible, row0, rows, nnz_hint, is_min, nfev, stream):
    _req(nfev, "null nfev")
    shgo_b200_star_csr(row_ptr, col, f, row0, rows, nnz_hint, is_min, stream)


def shgo_b200_star_csr_local(row_ptr, col, f_own, feasible, own0, own_rows, row0, rows, nnz_hint, ends,
                             flags, cand_ws, cand_ws_bytes, nfev, stream):
    _req(own0 + own_rows <= 0x7fffffff, "row range out of int32")
    _req(row0 >= own0 and row0 + rows <= own0 + own_rows, "classified rows outside the owned shard")


def shgo_b200_star_csr_remote(row_ptr, col, f_shards, f_own, nshards, rows_per_shard, my_shard, row0, rows,
                              ends, flags, cand_ws, cand_cap, tmp, tmp_bytes, stream):
    _req(row_ptr and f_shards and f_own and flags and (tmp or cand_ws), "null pointer")
    _req(nshards >= 1 and 0 <= my_shard < nshards, "bad shard index")
    _req(1 <= rows_per_shard <= 0x7fffffff, "bad shard size")
    own0 = my_shard * rows_per_shard
    _req(row0 >= own0 and row0 + rows <= own0 + rows_per_shard, "classified rows outside the owned shard")


def shgo_b200_star_csr_remote_filtered(row_ptr, col, f_shards, f_own, nshards, rows_per_shard, my_shard, row0, rows,
                                       ends, flags, cand_ws, cand_cap, tmp, tmp_bytes, block_min, plog, stream):
    _req(not block_min or 0 <= plog <= 5, "bad group layout")
    shgo_b200_star_csr_remote(row_ptr, col, f_shards, f_own, nshards, rows_per_shard, my_shard, row0, rows, ends,
                              flags, cand_ws, cand_cap, tmp, tmp_bytes, stream)


def shgo_b200_eval_sobol_blockmin(functor, gset, d, k0, n, sv, lb, ub, f, feasible, block_min, stream):
    _sobol_tables(sv, d, k0, n, lb, ub)
    if n == 0:
        return
    _req(f and block_min, "null output")


def shgo_b200_star_csr_fused(row_ptr, col, f_shards, f_own, feasible, nshards, rows_per_shard, my_shard, row0, rows,
                             nnz_hint, ready, epoch, flags, cand_ws, cand_ws_bytes, nfev, stream):
    _req(f_shards and f_own and ready and cand_ws, "null pointer")
    _req(nshards >= 1 and 0 <= my_shard < nshards, "bad shard index")
    _req(rows_per_shard >= 1 and rows_per_shard * nshards <= 0x7fffffff, "bad shard size")
    own0 = my_shard * rows_per_shard
    _req(row0 >= own0 and row0 + rows <= own0 + rows_per_shard, "classified rows outside the owned shard")


def shgo_b200_publish_ready(peer_flags, npeers, me, epoch, stream):
    _req(peer_flags and 1 <= npeers <= 32 and me >= 0, "bad arguments")


def shgo_b200_subspace_mask(g, m, n, keep, stream):
    if n == 0:
        return
    _req(keep, "null output")
    _req(m == 0 or g, "m > 0 but g is null")


def shgo_b200_gather_rows(x_soa, ld, d, idx, m, out, stream):
    _req(d >= 1, "d must be positive (got %d)", d)
    if m == 0:
        return
    _req(x_soa and idx and out, "null pointer")


def shgo_b200_pool_rank(pool, m, f, x_soa, ld, d, xl, k, mode, xref, ranked, count, tmp, tmp_bytes, stream):
    _req(count, "null count")
    _req(d >= 1 and k >= 0 and mode in (0, 1), "bad arguments")
    if m == 0:
        return
    _req(pool and ranked and tmp, "null pointer")
    _req(mode == 1 or f, "null f")
    _req((k == 0 and mode == 0) or x_soa, "coordinates needed")
    _req(k == 0 or xl, "null xl")
    _req(mode == 0 or xref, "null xref")


def shgo_b200_compact_pool(is_min, n, base, idx_out, cap, count, tmp, tmp_bytes, stream):
    _req(count and tmp, "null pointer")
    _req(is_min or n == 0, "null flags")
    _req(idx_out or cap == 0, "null idx_out with cap > 0")


def shgo_b200_argmin_ranked(f, present, rank, n, idx, val, tmp, tmp_bytes, stream):
    _req(idx and val and tmp, "null pointer")
    _req(f or n == 0, "null f")
    _req(tmp_bytes >= 1024 * 24, "argmin workspace: need %d bytes", 1024 * 24)


def shgo_b200_argmin(f, present, n, idx, val, tmp, tmp_bytes, stream):
    shgo_b200_argmin_ranked(f, present, None, n, idx, val, tmp, tmp_bytes, stream)


def shgo_b200_count_nfev(row_ptr, feasible, n, count, stream):
    _req(row_ptr and count, "null pointer")


def _lattice_shape(d, gen, dmax):
    _req(1 <= d <= dmax, "lattice supports 1 <= d <= %d (got %d)", dmax, d)
    _req(0 <= gen <= 20, "generation %d outside 0..20", gen)
    ng, nc = _lattice_counts(d, gen)
    _req(ng + nc < 4294967296.0, "lattice with more than 2^32 vertices is not supported")
    return int(ng) + int(nc)


def shgo_b200_lattice_eval(functor, gset, d, gen, tab, v0, nv, f, feasible, stream):
    nvert = _lattice_shape(d, gen, K_MAX_DIM)
    if nv == 0:
        return
    _req(tab and f, "null pointer")
    _req(v0 + nv <= nvert, "vertex range exceeds the lattice (%d vertices)", nvert)


def shgo_b200_lattice_eval_push(functor, gset, d, gen, tab, v0, nv, f, feasible, f_peers, npeers, part, nparts,
                                granule, stream):
    shgo_b200_lattice_eval(functor, gset, d, gen, tab, v0, nv, f, feasible, stream)
    _req(0 <= npeers <= 15, "at most 15 peer copies (got %d)", npeers)
    _req(nparts >= 1 and 0 <= part < nparts, "bad part %d of %d", part, nparts)
    _req(granule >= 4096 and granule & (granule - 1) == 0, "granule must be a power of two >= 4096")
    _req(nparts == 1 or v0 % granule == 0, "v0 must be a multiple of the granule")


def shgo_b200_lattice_coords(d, gen, tab, v0, nv, x_soa, ld, stream):
    nvert = _lattice_shape(d, gen, K_MAX_DIM)
    if nv == 0:
        return
    _req(tab and x_soa, "null pointer")
    _req(v0 + nv <= nvert and ld >= nv, "bad vertex range / ld")


def shgo_b200_lattice_star_part(d, gen, f, v0, nv, part, nparts, granule, is_min, tmp, tmp_bytes, stream):
    _req(nparts >= 1 and 0 <= part < nparts, "bad part %d of %d", part, nparts)
    _req(granule >= 32 and granule & (granule - 1) == 0, "granule must be a power of two >= 32")
    _req(nparts == 1 or v0 % 32 == 0, "v0 must be a multiple of 32 when the range is dealt out in parts")
    nvert = _lattice_shape(d, gen, 12)
    _req(v0 + nv <= nvert, "vertex range exceeds the lattice")
    if nv == 0:
        return
    _req(f and is_min and tmp, "null pointer")


def shgo_b200_local_bounds_csr(row_ptr, col, x_soa, ld, d, pool, m, lb, ub, lo, hi, stream):
    _req(d >= 1, "d must be positive (got %d)", d)
    if m == 0:
        return
    _req(row_ptr and x_soa and pool and lb and ub and lo and hi, "null pointer")


def shgo_b200_lattice_local_bounds(d, gen, tab, pool, m, lb, ub, lo, hi, stream):
    _lattice_shape(d, gen, 12)
    if m == 0:
        return
    _req(tab and pool and lb and ub and lo and hi, "null pointer")


def check(name, *args):
    fn = globals().get(name)
    if fn is None:
        raise AbiContractError("no contract written for %s" % name)
    fn(*[a if not hasattr(a, "value") else a.value for a in args])
    return 0
