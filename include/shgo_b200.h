/*
 * shgo_b200.h -- C ABI of the B200 (sm_100a) sampling-and-classification hot path of SHGO.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no torch / C++ types.
 * The reference (Stefan-Endres/shgo) is pure Python and has no FFI of its own; each entry point
 * below names the reference code it replaces (paths relative to the reference checkout) and
 * INTEGRATION.md shows the ctypes stub a maintainer would add on the reference side.
 *
 * Conventions
 *   - Unless a parameter is marked [host], every pointer is a DEVICE pointer owned by the
 *     caller.  Nothing is allocated behind the caller's back except inside a shgo_b200_ctx.
 *   - Every function enqueues on the caller's stream (`stream` is a cudaStream_t passed as
 *     void*; NULL = the legacy default stream) and returns immediately; results are valid
 *     after the stream is synchronised.
 *   - Return value: 0 on success, a SHGO_B200_ERR_* code otherwise; the message is available
 *     from shgo_b200_last_error() (thread local).  Functions never throw and never exit.
 *   - There is no CPU fallback: without a CUDA device every compute entry point returns
 *     SHGO_B200_ERR_CUDA.
 *   - Vertex / point indices are int32 in `col` (n <= 2^31) and int64 in `row_ptr` and pools.
 *     Sobol indices are limited to k0 + n <= 2^30 (30-bit generator, like scipy's).
 */
#ifndef SHGO_B200_H
#define SHGO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHGO_B200_VERSION 100

#if defined(__GNUC__)
#define SHGO_B200_API __attribute__((visibility("default")))
#else
#define SHGO_B200_API
#endif

enum {
    SHGO_B200_OK = 0,
    SHGO_B200_ERR_ARG = 1,      /* bad argument (null pointer, unsupported d, misalignment ...) */
    SHGO_B200_ERR_CUDA = 2,     /* CUDA runtime error, incl. "no device" */
    SHGO_B200_ERR_UNSUPPORTED = 3, /* functor / constraint-set combination not compiled in */
    SHGO_B200_ERR_WORKSPACE = 4 /* caller workspace too small; required size reported */
};

/* Registered device functors (objective).  The numpy definition of each one, i.e. the Python
 * callable a reference user passes as `func`, lives in shgo_b200/functors.py. */
enum {
    SHGO_B200_F_ROSENBROCK = 0,      /* scipy.optimize.rosen, any d >= 2 */
    SHGO_B200_F_RASTRIGIN = 1,       /* 10 d + sum(x^2 - 10 cos(2 pi x)) */
    SHGO_B200_F_ACKLEY = 2,
    SHGO_B200_F_STYBLINSKI_TANG = 3, /* 0.5 sum(x^4 - 16 x^2 + 5 x) */
    SHGO_B200_F_EGGHOLDER = 4,       /* d == 2, reference _shgo.py:329-333 */
    SHGO_B200_F_CATTLE_FEED = 5,     /* d == 4, reference _shgo.py:408-409 */
    SHGO_B200_F_SPHERE = 6           /* x0^2 + x1^2 + ... left to right */
};
/* Registered constraint sets g_i(x) >= 0 (reference: `ineq` constraints only, _shgo.py:544-556) */
enum {
    SHGO_B200_G_NONE = 0,
    SHGO_B200_G_CATTLE_FEED = 1, /* g1, g2 of reference _shgo.py:411-418, d == 4 */
    SHGO_B200_G_SUM_LE_6 = 2     /* -(sum(x) - 6) >= 0, reference tests/test__shgo.py:44-45 */
};

SHGO_B200_API int shgo_b200_version(void);
SHGO_B200_API const char *shgo_b200_last_error(void);
/* sm count, compute capability and memory of `device`; fails loudly when there is no GPU */
SHGO_B200_API int shgo_b200_device_info(int device, int *sm_count, int *cc_major, int *cc_minor,
                          size_t *total_mem);

/* ---- a1 + a2: Sobol points --------------------------------------------------------------
 * Replaces qmc.Sobol(d, scramble=False, seed=0).random(n) followed by the bounds scaling
 * loop (reference _shgo.py:703-711, 1441-1449).  Writes points k0 .. k0+n-1 of the
 * unscrambled 30-bit sequence, x = fl(fl(u * (ub - lb)) + lb), bit-identical to the reference.
 *   sv      [d][30] uint32 direction numbers (scipy's `_sv`; see shgo_b200/_sobol.py)
 *   lb, ub  [d] double
 *   x       layout 0: SoA [d][ld] (coordinate-major, ld >= n); layout 1: AoS [n][d] (the
 *           reference's C-order `self.C`)
 */
SHGO_B200_API int shgo_b200_sobol_fill(const uint32_t *sv, int d, uint64_t k0, uint64_t n, const double *lb,
                         const double *ub, double *x, uint64_t ld, int layout, void *stream);

/* ---- a1..a4 fused: generate -> scale -> evaluate f, g -> mask ---------------------------------
 * Replaces, for a registered functor, VertexCacheField.process_pools' evaluation half
 * (reference _vertex.py:324-385: feasibility_check + compute_sfield) on points that are never
 * materialised.  f[i] = +inf when some g_j(x_i) < 0 or when the objective is NaN
 * (_vertex.py:330-336, 351-352).  feasible may be NULL.
 */
SHGO_B200_API int shgo_b200_eval_sobol(int functor, int gset, int d, uint64_t k0, uint64_t n,
                         const uint32_t *sv, const double *lb, const double *ub, double *f,
                         uint8_t *feasible, void *stream);
/* same evaluation on caller-supplied points, SoA [d][ld] */
SHGO_B200_API int shgo_b200_eval_points(int functor, int gset, int d, uint64_t n, const double *x_soa,
                          uint64_t ld, double *f, uint8_t *feasible, void *stream);
/* arbitrary (torch-vectorised) callables: the caller produced f [n] and g [m][n] itself; this
 * applies the reference's masking rules in place: feasible <=> no g_j < 0 (NaN g is feasible),
 * infeasible or NaN f -> +inf.  g may be NULL (m = 0). */
SHGO_B200_API int shgo_b200_mask_infeasible(double *f, uint64_t n, const double *g, int m, uint8_t *feasible,
                              void *stream);

/* ---- a5: vertex-face mesh -> vertex-vertex graph ------------------------------------------------
 * Replaces Complex.vf_to_vv (reference _complex.py:1053-1074).  For dp1 >= 3 each simplex
 * contributes the edges among its first three slots only; for dp1 == 2 the pair itself.  Output is
 * a symmetric CSR with ascending, de-duplicated rows; vertices that never occupy a connecting
 * slot get an empty row (they do not exist in the reference's vertex cache).
 *   simplices [S][dp1] int32;  row_ptr [n+1] int64;  col [col_cap] int32 (6*S always suffices)
 *   nnz       device uint64 receiving the number of directed entries; ~0 (UINT64_MAX) when some
 *             simplex refers to a vertex outside [0, n) (such edges are dropped)
 *   tmp       workspace; call shgo_b200_csr_workspace() for the size
 */
SHGO_B200_API size_t shgo_b200_csr_workspace(uint64_t S, int dp1, uint64_t n);
SHGO_B200_API int shgo_b200_csr_from_simplices(const int32_t *simplices, uint64_t S, int dp1, uint64_t n,
                                 int64_t *row_ptr, int32_t *col, uint64_t col_cap, uint64_t *nnz,
                                 void *tmp, size_t tmp_bytes, void *stream);

/* Every undirected edge {a, b} of a symmetric CSR once, as the degenerate simplex [a, b, b] (a < b, any
 * order): feeding them to the next shgo_b200_csr_from_simplices call together with the new simplices
 * gives the union of the triangulations -- the reference's connect() only ever adds edges
 * (_vertex.py:116-131), so a second vf_to_vv call merges into the first.  out: [cap][3], *count edges
 * found (nnz / 2).  shgo_b200_row_present: present[r] = 1 iff row r is non-empty, i.e. the vertex exists in
 * the reference's cache (_vertex.py:304-317). */
SHGO_B200_API int shgo_b200_csr_edges3(const int64_t *row_ptr, const int32_t *col, uint64_t n, int32_t *out,
                         uint64_t cap, uint64_t *count, void *stream);
SHGO_B200_API int shgo_b200_row_present(const int64_t *row_ptr, uint64_t n, uint8_t *present, void *stream);

/* Position (3*s + slot) at which vf_to_vv first touches each vertex while it walks the simplices
 * (reference _complex.py:1064-1069), ~0 for vertices it never touches.  This is the insertion
 * order of the reference's vertex cache (_vertex.py:304-317), i.e. the order in which
 * SHGO.minimizers lists the pool (_shgo.py:1115-1152). */
SHGO_B200_API int shgo_b200_first_seen(const int32_t *simplices, uint64_t S, int dp1, uint64_t n,
                         uint64_t *first_seen, void *stream);

/* ---- locality-preserving vertex numbering for explicit meshes -----------------
 * qhull numbers a Delaunay mesh like the points it was handed, and Sobol order has no spatial
 * locality: the star test's gathers of neighbour f values are then random 8-byte reads.  These calls
 * build a Z-order (Morton) numbering of the vertices and move data between the two numberings; the
 * classification is the same set of vertices either way (the reference has no counterpart: its
 * vertices are Python objects in a dict, _vertex.py:304-317).
 *   shgo_b200_morton_order  32-bit Z-order key per point (floor(32/d) bits per axis, first 32 axes)
 *                           + stable radix sort: order[k] = old id at new position k, newid[old] = k
 *                           (newid may be NULL).  x_soa: [d][ld] like shgo_b200_eval_points.
 *   shgo_b200_relabel3      out [S][3] = newid[first three slots of every simplex] (dp1 == 2: the
 *                           second slot twice; newid NULL: plain copy) -- exactly what vf_to_vv reads
 *                           (_complex.py:1064-1069); feed it to shgo_b200_csr_from_simplices, dp1 = 3.
 *   gather / scatter        dst[i] = src[idx[i]]  /  dst[idx[i]] = src[i]
 */
SHGO_B200_API size_t shgo_b200_morton_workspace(uint64_t n);
SHGO_B200_API int shgo_b200_morton_order(const double *x_soa, uint64_t ld, int d, uint64_t n, const double *lb,
                           const double *ub, int32_t *order, int32_t *newid, void *tmp,
                           size_t tmp_bytes, void *stream);
SHGO_B200_API int shgo_b200_relabel3(const int32_t *simplices, uint64_t S, int dp1, uint64_t n,
                       const int32_t *newid, int32_t *out, void *stream);
SHGO_B200_API int shgo_b200_gather_f64(const double *src, const int32_t *idx, uint64_t n, double *dst, void *stream);
SHGO_B200_API int shgo_b200_gather_u8(const uint8_t *src, const int32_t *idx, uint64_t n, uint8_t *dst, void *stream);
SHGO_B200_API int shgo_b200_scatter_u8(const uint8_t *src, const int32_t *idx, uint64_t n, uint8_t *dst, void *stream);

/* ---- a6: star test ------------------------------------------------------------------------------------
 * Replaces VertexScalarField.minimiser / proc_minimisers (reference _vertex.py:144-151,
 * 422-426): is_min[r] = 1 iff row row0+r is non-empty and f[row0+r] < f[w] (strict) for every
 * neighbour w.  `row_ptr`/`col` address the whole graph, `f` the whole (gathered) vector; the
 * call classifies rows [row0, row0+rows) -- the shard of one GPU.
 */
SHGO_B200_API int shgo_b200_star_csr(const int64_t *row_ptr, const int32_t *col, const double *f,
                       uint64_t row0, uint64_t rows, uint64_t nnz_hint, uint8_t *is_min,
                       void *stream);
/* same, and in the same pass *nfev = number of classified rows that exist (non-empty) and are
 * feasible (feasible[r], indexed like is_min, may be NULL): the reference's HC.V.nfev
 * (_vertex.py:184,347).  nnz_hint (both calls): number of col entries of the classified rows, used
 * only to size the shared-memory staging buffers; 0 = unknown. */
SHGO_B200_API int shgo_b200_star_csr_nfev(const int64_t *row_ptr, const int32_t *col, const double *f,
                            const uint8_t *feasible, uint64_t row0, uint64_t rows,
                            uint64_t nnz_hint, uint8_t *is_min, uint64_t *nfev, void *stream);

/* ---- multi-GPU star test ------------------------------------------------------
 * The f vector lives in `nshards` equal shards of rows_per_shard values, shard g on GPU g; each GPU
 * classifies its own rows [row0, row0+rows), row0 = my_shard * rows_per_shard, in two phases:
 *
 *  1. shgo_b200_star_csr_local: the single-GPU kernel on the GPU's OWN shard of f (f_own, indexed
 *     by row - row0).  Neighbours owned by other GPUs are skipped.  flags[r] = 0 (not a minimiser),
 *     1 (minimiser, all neighbours local) or 2 (candidate: passed every local neighbour, has remote
 *     ones).  Needs no communication and no barrier; also counts nfev like shgo_b200_star_csr_nfev.
 *  2. shgo_b200_star_csr_remote: compacts the candidates (a few percent of the rows) and tests their
 *     remote neighbours, reading f in place from the peer-mapped shards (f_shards: DEVICE array of
 *     nshards base pointers, own shard = f_own, others mapped over NVLink with shgo_b200_ipc_*).
 *     Only 8 bytes per remote neighbour of a candidate cross NVLink; there is no exchange of f.
 *     Rewrites the 2s in flags to 0 / 1.  All ranks must have finished writing f before any rank
 *     starts phase 2 (caller's barrier).  remote_at_ends != 0 is the caller's promise that in every
 *     row the neighbours owned by other GPUs form a prefix and/or a suffix of the row -- true for
 *     ascending rows (shgo_b200_csr_from_simplices output) because a GPU owns a contiguous id range:
 *     the remote neighbours of a candidate are then looked up at the two ends of its row instead of
 *     by walking it.  cand_cap: room for candidate indices (rows is always
 *     enough); tmp: shgo_b200_star_remote_workspace(rows, cand_cap) bytes.
 * cand_ws (optional, shgo_b200_star_local_workspace(rows) bytes): phase 1 lists its candidates there
 * (dense per-warp segments) and phase 2, given the same pointer, consumes the list directly; with
 * cand_ws == NULL phase 2 compacts the flags itself (tmp / cand_cap).
 * Both calls classify a sub-range [row0, row0+rows) of the owned shard [own0, own0+own_rows)
 * (own0 = my_shard * rows_per_shard), so that a caller can pipeline chunks: phase 2 of one chunk
 * (NVLink bound) overlaps phase 1 of the next (HBM bound) on another stream.  flags / feasible are
 * indexed relative to row0, f_own relative to own0.
 * row_ptr / col are this shard's slices addressed with GLOBAL row ids / offsets (virtual bases).
 * remote_at_ends bit 0 (same promise as for phase 2, see above): every row is trimmed to its local
 * middle part once and then walked like a single-GPU row, without a per-neighbour ownership test.
 * remote_at_ends bit 1 (shgo_b200_star_csr_local only): add to *nfev instead of overwriting it (the
 * chunks of one shard accumulate into one counter).
 * remote_at_ends bits 8..15 (shgo_b200_star_csr_local only): when non-zero, the most warps per SM the
 * kernel may keep resident (default: as many as fit, 40) -- a caller that pipelines chunks leaves
 * room this way for phase 2 of the previous chunk to run underneath on another stream. */
SHGO_B200_API int shgo_b200_star_csr_local(const int64_t *row_ptr, const int32_t *col, const double *f_own,
                             const uint8_t *feasible, uint64_t own0, uint64_t own_rows,
                             uint64_t row0, uint64_t rows, uint64_t nnz_hint, int remote_at_ends,
                             uint8_t *flags, void *cand_ws, size_t cand_ws_bytes, uint64_t *nfev,
                             void *stream);
SHGO_B200_API size_t shgo_b200_star_local_workspace(uint64_t rows);
SHGO_B200_API size_t shgo_b200_star_remote_workspace(uint64_t rows, uint64_t cand_cap);
SHGO_B200_API int shgo_b200_star_csr_remote(const int64_t *row_ptr, const int32_t *col,
                              const double *const *f_shards, const double *f_own, int nshards,
                              uint64_t rows_per_shard, int my_shard, uint64_t row0, uint64_t rows,
                              int remote_at_ends, uint8_t *flags, const void *cand_ws,
                              uint64_t cand_cap, void *tmp, size_t tmp_bytes, void *stream);
/* Phase 2 with a conservative filter (round 2).  block_min holds, for ALL shards and in local memory,
 * the minimum of f over groups of 32 vertices: every rank produces its part with
 * shgo_b200_eval_sobol_blockmin and pushes it into the peers' copies in bulk (copy engines) while
 * phase 1 runs.  A candidate whose f is below the group minimum of a remote neighbour needs no 8-byte
 * read over NVLink for it: about 45 % of the remote reads disappear (a candidate is already the minimum
 * of its local star), for 1/32 of f in bulk traffic.  Results are identical to
 * shgo_b200_star_csr_remote.  A group is what the evaluation kernel can reduce cheaply: with P =
 * 2^layout_plog points per thread (shgo_b200_blockmin_layout(functor, gset)), vertex i belongs to group
 *   ((i >> TL) << (TL - 5)) + ((i & 255) >> (5 - layout_plog)),  TL = 8 + layout_plog. */
SHGO_B200_API int shgo_b200_star_csr_remote_filtered(const int64_t *row_ptr, const int32_t *col,
                              const double *const *f_shards, const double *f_own, int nshards,
                              uint64_t rows_per_shard, int my_shard, uint64_t row0, uint64_t rows,
                              int remote_at_ends, uint8_t *flags, const void *cand_ws,
                              uint64_t cand_cap, void *tmp, size_t tmp_bytes, const double *block_min,
                              int layout_plog, void *stream);
/* shgo_b200_eval_sobol that also writes the group minima of its outputs into block_min (the WHOLE
 * n_total / 32 vector; this call fills entries k0/32 .. (k0+n)/32).  k0 and n must be multiples of the
 * tile 256 << shgo_b200_blockmin_layout(functor, gset). */
SHGO_B200_API int shgo_b200_blockmin_layout(int functor, int gset);
SHGO_B200_API int shgo_b200_eval_sobol_blockmin(int functor, int gset, int d, uint64_t k0, uint64_t n,
                              const uint32_t *sv, const double *lb, const double *ub, double *f,
                              uint8_t *feasible, double *block_min, void *stream);

/* Fused form of the two phases (round 2): phase 1 with remote_at_ends, and in the SAME kernel every
 * candidate with at most 4 remote neighbours reads their f values in place over NVLink as soon as the
 * owning GPUs have published their shard (ready[q] >= epoch, see shgo_b200_publish_ready); the loads
 * are issued at the end of one tile and consumed at the end of the next, so the NVLink round trip
 * hides behind the local work.  Candidates that could not be finished inline (peer not ready yet, more
 * than 4 remote neighbours) keep flag 2 and are listed in cand_ws
 * (required) for shgo_b200_star_csr_remote, which the caller runs after its barrier as before.
 * Replaces the same reference code as shgo_b200_star_csr (_vertex.py:144-151, 422-426).
 *   f_shards DEVICE array of nshards device pointers, nshards <= 16
 *   ready   [nshards] uint32 in LOCAL device memory that the peers write (IPC-mapped on their side)
 *   epoch   value that marks "written for this iteration" (the caller increments it per iteration) */
SHGO_B200_API int shgo_b200_star_csr_fused(const int64_t *row_ptr, const int32_t *col,
                             const double *const *f_shards, const double *f_own,
                             const uint8_t *feasible, int nshards, uint64_t rows_per_shard,
                             int my_shard, uint64_t row0, uint64_t rows, uint64_t nnz_hint,
                             const uint32_t *ready, uint32_t epoch, uint8_t *flags, void *cand_ws,
                             size_t cand_ws_bytes, uint64_t *nfev, void *stream);
/* peer_flags: DEVICE array of npeers pointers, entry q = the `ready` array of GPU q (own or
 * IPC-mapped); stores `epoch` into peer_flags[q][me] for every q with system-scope release
 * semantics, stream-ordered after the kernels that wrote this GPU's shard of f. */
SHGO_B200_API int shgo_b200_publish_ready(uint32_t *const *peer_flags, int npeers, int me, uint32_t epoch,
                            void *stream);
/* cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle / cudaIpcCloseMemHandle / cudaFree;
 * handle64 is the 64-byte cudaIpcMemHandle_t to ship to the peer processes. */
SHGO_B200_API int shgo_b200_ipc_alloc(size_t bytes, void **ptr, unsigned char *handle64);
SHGO_B200_API int shgo_b200_ipc_open(const unsigned char *handle64, void **ptr);
SHGO_B200_API int shgo_b200_ipc_close(void *ptr);
/* one cudaMemcpyAsync(cudaMemcpyDefault): bulk pushes into peer-mapped memory go through the copy
 * engines and take no SM from the kernels running beside them */
SHGO_B200_API int shgo_b200_copy_async(void *dst, const void *src, size_t bytes, void *stream);
SHGO_B200_API int shgo_b200_ipc_free(void *ptr);

/* ---- a7: pool extraction ----------------------------------------------------------------------------
 * Replaces the scan over HC.V.cache in SHGO.minimizers (reference _shgo.py:1109-1157):
 * idx_out receives base + i for every i with is_min[i] != 0, ascending; *count the total.
 * idx_out needs room for cap entries; entries beyond cap are counted but not written.
 */
SHGO_B200_API size_t shgo_b200_compact_workspace(uint64_t n);
SHGO_B200_API int shgo_b200_compact_pool(const uint8_t *is_min, uint64_t n, int64_t base, int64_t *idx_out,
                           uint64_t cap, uint64_t *count, void *tmp, size_t tmp_bytes,
                           void *stream);

/* ---- a4': SHGO.sampling_subspace (reference _shgo.py:1452-1463) --------------------------------------
 * With options['infty_constraints'] = False the driver drops the sample points that violate a
 * constraint BEFORE triangulating: C = C[g(C.T) >= 0.0] for every g.  keep[i] = 1 iff every
 * g[c][i] >= 0.0 (g: [m][n], e.g. the constraint callables evaluated on the device tensor); a NaN
 * drops the point -- the opposite of feasibility_check's rule (shgo_b200_mask_infeasible).  The kept
 * row indices come from shgo_b200_compact_pool(keep), the surviving points in the driver's C-order
 * (m, d) layout from shgo_b200_gather_rows (out[r][j] = x_soa[j][idx[r]]). */
SHGO_B200_API int shgo_b200_subspace_mask(const double *g, int m, uint64_t n, uint8_t *keep, void *stream);
SHGO_B200_API int shgo_b200_gather_rows(const double *x_soa, uint64_t ld, int d, const int64_t *idx, uint64_t m,
                          double *out, void *stream);

/* ---- f3: pool post-processing (reference _shgo.py:1116-1122, 1213-1219, 1227-1243) ------------------
 * SHGO.minimizers skips pool vertices whose coordinates equal an earlier local minimum (LMC.xl_maps),
 * sort_min_pool orders the pool by f, g_topograph by distance from a point.  pool: m vertex ids in the
 * reference's cache order; xl: [k][d] coordinates to exclude (k may be 0).  ranked receives the
 * remaining ids ordered by mode 0: f[id] ascending, mode 1: euclidean distance from xref[d] ascending
 * (sum of squares accumulated axis by axis, then sqrt, like scipy's cdist); equal keys keep the
 * pool's order (the reference's np.argsort does not promise an order among equal keys).  *count =
 * number of ids written.  tmp: shgo_b200_pool_rank_workspace(m) bytes. */
SHGO_B200_API size_t shgo_b200_pool_rank_workspace(uint64_t m);
SHGO_B200_API int shgo_b200_pool_rank(const int64_t *pool, uint64_t m, const double *f, const double *x_soa,
                        uint64_t ld, int d, const double *xl, int k, int mode, const double *xref,
                        int64_t *ranked, uint64_t *count, void *tmp, size_t tmp_bytes, void *stream);

/* ---- a11: find_lowest_vertex (reference _shgo.py:863-880) ---------------------------------------
 * arg-min of f over vertices with present[i] != 0 (present may be NULL = all); ties resolve to the
 * lowest index; *idx = -1 and *val = +inf when nothing is finite.  tmp: 32 * 1024 bytes.
 * _ranked: ties resolve to the lowest rank[i] instead (rank = insertion order of the reference's
 * vertex cache, e.g. shgo_b200_first_seen: the reference keeps the FIRST strictly lower vertex it
 * meets in that order, so among equal values the earliest inserted one is its answer). */
SHGO_B200_API int shgo_b200_argmin(const double *f, const uint8_t *present, uint64_t n, int64_t *idx,
                     double *val, void *tmp, size_t tmp_bytes, void *stream);
SHGO_B200_API int shgo_b200_argmin_ranked(const double *f, const uint8_t *present, const int64_t *rank,
                            uint64_t n, int64_t *idx, double *val, void *tmp, size_t tmp_bytes,
                            void *stream);

/* nfev bookkeeping (reference _vertex.py:184,347): number of vertices that exist (non-empty
 * row) and are feasible (feasible may be NULL). */
SHGO_B200_API int shgo_b200_count_nfev(const int64_t *row_ptr, const uint8_t *feasible, uint64_t n,
                         uint64_t *count, void *stream);

/* ---- a8 + a9: whole-generation simplicial complex ---------------------------------------------
 * Replaces Complex.triangulate / refine_all (reference _complex.py:192-472, 514-976) by the
 * closed form of SURVEY.md Appendix B.  gen = completed refinements, K = 2^gen cells per axis.
 * Vertex ids: grid vertex k in [0,K]^d -> sum k_i (K+1)^i ; centroid c in [0,K)^d ->
 * (K+1)^d + sum c_i K^i.  tab [d][2^(gen+1)+1] holds the per-axis coordinates (built on the host
 * by recursive midpointing, shgo_b200/lattice.py).
 */
SHGO_B200_API int shgo_b200_lattice_eval(int functor, int gset, int d, int gen, const double *tab,
                           uint64_t v0, uint64_t nv, double *f, uint8_t *feasible, void *stream);
/* multi-GPU form: every value is also stored at the same offset of npeers further vectors
 * (f_peers: HOST array of device pointers, e.g. the peers' copies of f opened with
 * shgo_b200_ipc_open, each already advanced to the position of vertex v0) -- evaluation and the
 * exchange of f are one kernel.  npeers <= 15.
 * part / nparts / granule: the id range is cut into granules of `granule` consecutive vertices (a
 * power of two >= SHGO_B200_LATTICE_MIN_GRANULE) and only the granules with (granule index %
 * nparts) == part are evaluated (the others are left untouched) -- GPU `part` of `nparts`;
 * granule-cyclic so that the very uneven star-test cost of a generation spreads over the GPUs.
 * v0 must be a multiple of the granule if nparts > 1. */
#define SHGO_B200_LATTICE_MIN_GRANULE 4096
SHGO_B200_API int shgo_b200_lattice_eval_push(int functor, int gset, int d, int gen, const double *tab,
                                uint64_t v0, uint64_t nv, double *f, uint8_t *feasible,
                                double *const *f_peers, int npeers, int part, int nparts, uint32_t granule,
                                void *stream);
/* coordinates of vertices v0 .. v0+nv-1, SoA [d][ld] */
SHGO_B200_API int shgo_b200_lattice_coords(int d, int gen, const double *tab, uint64_t v0, uint64_t nv,
                             double *x_soa, uint64_t ld, void *stream);
/* implicit-adjacency star test over vertices v0 .. v0+nv-1; f is the whole vector.
 * tmp: shgo_b200_lattice_star_workspace(nv) bytes (list of the few vertices whose star has to be
 * enumerated in full by a whole thread block). */
SHGO_B200_API size_t shgo_b200_lattice_star_workspace(uint64_t nv);
SHGO_B200_API int shgo_b200_lattice_star(int d, int gen, const double *f, uint64_t v0, uint64_t nv,
                           uint8_t *is_min, void *tmp, size_t tmp_bytes, void *stream);
/* the same for the granules owned by `part` of `nparts` (see shgo_b200_lattice_eval_push); is_min
 * still has nv entries, those of other parts' vertices are set to 0 */
SHGO_B200_API int shgo_b200_lattice_star_part(int d, int gen, const double *f, uint64_t v0, uint64_t nv,
                                int part, int nparts, uint32_t granule, uint8_t *is_min, void *tmp,
                                size_t tmp_bytes, void *stream);

/* ---- local bounds of pool vertices -------------------------------------
 * Replaces SHGO.construct_lcb_simplicial, shgo/_shgo.py:1246-1270: per pool vertex and axis, the
 * neighbour coordinate nearest below (lo) / above (hi) the vertex's own, the global bound lb / ub
 * where there is none.  lo, hi: row-major [m][d]; pool: m vertex ids.
 *   _csr     explicit graph, coordinates SoA [d][ld] (as shgo_b200_eval_points takes them)
 *   lattice  whole-generation hypercube complex (closed form on the coordinate table `tab`)
 */
SHGO_B200_API int shgo_b200_local_bounds_csr(const int64_t *row_ptr, const int32_t *col, const double *x_soa,
                               uint64_t ld, int d, const int64_t *pool, uint64_t m, const double *lb,
                               const double *ub, double *lo, double *hi, void *stream);
SHGO_B200_API int shgo_b200_lattice_local_bounds(int d, int gen, const double *tab, const int64_t *pool,
                                   uint64_t m, const double *lb, const double *ub, double *lo, double *hi,
                                   void *stream);

/* ---- host-buffer convenience path (what bench.py's `e2e` times) ---------------------------------
 * One shgo iteration of the Sobol path with HOST inputs and outputs: uploads the CSR (pinned or
 * pageable host memory), generates + evaluates + masks on the device, classifies, compacts and
 * copies the pool indices back.  All device memory lives in the ctx and is reused across calls.
 */
typedef struct shgo_b200_ctx shgo_b200_ctx;
SHGO_B200_API int shgo_b200_ctx_create(int device, shgo_b200_ctx **out);
SHGO_B200_API void shgo_b200_ctx_destroy(shgo_b200_ctx *ctx);
SHGO_B200_API int shgo_b200_iterate_sobol_host(shgo_b200_ctx *ctx, int functor, int gset, int d, uint64_t k0,
                                 uint64_t n, const uint32_t *sv /*[host]*/,
                                 const double *lb /*[host]*/, const double *ub /*[host]*/,
                                 const int64_t *row_ptr /*[host] n+1*/,
                                 const int32_t *col /*[host] nnz*/, int64_t *pool_idx /*[host]*/,
                                 uint64_t pool_cap, uint64_t *pool_count /*[host]*/,
                                 uint64_t *nfev /*[host]*/, double *f_out /*[host] n or NULL*/);

/* Sharded form of the host-buffer path (one ctx per rank / GPU): GPU r of nshards owns Sobol indices
 * and graph rows [r * rows_per_shard, (r+1) * rows_per_shard) and holds its shard of f in two
 * cudaMalloc'ed buffers (used alternately, so one barrier per iteration suffices) that every peer maps
 * over NVLink with CUDA IPC -- the same data path as the device-resident multi-GPU step.
 *   1. _shard_init     allocates the f buffers, returns their two 64-byte IPC handles
 *   2. the caller exchanges the handles of all ranks (MPI_Allgather, torch.distributed, ...)
 *   3. _shard_connect  maps the peers' buffers: all_handles = [nshards][2][64]
 *   per iteration:
 *   4. _shard_begin    uploads the rank's CSR rows (row_ptr: rows_per_shard + 1 GLOBAL offsets, col: the
 *                      matching slice with GLOBAL vertex ids), evaluates its Sobol shard, enqueues phase 1
 *                      of the star test; returns once the rank's shard of f is complete
 *   5. the caller's BARRIER ("every shard of f is written")
 *   6. _shard_finish   phase 2 (remote neighbours of the candidates, read in place over NVLink), pool
 *                      compaction, read-back: pool_idx are global vertex ids of this rank's rows
 * Replaces the same reference code as shgo_b200_iterate_sobol_host. */
SHGO_B200_API int shgo_b200_ctx_shard_init(shgo_b200_ctx *ctx, int nshards, int my_shard,
                             uint64_t rows_per_shard, unsigned char *handles_out /*[host] [2][64]*/);
SHGO_B200_API int shgo_b200_ctx_shard_connect(shgo_b200_ctx *ctx,
                                const unsigned char *all_handles /*[host] [nshards][2][64]*/);
SHGO_B200_API int shgo_b200_iterate_sobol_host_shard_begin(shgo_b200_ctx *ctx, int functor, int gset, int d,
                                             const uint32_t *sv /*[host]*/, const double *lb /*[host]*/,
                                             const double *ub /*[host]*/,
                                             const int64_t *row_ptr /*[host] rows_per_shard+1*/,
                                             const int32_t *col /*[host]*/, int remote_at_ends);
SHGO_B200_API int shgo_b200_iterate_sobol_host_shard_finish(shgo_b200_ctx *ctx, int64_t *pool_idx /*[host]*/,
                                              uint64_t pool_cap, uint64_t *pool_count /*[host]*/,
                                              uint64_t *nfev /*[host]*/);

/* ---- measurement helper ------------------------------------------------------------------------------
 * Dependent-chain-free DFMA loop; *flops = FP64 flops performed.  Used by bench.py to measure the
 * FP64 pipe peak on the box (no published number is trusted). */
SHGO_B200_API int shgo_b200_fp64_peak(int iters, double *sink, uint64_t *flops, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SHGO_B200_H */
