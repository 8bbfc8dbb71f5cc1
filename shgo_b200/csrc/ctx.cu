// ctx.cu -- host-buffer convenience path: one Sobol-path shgo iteration with HOST inputs/outputs.
//
// This is what a reference-side caller that holds numpy arrays uses (and what bench.py's `e2e`
// figure times): the CSR derived from qhull's simplices arrives from host memory, is uploaded on
// a copy stream while the compute stream already generates and evaluates the Sobol shard, then the
// star test, pool compaction and nfev count run and the (small) results are copied back.
// Reference path replaced: SHGO.iterate_delaunay's sampling + process_pools + minimizers
// (_shgo.py:1052-1157) minus qhull.
#include <new>

#include "common.cuh"

struct shgo_b200_ctx {
    int device = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t uploaded = nullptr;
    // device buffers (grown on demand, never shrunk)
    int64_t *row_ptr = nullptr;
    int32_t *col = nullptr;
    double *f = nullptr;
    uint8_t *feas = nullptr, *is_min = nullptr;
    int64_t *pool = nullptr;
    void *tmp = nullptr;
    uint32_t *sv = nullptr;
    double *lbub = nullptr;
    uint64_t *scalars = nullptr;  // [0] pool count, [1] nfev
    uint64_t *h_scalars = nullptr;  // pinned mirror
    uint64_t cap_n = 0, cap_nnz = 0, cap_pool = 0;
    size_t cap_tmp = 0;
    // sharded form (one ctx per rank): two peer-mapped f buffers used alternately, see dist.py
    int nshards = 0, my_shard = 0;
    uint64_t shard_rows = 0, step = 0;
    double *f_ipc[2] = {nullptr, nullptr};
    double **peer_table[2] = {nullptr, nullptr};  // device arrays of nshards base pointers
    void *opened[2][32] = {};
    void *cand_ws = nullptr;
    size_t cand_ws_bytes = 0;
    cudaEvent_t evaluated = nullptr;
    int pending_ends = 0;
    uint64_t pending_nnz = 0;
};

namespace {
using namespace shgo_b200;

template <class T>
int grow(T *&p, uint64_t &cap, uint64_t need, size_t extra_bytes = 0)
{
    if (need <= cap && p) return SHGO_B200_OK;
    if (p) SHGO_CUDA_TRY(cudaFree(p));
    p = nullptr;
    SHGO_CUDA_TRY(cudaMalloc(&p, need * sizeof(T) + extra_bytes));
    cap = need;
    return SHGO_B200_OK;
}
}  // namespace

extern "C" {

int shgo_b200_ctx_create(int device, shgo_b200_ctx **out)
{
    SHGO_REQUIRE(out, "null out");
    int count = 0;
    SHGO_CUDA_TRY(cudaGetDeviceCount(&count));
    SHGO_REQUIRE(device >= 0 && device < count, "device %d out of range (%d visible)", device, count);
    SHGO_CUDA_TRY(cudaSetDevice(device));
    shgo_b200_ctx *c = new (std::nothrow) shgo_b200_ctx();
    SHGO_REQUIRE(c, "out of host memory");
    c->device = device;
    SHGO_CUDA_TRY(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
    SHGO_CUDA_TRY(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
    SHGO_CUDA_TRY(cudaEventCreateWithFlags(&c->uploaded, cudaEventDisableTiming));
    SHGO_CUDA_TRY(cudaMalloc(&c->sv, kMaxDim * kSobolBits * sizeof(uint32_t)));
    SHGO_CUDA_TRY(cudaMalloc(&c->lbub, 2 * kMaxDim * sizeof(double)));
    SHGO_CUDA_TRY(cudaMalloc(&c->scalars, 4 * sizeof(uint64_t)));
    SHGO_CUDA_TRY(cudaMallocHost(&c->h_scalars, 4 * sizeof(uint64_t)));
    *out = c;
    return SHGO_B200_OK;
}

void shgo_b200_ctx_destroy(shgo_b200_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->row_ptr);
    cudaFree(c->col);
    cudaFree(c->f);
    cudaFree(c->feas);
    cudaFree(c->is_min);
    cudaFree(c->pool);
    cudaFree(c->tmp);
    cudaFree(c->sv);
    cudaFree(c->lbub);
    cudaFree(c->scalars);
    cudaFreeHost(c->h_scalars);
    for (int b = 0; b < 2; ++b) {
        for (int q = 0; q < 32; ++q)
            if (c->opened[b][q]) cudaIpcCloseMemHandle(c->opened[b][q]);
        cudaFree(c->f_ipc[b]);
        cudaFree(c->peer_table[b]);
    }
    cudaFree(c->cand_ws);
    if (c->evaluated) cudaEventDestroy(c->evaluated);
    if (c->uploaded) cudaEventDestroy(c->uploaded);
    if (c->compute) cudaStreamDestroy(c->compute);
    if (c->copy) cudaStreamDestroy(c->copy);
    delete c;
}

int shgo_b200_iterate_sobol_host(shgo_b200_ctx *c, int functor, int gset, int d, uint64_t k0, uint64_t n,
                                 const uint32_t *sv, const double *lb, const double *ub,
                                 const int64_t *row_ptr, const int32_t *col, int64_t *pool_idx,
                                 uint64_t pool_cap, uint64_t *pool_count, uint64_t *nfev, double *f_out)
{
    SHGO_REQUIRE(c && sv && lb && ub && row_ptr && pool_count, "null pointer");
    SHGO_REQUIRE(d >= 1 && d <= kMaxDim, "dimension %d outside 1..%d", d, kMaxDim);
    SHGO_REQUIRE(n >= 1, "empty problem");
    SHGO_CUDA_TRY(cudaSetDevice(c->device));
    const uint64_t nnz = (uint64_t)row_ptr[n];
    SHGO_REQUIRE(col || nnz == 0, "null col");
    // capacities
    if (n > c->cap_n || !c->f) {
        uint64_t cap = 0;
        if (int rc = grow(c->row_ptr, cap, n + 1)) return rc;
        cap = 0;
        if (int rc = grow(c->f, cap, n)) return rc;
        cap = 0;
        if (int rc = grow(c->feas, cap, n)) return rc;
        cap = 0;
        if (int rc = grow(c->is_min, cap, n)) return rc;
        c->cap_n = n;
        size_t need = shgo_b200_compact_workspace(n);
        if (need > c->cap_tmp) {
            if (c->tmp) SHGO_CUDA_TRY(cudaFree(c->tmp));
            c->tmp = nullptr;
            SHGO_CUDA_TRY(cudaMalloc(&c->tmp, need));
            c->cap_tmp = need;
        }
    }
    if (int rc = grow(c->col, c->cap_nnz, nnz < 4 ? 4 : nnz, 16)) return rc;
    if (pool_cap > 0)
        if (int rc = grow(c->pool, c->cap_pool, pool_cap)) return rc;

    // upload the graph on the copy stream; generate + evaluate concurrently on the compute stream
    SHGO_CUDA_TRY(cudaMemcpyAsync(c->row_ptr, row_ptr, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice,
                                  c->copy));
    if (nnz)
        SHGO_CUDA_TRY(cudaMemcpyAsync(c->col, col, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->copy));
    SHGO_CUDA_TRY(cudaEventRecord(c->uploaded, c->copy));

    SHGO_CUDA_TRY(cudaMemcpyAsync(c->sv, sv, (size_t)d * kSobolBits * sizeof(uint32_t),
                                  cudaMemcpyHostToDevice, c->compute));
    SHGO_CUDA_TRY(cudaMemcpyAsync(c->lbub, lb, d * sizeof(double), cudaMemcpyHostToDevice, c->compute));
    SHGO_CUDA_TRY(cudaMemcpyAsync(c->lbub + kMaxDim, ub, d * sizeof(double), cudaMemcpyHostToDevice,
                                  c->compute));
    if (int rc = shgo_b200_eval_sobol(functor, gset, d, k0, n, c->sv, c->lbub, c->lbub + kMaxDim, c->f,
                                      c->feas, c->compute))
        return rc;
    SHGO_CUDA_TRY(cudaStreamWaitEvent(c->compute, c->uploaded, 0));
    if (int rc = shgo_b200_star_csr_nfev(c->row_ptr, c->col, c->f, c->feas, 0, n, nnz, c->is_min,
                                         c->scalars + 1, c->compute))
        return rc;
    if (int rc = shgo_b200_compact_pool(c->is_min, n, (int64_t)0, c->pool, pool_cap, c->scalars, c->tmp,
                                        c->cap_tmp, c->compute))
        return rc;
    SHGO_CUDA_TRY(cudaMemcpyAsync(c->h_scalars, c->scalars, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                  c->compute));
    if (f_out)
        SHGO_CUDA_TRY(cudaMemcpyAsync(f_out, c->f, n * sizeof(double), cudaMemcpyDeviceToHost, c->compute));
    SHGO_CUDA_TRY(cudaStreamSynchronize(c->compute));
    *pool_count = c->h_scalars[0];
    if (nfev) *nfev = c->h_scalars[1];
    uint64_t m = c->h_scalars[0] < pool_cap ? c->h_scalars[0] : pool_cap;
    if (m && pool_idx) {
        SHGO_CUDA_TRY(cudaMemcpyAsync(pool_idx, c->pool, m * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                      c->compute));
        SHGO_CUDA_TRY(cudaStreamSynchronize(c->compute));
    }
    return SHGO_B200_OK;
}


// ---- sharded form: one ctx per rank -------------------------------------------------------------
int shgo_b200_ctx_shard_init(shgo_b200_ctx *c, int nshards, int my_shard, uint64_t rows_per_shard,
                             unsigned char *handles_out /*[2][64]*/)
{
    SHGO_REQUIRE(c && handles_out, "null pointer");
    SHGO_REQUIRE(nshards >= 1 && nshards <= 32 && my_shard >= 0 && my_shard < nshards, "bad shard index");
    SHGO_REQUIRE(rows_per_shard >= 1 && rows_per_shard * (uint64_t)nshards <= 0x7fffffffull, "bad shard size");
    SHGO_REQUIRE(c->nshards == 0, "the context is already sharded");
    SHGO_CUDA_TRY(cudaSetDevice(c->device));
    for (int b = 0; b < 2; ++b) {
        void *p = nullptr;
        if (int rc = shgo_b200_ipc_alloc(rows_per_shard * sizeof(double), &p, handles_out + 64 * b)) return rc;
        c->f_ipc[b] = static_cast<double *>(p);
        SHGO_CUDA_TRY(cudaMalloc(&c->peer_table[b], 32 * sizeof(double *)));
    }
    SHGO_CUDA_TRY(cudaEventCreateWithFlags(&c->evaluated, cudaEventDisableTiming));
    c->nshards = nshards, c->my_shard = my_shard, c->shard_rows = rows_per_shard, c->step = 0;
    return SHGO_B200_OK;
}

int shgo_b200_ctx_shard_connect(shgo_b200_ctx *c, const unsigned char *all_handles /*[nshards][2][64]*/)
{
    SHGO_REQUIRE(c && all_handles && c->nshards >= 1, "shgo_b200_ctx_shard_init first");
    SHGO_CUDA_TRY(cudaSetDevice(c->device));
    for (int b = 0; b < 2; ++b) {
        double *table[32];
        for (int q = 0; q < c->nshards; ++q) {
            if (q == c->my_shard) {
                table[q] = c->f_ipc[b];
                continue;
            }
            void *p = nullptr;
            if (int rc = shgo_b200_ipc_open(all_handles + ((size_t)q * 2 + b) * 64, &p)) return rc;
            c->opened[b][q] = p;
            table[q] = static_cast<double *>(p);
        }
        SHGO_CUDA_TRY(cudaMemcpy(c->peer_table[b], table, c->nshards * sizeof(double *), cudaMemcpyHostToDevice));
    }
    return SHGO_B200_OK;
}

int shgo_b200_iterate_sobol_host_shard_begin(shgo_b200_ctx *c, int functor, int gset, int d, const uint32_t *sv,
                                             const double *lb, const double *ub, const int64_t *row_ptr,
                                             const int32_t *col, int remote_at_ends)
{
    SHGO_REQUIRE(c && sv && lb && ub && row_ptr, "null pointer");
    SHGO_REQUIRE(c->nshards >= 1 && c->peer_table[0], "shgo_b200_ctx_shard_init / _connect first");
    SHGO_REQUIRE(d >= 1 && d <= kMaxDim, "dimension %d outside 1..%d", d, kMaxDim);
    SHGO_CUDA_TRY(cudaSetDevice(c->device));
    const uint64_t n = c->shard_rows, own0 = (uint64_t)c->my_shard * n;
    const uint64_t off0 = (uint64_t)row_ptr[0], nnz = (uint64_t)row_ptr[n] - off0;
    SHGO_REQUIRE(col || nnz == 0, "null col");
    if (n > c->cap_n || !c->feas) {
        uint64_t cap = 0;
        if (int rc = grow(c->row_ptr, cap, n + 1)) return rc;
        cap = 0;
        if (int rc = grow(c->feas, cap, n)) return rc;
        cap = 0;
        if (int rc = grow(c->is_min, cap, n)) return rc;
        c->cap_n = n;
        size_t need = shgo_b200_compact_workspace(n);
        if (need > c->cap_tmp) {
            if (c->tmp) SHGO_CUDA_TRY(cudaFree(c->tmp));
            c->tmp = nullptr;
            SHGO_CUDA_TRY(cudaMalloc(&c->tmp, need));
            c->cap_tmp = need;
        }
        need = shgo_b200_star_local_workspace(n);
        if (need > c->cand_ws_bytes) {
            if (c->cand_ws) SHGO_CUDA_TRY(cudaFree(c->cand_ws));
            c->cand_ws = nullptr;
            SHGO_CUDA_TRY(cudaMalloc(&c->cand_ws, need));
            c->cand_ws_bytes = need;
        }
    }
    if (int rc = grow(c->col, c->cap_nnz, nnz < 4 ? 4 : nnz, 16)) return rc;
    const int b = (int)(c->step & 1);
    double *f_own = c->f_ipc[b];
    SHGO_CUDA_TRY(cudaMemcpyAsync(c->row_ptr, row_ptr, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->copy));
    if (nnz) SHGO_CUDA_TRY(cudaMemcpyAsync(c->col, col, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->copy));
    SHGO_CUDA_TRY(cudaEventRecord(c->uploaded, c->copy));
    SHGO_CUDA_TRY(cudaMemcpyAsync(c->sv, sv, (size_t)d * kSobolBits * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                  c->compute));
    SHGO_CUDA_TRY(cudaMemcpyAsync(c->lbub, lb, d * sizeof(double), cudaMemcpyHostToDevice, c->compute));
    SHGO_CUDA_TRY(cudaMemcpyAsync(c->lbub + kMaxDim, ub, d * sizeof(double), cudaMemcpyHostToDevice, c->compute));
    if (int rc = shgo_b200_eval_sobol(functor, gset, d, own0, n, c->sv, c->lbub, c->lbub + kMaxDim, f_own, c->feas,
                                      c->compute))
        return rc;
    SHGO_CUDA_TRY(cudaEventRecord(c->evaluated, c->compute));
    SHGO_CUDA_TRY(cudaStreamWaitEvent(c->compute, c->uploaded, 0));
    // phase 1 (local neighbours) needs nobody else; virtual bases: the shard's slices are addressed
    // with GLOBAL row ids / offsets
    if (int rc = shgo_b200_star_csr_local(c->row_ptr - own0, c->col - off0, f_own, c->feas, own0, n, own0, n, nnz,
                                          remote_at_ends ? 1 : 0, c->is_min, c->cand_ws, c->cand_ws_bytes,
                                          c->scalars + 1, c->compute))
        return rc;
    c->pending_ends = remote_at_ends ? 1 : 0;
    c->pending_nnz = off0;
    // the caller's barrier after this call must mean "every shard of f is written"
    SHGO_CUDA_TRY(cudaEventSynchronize(c->evaluated));
    return SHGO_B200_OK;
}

int shgo_b200_iterate_sobol_host_shard_finish(shgo_b200_ctx *c, int64_t *pool_idx, uint64_t pool_cap,
                                              uint64_t *pool_count, uint64_t *nfev)
{
    SHGO_REQUIRE(c && pool_count, "null pointer");
    SHGO_REQUIRE(c->nshards >= 1 && c->cand_ws, "shgo_b200_iterate_sobol_host_shard_begin first");
    SHGO_CUDA_TRY(cudaSetDevice(c->device));
    const uint64_t n = c->shard_rows, own0 = (uint64_t)c->my_shard * n, off0 = c->pending_nnz;
    const int b = (int)(c->step & 1);
    if (pool_cap > 0)
        if (int rc = grow(c->pool, c->cap_pool, pool_cap)) return rc;
    if (c->nshards > 1)
        if (int rc = shgo_b200_star_csr_remote(c->row_ptr - own0, c->col - off0, c->peer_table[b], c->f_ipc[b],
                                               c->nshards, n, c->my_shard, own0, n, c->pending_ends, c->is_min,
                                               c->cand_ws, n, nullptr, 0, c->compute))
            return rc;
    if (int rc = shgo_b200_compact_pool(c->is_min, n, (int64_t)own0, c->pool, pool_cap, c->scalars, c->tmp,
                                        c->cap_tmp, c->compute))
        return rc;
    SHGO_CUDA_TRY(cudaMemcpyAsync(c->h_scalars, c->scalars, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->compute));
    SHGO_CUDA_TRY(cudaStreamSynchronize(c->compute));
    *pool_count = c->h_scalars[0];
    if (nfev) *nfev = c->h_scalars[1];
    const uint64_t m = c->h_scalars[0] < pool_cap ? c->h_scalars[0] : pool_cap;
    if (m && pool_idx) {
        SHGO_CUDA_TRY(cudaMemcpyAsync(pool_idx, c->pool, m * sizeof(int64_t), cudaMemcpyDeviceToHost, c->compute));
        SHGO_CUDA_TRY(cudaStreamSynchronize(c->compute));
    }
    ++c->step;
    return SHGO_B200_OK;
}

}  // extern "C"
