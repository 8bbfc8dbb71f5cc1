// graph.cu -- K3 (vertex-face mesh -> CSR), K4 (CSR star test), pool compaction, arg-min and
// nfev bookkeeping.  Integer / byte work bound by HBM; no tensor cores involved.
#include <cstring>

#include "common.cuh"

namespace shgo_b200 {

// =============================================================================================
// K4: star test over an explicit CSR  (reference _vertex.py:144-151, 422-426)
//
// One WARP owns a tile of 32 consecutive rows; one thread walks one row.  The tile's slice of
// `col` is the contiguous range col[row_ptr[r0] .. row_ptr[r0+32]) and is staged into the warp's
// private shared-memory slice with 16-byte cp.async (LDGSTS) copies, so `col` -- 4*nnz bytes, the
// bulk of the compulsory traffic -- streams from HBM exactly once at full sector efficiency and
// without register staging.  Warps never wait on one another (only __syncwarp), so the 40 resident
// warps of an SM are naturally staggered between "streaming col" and "gathering f".
// The walk itself: neighbour ids come from shared memory, neighbour f values are gathered through
// L1/L2 with ld.global.nc in rounds of 4, 8, then 16, 16, ...; a row stops as soon as one neighbour
// is not strictly larger (the reference's all() short-circuits as well).  On a random field a row
// survives k neighbours with probability 1/(k+1), so ~7 gathers per row are issued instead of deg,
// while the widening rounds keep the dependent-latency chain of the few long-lived rows short
// (measured: 4/8/16 beats 4/8/8 and 8/8/8 on lattice-like AND on random graphs, profiles/).
// Mapping thread <-> row (not warp <-> row) makes the gathers of a warp hit neighbouring sectors
// whenever consecutive rows have "parallel" neighbour lists (lattice-like graphs).
// A tile whose slice exceeds the warp's staging buffer reads `col` from global memory instead.
// The same pass counts nfev (rows that exist and are feasible, reference _vertex.py:184,347).
// =============================================================================================
constexpr int kStarWarps = 8;
constexpr int kStarThreads = kStarWarps * 32;
constexpr int kMaxStarWarps = 16384;  // upper bound on the warps of one star_csr_kernel launch

// L2 policies: `col` and `row_ptr` are read exactly once -> evict_first, so that the 126 MB L2 is
// spent on the f values the gathers come back for instead of on the 30 GB stream passing through.
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_evict_last_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, uint64_t pol)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ int64_t ld_stream_i64(const int64_t *p, uint64_t pol)
{
    int64_t v;
    asm volatile("ld.global.nc.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_keep_f64(const double *p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
// gather of a neighbour's f: same policy, and an L2 miss fetches 64 instead of 128 bytes from DRAM
// (ncu: 8 % less DRAM traffic on the bench graph; cudaLimitMaxL2FetchGranularity is ignored here)
__device__ __forceinline__ double ld_gather_f64(const double *p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.L2::64B.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// One round of the row walk on the staged slice: the next W neighbours (32-bit local indices; the
// tail of a row is handled by clamping, so a short row simply re-tests its last neighbour instead
// of paying for per-entry predicates and +inf selects).
// LOCAL (multi-GPU, phase 1): f is this GPU's shard, indexed by row - row0; neighbours owned by
// another GPU are skipped here and remembered in `remote`.
template <int W, bool LOCAL>
__device__ __forceinline__ bool star_round(const int32_t *__restrict__ sc, const double *__restrict__ f,
                                           double fv, int &k, int last, uint32_t row0, uint32_t rows,
                                           bool &remote, uint64_t keep)
{
    int32_t nb[W];
    double fn[W];
#pragma unroll
    for (int u = 0; u < W; ++u) nb[u] = sc[min(k + u, last)];
#pragma unroll
    for (int u = 0; u < W; ++u) {
        if (LOCAL) {
            const uint32_t l = (uint32_t)nb[u] - row0;
            const bool mine = l < rows;
            remote = remote || !mine;
            fn[u] = mine ? ld_gather_f64(f + l, keep) : pos_inf();
        } else {
            fn[u] = ld_gather_f64(f + nb[u], keep);
        }
    }
    bool ok = true;
#pragma unroll
    for (int u = 0; u < W; ++u) ok = ok && (fv < fn[u]);
    k += W;
    return ok;
}

struct StarMeta {  // what a lane knows about its row of a tile before the slice arrives
    int64_t b, e;
    double fv;
    bool valid;
};
// f_rows: f addressed by GLOBAL row id (the caller passes the virtual base of a shard)
__device__ __forceinline__ StarMeta star_meta(const int64_t *__restrict__ row_ptr,
                                              const double *__restrict__ f_rows, uint64_t row0,
                                              uint64_t rows, uint64_t tile, uint64_t ntiles, int lane,
                                              uint64_t stream_pol, uint64_t keep)
{
    StarMeta m;
    const uint64_t lr = (tile << 5) + lane;
    m.valid = tile < ntiles && lr < rows;
    const uint64_t r = row0 + (m.valid ? lr : rows);  // invalid lanes: empty row at the end
    m.b = ld_stream_i64(row_ptr + r, stream_pol);
    m.e = m.valid ? ld_stream_i64(row_ptr + r + 1, stream_pol) : m.b;
    m.fv = m.valid ? ld_keep_f64(f_rows + r, keep) : 0.0;
    return m;
}

// LOCAL = false: `f` is the whole vector, flags are 0 / 1.
// LOCAL = true : `f` is this GPU's shard (rows row0 .. row0+rows); a row that passed every LOCAL
//                neighbour but also has neighbours on other GPUs gets flag 2 ("candidate") and is
//                finished by star_remote_kernel.
// ENDS (LOCAL only): the caller promises that in every row the neighbours owned by other GPUs sit at
// the two ends (ascending rows; the bench graph).  The row is trimmed to its local middle part once
// and then walked exactly like a single-GPU row -- the per-neighbour ownership test of the general
// LOCAL walk costs ~50 % more instructions per tile, 30 % of the kernel's time (ncu).
template <bool LOCAL, bool ENDS>
__global__ void __launch_bounds__(kStarThreads, 5)
star_csr_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                const double *__restrict__ f, const uint8_t *__restrict__ feasible, uint64_t row0,
                uint64_t rows, uint64_t own0, uint64_t own_rows, int cap_w,
                uint8_t *__restrict__ is_min, uint32_t *__restrict__ cand_ws,
                unsigned long long *__restrict__ nfev)
{
    extern __shared__ __align__(16) int32_t star_smem[];
    const int lane = threadIdx.x & 31;
    int32_t *sc = star_smem + (threadIdx.x >> 5) * cap_w;
    const uint64_t ntiles = (rows + 31) >> 5;
    const uint64_t gw = ((uint64_t)blockIdx.x * kStarThreads + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * kStarThreads) >> 5;
    const double *f_rows = LOCAL ? f - own0 : f;  // LOCAL: f is the owned shard, rows own0 .. own0+own_rows
    unsigned present_feasible = 0;
    // LOCAL: every warp appends its candidates (flag 2) to a private, dense segment of cand_ws, so
    // that phase 2 needs no compaction pass: [0] = warps, [1] = segment capacity, [2..2+kMaxStarWarps)
    // = per-warp counts, then the segments
    const uint32_t seg_cap = (uint32_t)(((ntiles + nw - 1) / nw) << 5);
    uint32_t *seg = (LOCAL && cand_ws) ? cand_ws + 2 + kMaxStarWarps + gw * (uint64_t)seg_cap : nullptr;
    uint32_t ncand = 0;
    const uint64_t stream_pol = l2_evict_first_policy(), keep = l2_evict_last_policy();
    // row_ptr / f of the NEXT tile are fetched while the current one is walked
    StarMeta cur = star_meta(row_ptr, f_rows, row0, rows, gw, ntiles, lane, stream_pol, keep);
    for (uint64_t tile = gw; tile < ntiles; tile += nw) {
        const uint64_t lr = (tile << 5) + lane;
        const int64_t b = cur.b, e = cur.e;
        const double fv = cur.fv;
        const bool valid = cur.valid;
        const int64_t s0 = __shfl_sync(0xffffffffu, b, 0), s1 = __shfl_sync(0xffffffffu, e, 31);
        // align the slice start down to a 16-byte ADDRESS (col may be a shard's virtual base)
        const int64_t s0a = s0 - (int64_t)((reinterpret_cast<uintptr_t>(col + s0) >> 2) & 3);
        const bool staged = s1 - s0a <= cap_w;
        if (staged) {
            const int len = (int)(s1 - s0a);
            const int nvec = len >> 2;  // full 16-byte vectors inside the slice
            const int32_t *src = col + s0a;
            for (int v = lane; v < nvec; v += 32) cp_async16(sc + (v << 2), src + (v << 2), stream_pol);
            for (int k = (nvec << 2) + lane; k < len; k += 32) sc[k] = __ldg(src + k);
        }
        cur = star_meta(row_ptr, f_rows, row0, rows, tile + nw, ntiles, lane, stream_pol, keep);
        bool ok = valid && e > b;
        bool remote = false;
        if (nfev != nullptr && ok && (feasible == nullptr || feasible[lr])) ++present_feasible;
        if (staged) {
            cp_async_wait_all();
            __syncwarp();
            int k = (int)(b - s0a);
            int le = (int)(e - s0a);
            const uint32_t r0 = (uint32_t)own0, nr = (uint32_t)own_rows;
            if (LOCAL && ENDS) {
                const int k0 = k, le0 = le;
                while (k < le && (uint32_t)sc[k] - r0 >= nr) ++k;
                while (le > k && (uint32_t)sc[le - 1] - r0 >= nr) --le;
                remote = k != k0 || le != le0;
                const int last = le - 1;
                // f_rows is addressed by GLOBAL row id, like f of the single-GPU kernel
                if (ok && k < le) ok = star_round<4, false>(sc, f_rows, fv, k, last, r0, nr, remote, keep);
                if (ok && k < le) ok = star_round<8, false>(sc, f_rows, fv, k, last, r0, nr, remote, keep);
                while (ok && k < le) ok = star_round<16, false>(sc, f_rows, fv, k, last, r0, nr, remote, keep);
            } else {
                const int last = le - 1;
                if (ok) ok = star_round<4, LOCAL>(sc, f, fv, k, last, r0, nr, remote, keep);
                if (ok && k < le) ok = star_round<8, LOCAL>(sc, f, fv, k, last, r0, nr, remote, keep);
                while (ok && k < le) ok = star_round<16, LOCAL>(sc, f, fv, k, last, r0, nr, remote, keep);
            }
        } else {
            for (int64_t k = b; k < e && ok; ++k) {
                const int32_t w = __ldg(col + k);
                if (LOCAL) {
                    const uint32_t l = (uint32_t)w - (uint32_t)own0;
                    if (l < (uint32_t)own_rows) ok = fv < ld_nc_f64(f + l); else remote = true;
                } else {
                    ok = fv < ld_nc_f64(f + w);
                }
            }
        }
        const bool candidate = LOCAL && valid && ok && remote;
        if (valid) is_min[lr] = ok ? (candidate ? 2 : 1) : 0;
        if (LOCAL && seg != nullptr) {
            const unsigned m = __ballot_sync(0xffffffffu, candidate);
            if (candidate) seg[ncand + __popc(m & ((1u << lane) - 1u))] = (uint32_t)lr;
            ncand += __popc(m);
        }
        __syncwarp();  // the slice is rewritten by the next tile
    }
    if (LOCAL && seg != nullptr && lane == 0) {
        cand_ws[2 + gw] = ncand;
        if (gw == 0) {
            cand_ws[0] = (uint32_t)nw;
            cand_ws[1] = seg_cap;
        }
    }
    if (nfev != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) present_feasible += __shfl_down_sync(0xffffffffu, present_feasible, o);
        if (lane == 0 && present_feasible) atomicAdd(nfev, (unsigned long long)present_feasible);
    }
}

// ---- multi-GPU, phase 2 ---------------------------------------------------------------------------
// f lives in N equal shards, one per GPU, all mapped into this process over NVLink (CUDA IPC).
// One thread per CANDIDATE row (compacted list: a few percent of the rows) re-reads its neighbour
// list and tests only the neighbours owned by other GPUs, reading their f in place through the
// peer pointers -- plain ld.global over NVLink, all candidates in flight at once, so the link
// latency is hidden by parallelism and only 8 bytes per remote neighbour of a candidate cross it.
struct ShardedF {
    const double *const *shards;  // device array of nshards base pointers (own shard = local HBM)
    uint32_t rows_per_shard;
    __device__ __forceinline__ double operator()(int32_t w) const
    {
        const uint32_t q = (uint32_t)w / rows_per_shard;
        const double *p = shards[q] + ((uint32_t)w - q * rows_per_shard);
        double v;
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
        return v;
    }
};

__device__ __forceinline__ void star_remote_one(uint64_t lr, const int64_t *__restrict__ row_ptr,
                                                const int32_t *__restrict__ col, const ShardedF &facc,
                                                const double *__restrict__ f_own, uint64_t own0,
                                                uint64_t own_rows, uint64_t row0, int remote_at_ends,
                                                uint8_t *__restrict__ is_min)
{
    const uint64_t r = row0 + lr;
    const double fv = ld_nc_f64(f_own + (r - own0));
    bool ok = true;
    const int64_t b = row_ptr[r], e = row_ptr[r + 1];
    const uint32_t o0 = (uint32_t)own0, on = (uint32_t)own_rows;
    if (remote_at_ends) {
        // Promise of the caller: the neighbours owned by other GPUs sit at the two ENDS of the
        // row (e.g. ascending rows: ids below own0 first, ids >= own0+own_rows last).  Read four entries from each end together and
        // fetch the remote ones' values over NVLink: two latency round trips and two sectors of
        // `col` per candidate instead of a walk over the whole row.  A side whose four entries
        // are all remote is continued inwards.
        int64_t klo = b, khi = e;  // [klo, khi) is what has not been looked at yet
        bool more_lo = true, more_hi = true;
        while (ok && klo < khi && (more_lo || more_hi)) {
            int32_t w[8];
            bool use[8];
            double fn[8];
            const int64_t nlo = more_lo ? min((int64_t)4, khi - klo) : 0;
            const int64_t nhi = more_hi ? min((int64_t)4, khi - klo - nlo) : 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                w[u] = u < nlo ? __ldg(col + klo + u) : 0;
                w[4 + u] = u < nhi ? __ldg(col + khi - 1 - u) : 0;
            }
            bool all_lo = nlo == 4, all_hi = nhi == 4;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                use[u] = u < nlo && ((uint32_t)w[u] - o0 >= on);
                use[4 + u] = u < nhi && ((uint32_t)w[4 + u] - o0 >= on);
                all_lo = all_lo && (use[u] || u >= nlo);
                all_hi = all_hi && (use[4 + u] || u >= nhi);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) fn[u] = use[u] ? facc(w[u]) : pos_inf();
#pragma unroll
            for (int u = 0; u < 8; ++u) ok = ok && (fv < fn[u]);
            klo += nlo;
            khi -= nhi;
            more_lo = more_lo && all_lo;
            more_hi = more_hi && all_hi;
        }
    } else {
        // 8 neighbour ids, then their (remote only) f values, are in flight together
        for (int64_t k = b; k < e && ok; k += 8) {
            int32_t w[8];
            double fn[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) w[u] = __ldg(col + (k + u < e ? k + u : e - 1));
#pragma unroll
            for (int u = 0; u < 8; ++u) fn[u] = ((uint32_t)w[u] - o0 < on) ? pos_inf() : facc(w[u]);
#pragma unroll
            for (int u = 0; u < 8; ++u) ok = ok && (fv < fn[u]);
        }
    }
    is_min[lr] = ok ? 1 : 0;
}

// candidates from a compacted index list (phase 1 ran without a candidate workspace)
__global__ void __launch_bounds__(256)
star_remote_kernel(const int64_t *__restrict__ cand, const uint64_t *__restrict__ ncand, uint64_t cap,
                   const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col, ShardedF facc,
                   const double *__restrict__ f_own, uint64_t own0, uint64_t own_rows, uint64_t row0,
                   int remote_at_ends, uint8_t *__restrict__ is_min)
{
    uint64_t n = *ncand;
    if (n > cap) n = cap;
    for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < n;
         c += (uint64_t)gridDim.x * blockDim.x)
        star_remote_one((uint64_t)cand[c], row_ptr, col, facc, f_own, own0, own_rows, row0, remote_at_ends,
                        is_min);
}

// candidates from the per-warp segments phase 1 wrote: no compaction pass in between
__global__ void __launch_bounds__(256)
star_remote_seg_kernel(const uint32_t *__restrict__ cand_ws, const int64_t *__restrict__ row_ptr,
                       const int32_t *__restrict__ col, ShardedF facc, const double *__restrict__ f_own,
                       uint64_t own0, uint64_t own_rows, uint64_t row0, int remote_at_ends,
                       uint8_t *__restrict__ is_min)
{
    const uint32_t nw = cand_ws[0], seg_cap = cand_ws[1];
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, tw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t sw = gw; sw < nw; sw += tw) {
        const uint32_t n = cand_ws[2 + sw];
        const uint32_t *seg = cand_ws + 2 + kMaxStarWarps + sw * (uint64_t)seg_cap;
        for (uint32_t i = lane; i < n; i += 32)
            star_remote_one((uint64_t)seg[i], row_ptr, col, facc, f_own, own0, own_rows, row0, remote_at_ends,
                            is_min);
    }
}

// =============================================================================================
// Multi-block exclusive scan of small counts (uint8 flags or uint32 degrees) into 64-bit offsets.
//   pass 1: per-block sums (4096 items per block)   pass 2: one CTA scans the block sums
//   pass 3: callers re-scan inside each block and add the block offset
// =============================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanBlock = kScanThreads * kScanItems;  // 4096

// uint8 flags count 1 when they equal `match` (match == 0: when non-zero); uint32 items count as is.
// A thread owns kScanItems consecutive items; whole, 16-byte aligned groups are fetched with
// 128-bit loads (byte-wide loads of a 16-byte-strided warp waste 15/16 of every sector request).
template <class T>
__device__ __forceinline__ uint32_t load_items(const T *in, uint64_t base, uint64_t n,
                                               uint32_t (&v)[kScanItems], int match = 0)
{
    static_assert(kScanItems == 16, "vector path assumes 16 items per thread");
    uint32_t s = 0;
    const bool vec = base + kScanItems <= n && (reinterpret_cast<uintptr_t>(in + base) & 15) == 0;
    if (vec) {
        if (sizeof(T) == 1) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(in + base));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < kScanItems; ++i) v[i] = (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(in + base) + g);
                v[4 * g] = q.x, v[4 * g + 1] = q.y, v[4 * g + 2] = q.z, v[4 * g + 3] = q.w;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            uint64_t k = base + i;
            v[i] = k < n ? (uint32_t)in[k] : 0u;
        }
    }
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (sizeof(T) == 1) v[i] = match ? (v[i] == (uint32_t)match) : (v[i] != 0);
        s += v[i];
    }
    return s;
}

// exclusive scan of one value per thread across the CTA; returns the exclusive prefix, *total
// the CTA sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t x, uint32_t *total)
{
    __shared__ uint32_t warp_sums[kScanThreads / 32];
    __shared__ uint32_t block_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t ws = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
        uint32_t winc = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += y;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = winc - ws;
        if (lane == 31) block_total = winc;
    }
    __syncthreads();
    uint32_t res = warp_sums[w] + inc - x;
    *total = block_total;
    __syncthreads();
    return res;
}

template <class T>
__global__ void __launch_bounds__(kScanThreads)
block_sum_kernel(const T *__restrict__ in, uint64_t n, uint64_t *__restrict__ sums, int match)
{
    uint32_t v[kScanItems];
    uint64_t base = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * kScanItems;
    uint32_t s = load_items(in, base, n, v, match);
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// one CTA: in-place exclusive scan of nb 64-bit block sums; grand total -> *total
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint64_t *sums, uint64_t nb, uint64_t *total)
{
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t chunk_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t carry = 0;  // identical in every thread
    for (uint64_t base = 0; base < nb; base += 1024) {
        const uint64_t k = base + threadIdx.x;
        const uint64_t x = k < nb ? sums[k] : 0;
        uint64_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            const uint64_t ws = wsum[lane];
            uint64_t winc = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint64_t y = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += y;
            }
            wsum[lane] = winc - ws;
            if (lane == 31) chunk_total = winc;
        }
        __syncthreads();
        if (k < nb) sums[k] = carry + wsum[w] + inc - x;
        carry += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *total = carry;
        sums[nb] = carry;  // offset of a virtual block past the end (row_ptr[n] when n % 4096 == 0)
    }
}

// pool compaction, pass 3: ascending indices of the non-zero flags
__global__ void __launch_bounds__(kScanThreads)
compact_scatter_kernel(const uint8_t *__restrict__ flags, uint64_t n, const uint64_t *__restrict__ offs,
                       int64_t base_index, int64_t *__restrict__ idx_out, uint64_t cap, int match)
{
    uint32_t v[kScanItems];
    const uint64_t base = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * kScanItems;
    uint32_t s = load_items(flags, base, n, v, match);
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total);
    if (s == 0) return;
    uint64_t o = offs[blockIdx.x] + ex;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (v[i]) {
            if (o < cap) idx_out[o] = base_index + (int64_t)(base + i);
            ++o;
        }
}

// CSR build, pass 3: row_ptr[i] = exclusive prefix of deg; row_ptr[n] = nnz
__global__ void __launch_bounds__(kScanThreads)
rowptr_kernel(const uint32_t *__restrict__ deg, uint64_t n, const uint64_t *__restrict__ offs,
              int64_t *__restrict__ row_ptr)
{
    uint32_t v[kScanItems];
    const uint64_t base = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * kScanItems;
    uint32_t s = load_items(deg, base, n, v);
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total);
    uint64_t o = offs[blockIdx.x] + ex;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i <= n) row_ptr[base + i] = (int64_t)o;  // base+i == n writes nnz
        o += v[i];
    }
}

// =============================================================================================
// K3: vf_to_vv (reference _complex.py:1053-1074).  Undirected edges among simplex slots 0,1,2
// (dp1 >= 3) or the pair itself (dp1 == 2) are de-duplicated in an open-addressing hash set
// (64-bit keys min<<32|max, linear probing, atomicCAS), degrees are counted on first insertion,
// rows are filled by atomic cursors and finally rank-sorted, one warp per row, so that the CSR
// is deterministic.
// =============================================================================================
constexpr uint64_t kEmptyKey = ~0ull;

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

__global__ void edge_insert_kernel(const int32_t *__restrict__ simplices, uint64_t S, int dp1,
                                   uint64_t n, unsigned long long *__restrict__ table,
                                   uint64_t mask, uint32_t *__restrict__ deg, int *__restrict__ bad)
{
    const int npair = dp1 >= 3 ? 3 : 1;
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < S;
         s += (uint64_t)gridDim.x * blockDim.x) {
        const int32_t *sp = simplices + s * (uint64_t)dp1;
        int32_t v[3];
        v[0] = __ldg(sp);
        v[1] = __ldg(sp + 1);
        v[2] = dp1 >= 3 ? __ldg(sp + 2) : 0;
        for (int e = 0; e < npair; ++e) {
            int32_t a = v[e == 2 ? 1 : 0], b = v[e == 0 ? 1 : 2];
            if (a == b) continue;  // connect(): `v is not self` (_vertex.py:123)
            if ((uint32_t)a >= n || (uint32_t)b >= n) {
                *bad = 1;
                continue;
            }
            uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
            unsigned long long key = ((unsigned long long)lo << 32) | hi;
            uint64_t h = mix64(key) & mask;
            while (true) {
                unsigned long long prev = atomicCAS(table + h, kEmptyKey, key);
                if (prev == kEmptyKey) {
                    atomicAdd(deg + lo, 1u);
                    atomicAdd(deg + hi, 1u);
                    break;
                }
                if (prev == key) break;
                h = (h + 1) & mask;
            }
        }
    }
}

// order in which Complex.vf_to_vv first touches each vertex (simplex by simplex, slots 0,1,2):
// the insertion order of the reference's vertex cache, which decides the order of its pool
__global__ void first_seen_kernel(const int32_t *__restrict__ simplices, uint64_t S, int dp1, uint64_t n,
                                  unsigned long long *__restrict__ first)
{
    const int nslot = dp1 >= 3 ? 3 : 2;
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < S;
         s += (uint64_t)gridDim.x * blockDim.x) {
        const int32_t *sp = simplices + s * (uint64_t)dp1;
        for (int k = 0; k < nslot; ++k) {
            const int32_t v = __ldg(sp + k);
            if ((uint32_t)v < n) atomicMin(first + v, (unsigned long long)(s * 3 + k));
        }
    }
}

// a simplex index outside [0, n) poisons the result: report it through nnz = ~0
__global__ void flag_bad_kernel(const int *__restrict__ bad, uint64_t *__restrict__ nnz)
{
    if (*bad) *nnz = ~0ull;
}

__global__ void edge_fill_kernel(const unsigned long long *__restrict__ table, uint64_t slots,
                                 const int64_t *__restrict__ row_ptr, uint32_t *__restrict__ cursor,
                                 int32_t *__restrict__ col_tmp)
{
    for (uint64_t h = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; h < slots;
         h += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long key = table[h];
        if (key == kEmptyKey) continue;
        uint32_t lo = (uint32_t)(key >> 32), hi = (uint32_t)key;
        col_tmp[row_ptr[lo] + atomicAdd(cursor + lo, 1u)] = (int32_t)hi;
        col_tmp[row_ptr[hi] + atomicAdd(cursor + hi, 1u)] = (int32_t)lo;
    }
}

// one warp per row: rank sort (entries of a row are distinct), col_tmp -> col
__global__ void __launch_bounds__(256)
row_sort_kernel(const int64_t *__restrict__ row_ptr, uint64_t n, const int32_t *__restrict__ col_tmp,
                int32_t *__restrict__ col)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
        const int64_t b = row_ptr[r], e = row_ptr[r + 1];
        for (int64_t i = b + lane; i < e; i += 32) {
            const int32_t v = col_tmp[i];
            int64_t rank = 0;
            for (int64_t k = b; k < e; ++k) rank += col_tmp[k] < v;
            col[b + rank] = v;
        }
    }
}

// =============================================================================================
// arg-min (find_lowest_vertex, reference _shgo.py:863-880) and nfev (reference _vertex.py:184,347)
// =============================================================================================
struct MinPair {
    double v;
    int64_t i;
};
__device__ __forceinline__ MinPair min_pair(MinPair a, MinPair b)
{
    if (b.i >= 0 && (a.i < 0 || b.v < a.v || (b.v == a.v && b.i < a.i))) return b;
    return a;
}
__device__ __forceinline__ MinPair block_min(MinPair m)
{
    __shared__ double sv[32];
    __shared__ long long si[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MinPair y{__shfl_down_sync(0xffffffffu, m.v, o), __shfl_down_sync(0xffffffffu, m.i, o)};
        m = min_pair(m, y);
    }
    if (lane == 0) {
        sv[w] = m.v;
        si[w] = m.i;
    }
    __syncthreads();
    if (w == 0) {
        m = lane < nw ? MinPair{sv[lane], si[lane]} : MinPair{pos_inf(), -1};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            MinPair y{__shfl_down_sync(0xffffffffu, m.v, o), __shfl_down_sync(0xffffffffu, m.i, o)};
            m = min_pair(m, y);
        }
    }
    return m;  // valid in thread 0
}

__global__ void __launch_bounds__(256)
argmin_partial_kernel(const double *__restrict__ f, const uint8_t *__restrict__ present, uint64_t n,
                      MinPair *__restrict__ partial)
{
    MinPair m{pos_inf(), -1};
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        if (present && !present[i]) continue;
        double v = ld_nc_f64(f + i);
        if (v < m.v) {  // strict: +inf never wins, first (lowest) index wins ties
            m.v = v;
            m.i = (int64_t)i;
        }
    }
    m = block_min(m);
    if (threadIdx.x == 0) partial[blockIdx.x] = m;
}

__global__ void __launch_bounds__(1024)
argmin_final_kernel(const MinPair *__restrict__ partial, int nparts, int64_t *idx, double *val)
{
    MinPair m{pos_inf(), -1};
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) m = min_pair(m, partial[i]);
    m = block_min(m);
    if (threadIdx.x == 0) {
        *idx = m.i;
        *val = m.i >= 0 ? m.v : pos_inf();
    }
}

__global__ void __launch_bounds__(256)
nfev_kernel(const int64_t *__restrict__ row_ptr, const uint8_t *__restrict__ feasible, uint64_t n,
            unsigned long long *__restrict__ count)
{
    unsigned long long c = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        c += (row_ptr[i + 1] > row_ptr[i]) && (!feasible || feasible[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

static inline uint64_t next_pow2(uint64_t x)
{
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline unsigned grid_for(uint64_t items, int threads, int per_sm)
{
    uint64_t b = (items + threads - 1) / threads;
    uint64_t cap = (uint64_t)sm_count() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace shgo_b200

using namespace shgo_b200;

static size_t star_smem_bytes(uint64_t rows, uint64_t nnz_hint, long *cap_out)
{
    // staging slice per warp: 32 rows * ~1.125 * mean degree (+ alignment slack), 16-byte multiple
    double mean_deg = nnz_hint ? (double)nnz_hint / (double)rows : 32.0;
    long cap = (long)(mean_deg * 32.0 * 1.125) + 16;
    if (cap < 256) cap = 256;
    if (cap > 6144) cap = 6144;  // 24 KB per warp at most (1 CTA of 8 warps per SM)
    cap = (cap + 63) & ~63L;
    *cap_out = cap;
    return (size_t)cap * kStarWarps * sizeof(int32_t);
}

template <bool LOCAL, bool ENDS = false>
static int launch_star(const int64_t *row_ptr, const int32_t *col, const double *f,
                       const uint8_t *feasible, uint64_t row0, uint64_t rows, uint64_t own0,
                       uint64_t own_rows, uint64_t nnz_hint, uint8_t *is_min, uint32_t *cand_ws,
                       uint64_t *nfev, cudaStream_t st)
{
    if (nfev) SHGO_CUDA_TRY(cudaMemsetAsync(nfev, 0, sizeof(uint64_t), st));
    if (rows == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(row_ptr && f && is_min, "null pointer");
    SHGO_REQUIRE((reinterpret_cast<uintptr_t>(col) & 3) == 0, "col must be 4-byte aligned");
    long cap = 0;
    const size_t smem = star_smem_bytes(rows, nnz_hint, &cap);
    auto kern = star_csr_kernel<LOCAL, ENDS>;
    if (smem > 48 * 1024)
        SHGO_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int occ = 0;
    SHGO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kStarThreads, smem));
    if (occ < 1) occ = 1;
    uint64_t tiles = (rows + kStarThreads - 1) / kStarThreads;
    uint64_t grid = (uint64_t)sm_count() * occ;
    if (grid > tiles) grid = tiles;
    if (grid * kStarWarps > (uint64_t)kMaxStarWarps) grid = kMaxStarWarps / kStarWarps;
    kern<<<(unsigned)grid, kStarThreads, smem, st>>>(row_ptr, col, f, feasible, row0, rows, own0, own_rows,
                                                    (int)cap, is_min, cand_ws,
                                                    reinterpret_cast<unsigned long long *>(nfev));
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

int shgo_b200::compact_flags(const uint8_t *flags, uint64_t n, int match, int64_t base, int64_t *idx_out,
                             uint64_t cap, uint64_t *count, void *tmp, cudaStream_t st)
{
    if (n == 0) {
        SHGO_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(uint64_t), st));
        return SHGO_B200_OK;
    }
    uint64_t nb = (n + kScanBlock - 1) / kScanBlock;
    uint64_t *sums = static_cast<uint64_t *>(tmp);
    block_sum_kernel<uint8_t><<<(unsigned)nb, kScanThreads, 0, st>>>(flags, n, sums, match);
    scan_sums_kernel<<<1, 1024, 0, st>>>(sums, nb, count);
    compact_scatter_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(flags, n, sums, base, idx_out, cap, match);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

extern "C" {

int shgo_b200_star_csr(const int64_t *row_ptr, const int32_t *col, const double *f, uint64_t row0,
                       uint64_t rows, uint64_t nnz_hint, uint8_t *is_min, void *stream)
{
    return launch_star<false>(row_ptr, col, f, nullptr, row0, rows, 0, 0, nnz_hint, is_min, nullptr, nullptr,
                              as_stream(stream));
}

int shgo_b200_star_csr_nfev(const int64_t *row_ptr, const int32_t *col, const double *f,
                            const uint8_t *feasible, uint64_t row0, uint64_t rows, uint64_t nnz_hint,
                            uint8_t *is_min, uint64_t *nfev, void *stream)
{
    SHGO_REQUIRE(nfev, "null nfev");
    return launch_star<false>(row_ptr, col, f, feasible, row0, rows, 0, 0, nnz_hint, is_min, nullptr, nfev,
                              as_stream(stream));
}

int shgo_b200_star_csr_local(const int64_t *row_ptr, const int32_t *col, const double *f_own,
                             const uint8_t *feasible, uint64_t own0, uint64_t own_rows, uint64_t row0,
                             uint64_t rows, uint64_t nnz_hint, int remote_at_ends, uint8_t *flags,
                             void *cand_ws, size_t cand_ws_bytes, uint64_t *nfev, void *stream)
{
    SHGO_REQUIRE(own0 + own_rows <= 0x7fffffffull, "row range out of int32");
    SHGO_REQUIRE(row0 >= own0 && row0 + rows <= own0 + own_rows, "classified rows outside the owned shard");
    if (cand_ws && cand_ws_bytes < shgo_b200_star_local_workspace(rows))
        return fail(SHGO_B200_ERR_WORKSPACE, "candidate workspace: need %zu bytes, got %zu",
                    shgo_b200_star_local_workspace(rows), cand_ws_bytes);
    if (remote_at_ends)
        return launch_star<true, true>(row_ptr, col, f_own, feasible, row0, rows, own0, own_rows, nnz_hint,
                                       flags, static_cast<uint32_t *>(cand_ws), nfev, as_stream(stream));
    return launch_star<true>(row_ptr, col, f_own, feasible, row0, rows, own0, own_rows, nnz_hint, flags,
                             static_cast<uint32_t *>(cand_ws), nfev, as_stream(stream));
}

size_t shgo_b200_star_local_workspace(uint64_t rows)
{
    // header + per-warp counts + per-warp segments (each rounded up to whole tiles of 32 rows)
    return align256((2 + (size_t)kMaxStarWarps + rows + 32ull * kMaxStarWarps) * sizeof(uint32_t));
}

size_t shgo_b200_star_remote_workspace(uint64_t rows, uint64_t cand_cap)
{
    return align256(cand_cap * sizeof(int64_t)) + shgo_b200_compact_workspace(rows) + align256(16);
}

int shgo_b200_star_csr_remote(const int64_t *row_ptr, const int32_t *col, const double *const *f_shards,
                              const double *f_own, int nshards, uint64_t rows_per_shard, int my_shard,
                              uint64_t row0, uint64_t rows, int remote_at_ends, uint8_t *flags,
                              const void *cand_ws, uint64_t cand_cap, void *tmp, size_t tmp_bytes,
                              void *stream)
{
    cudaStream_t st = as_stream(stream);
    if (rows == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(row_ptr && f_shards && f_own && flags && (tmp || cand_ws), "null pointer");
    SHGO_REQUIRE(nshards >= 1 && my_shard >= 0 && my_shard < nshards, "bad shard index");
    SHGO_REQUIRE(rows_per_shard >= 1 && rows_per_shard <= 0x7fffffffull, "bad shard size");
    const uint64_t own0 = (uint64_t)my_shard * rows_per_shard;
    SHGO_REQUIRE(row0 >= own0 && row0 + rows <= own0 + rows_per_shard, "classified rows outside the owned shard");
    if (cand_ws) {  // phase 1 already listed the candidates
        ShardedF fa{f_shards, (uint32_t)rows_per_shard};
        star_remote_seg_kernel<<<sm_count() * 8, 256, 0, st>>>(static_cast<const uint32_t *>(cand_ws), row_ptr, col,
                                                             fa, f_own, own0, rows_per_shard, row0,
                                                             remote_at_ends, flags);
        SHGO_CUDA_TRY(cudaGetLastError());
        return SHGO_B200_OK;
    }
    if (tmp_bytes < shgo_b200_star_remote_workspace(rows, cand_cap))
        return fail(SHGO_B200_ERR_WORKSPACE, "remote-phase workspace: need %zu bytes, got %zu",
                    shgo_b200_star_remote_workspace(rows, cand_cap), tmp_bytes);
    unsigned char *p = static_cast<unsigned char *>(tmp);
    int64_t *cand = reinterpret_cast<int64_t *>(p);
    p += align256(cand_cap * sizeof(int64_t));
    void *ctmp = p;
    p += shgo_b200_compact_workspace(rows);
    uint64_t *ncand = reinterpret_cast<uint64_t *>(p);
    if (int rc = compact_flags(flags, rows, /*match=*/2, 0, cand, cand_cap, ncand, ctmp, st)) return rc;
    ShardedF facc{f_shards, (uint32_t)rows_per_shard};
    // candidates beyond cand_cap would be left with flag 2; callers size cand_cap = rows to be safe
    uint64_t blocks = (rows / 16 + 255) / 256;  // ~ one thread per expected candidate, capped
    const uint64_t capb = (uint64_t)sm_count() * 8;
    if (blocks > capb) blocks = capb;
    if (blocks < 1) blocks = 1;
    star_remote_kernel<<<(unsigned)blocks, 256, 0, st>>>(cand, ncand, cand_cap, row_ptr, col, facc, f_own, own0,
                                                        rows_per_shard, row0, remote_at_ends, flags);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

// ---- CUDA IPC plumbing for the peer-mapped f shards ---------------------------------------------
int shgo_b200_ipc_alloc(size_t bytes, void **ptr, unsigned char *handle64)
{
    SHGO_REQUIRE(ptr && handle64 && bytes > 0, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    SHGO_CUDA_TRY(cudaMalloc(ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) {
        cudaFree(*ptr);
        *ptr = nullptr;
        return fail(SHGO_B200_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle64, &h, 64);
    return SHGO_B200_OK;
}

int shgo_b200_ipc_open(const unsigned char *handle64, void **ptr)
{
    SHGO_REQUIRE(ptr && handle64, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SHGO_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SHGO_B200_OK;
}

int shgo_b200_ipc_close(void *ptr)
{
    if (ptr) SHGO_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return SHGO_B200_OK;
}

int shgo_b200_ipc_free(void *ptr)
{
    if (ptr) SHGO_CUDA_TRY(cudaFree(ptr));
    return SHGO_B200_OK;
}

size_t shgo_b200_compact_workspace(uint64_t n)
{
    uint64_t nb = (n + kScanBlock - 1) / kScanBlock;
    return align256((nb + 2) * sizeof(uint64_t));
}

int shgo_b200_compact_pool(const uint8_t *is_min, uint64_t n, int64_t base, int64_t *idx_out,
                           uint64_t cap, uint64_t *count, void *tmp, size_t tmp_bytes, void *stream)
{
    SHGO_REQUIRE(count && tmp, "null pointer");
    SHGO_REQUIRE(is_min || n == 0, "null flags");
    SHGO_REQUIRE(idx_out || cap == 0, "null idx_out with cap > 0");
    if (tmp_bytes < shgo_b200_compact_workspace(n))
        return fail(SHGO_B200_ERR_WORKSPACE, "compact workspace: need %zu bytes, got %zu",
                    shgo_b200_compact_workspace(n), tmp_bytes);
    return compact_flags(is_min, n, /*match=*/1, base, idx_out, cap, count, tmp, as_stream(stream));
}

size_t shgo_b200_csr_workspace(uint64_t S, int dp1, uint64_t n)
{
    uint64_t pairs = dp1 >= 3 ? 3 * S : S;
    uint64_t slots = next_pow2(pairs * 2 < 64 ? 64 : pairs * 2);
    uint64_t nb = (n + kScanBlock - 1) / kScanBlock;
    return align256(slots * 8) + 2 * align256(n * 4) + align256(pairs * 2 * 4 + 16) +
           align256((nb + 2) * 8) + align256(16);
}

int shgo_b200_csr_from_simplices(const int32_t *simplices, uint64_t S, int dp1, uint64_t n,
                                 int64_t *row_ptr, int32_t *col, uint64_t col_cap, uint64_t *nnz,
                                 void *tmp, size_t tmp_bytes, void *stream)
{
    SHGO_REQUIRE(row_ptr && nnz && tmp, "null pointer");
    SHGO_REQUIRE(simplices || S == 0, "null simplices");
    SHGO_REQUIRE(dp1 >= 2, "simplices need at least two vertex slots (dp1 = %d)", dp1);
    SHGO_REQUIRE(n >= 1 && n <= 0x7fffffffull, "vertex count out of int32 range");
    const uint64_t pairs = dp1 >= 3 ? 3 * S : S;
    SHGO_REQUIRE(col_cap >= 2 * pairs, "col capacity %llu < %llu", (unsigned long long)col_cap,
                 (unsigned long long)(2 * pairs));
    SHGO_REQUIRE(col || pairs == 0, "null col");
    if (tmp_bytes < shgo_b200_csr_workspace(S, dp1, n))
        return fail(SHGO_B200_ERR_WORKSPACE, "csr workspace: need %zu bytes, got %zu",
                    shgo_b200_csr_workspace(S, dp1, n), tmp_bytes);
    cudaStream_t st = as_stream(stream);
    const uint64_t slots = next_pow2(pairs * 2 < 64 ? 64 : pairs * 2);
    const uint64_t nb = (n + kScanBlock - 1) / kScanBlock;
    unsigned char *p = static_cast<unsigned char *>(tmp);
    unsigned long long *table = reinterpret_cast<unsigned long long *>(p);
    p += align256(slots * 8);
    uint32_t *deg = reinterpret_cast<uint32_t *>(p);
    p += align256(n * 4);
    uint32_t *cursor = reinterpret_cast<uint32_t *>(p);
    p += align256(n * 4);
    int32_t *col_tmp = reinterpret_cast<int32_t *>(p);
    p += align256(pairs * 2 * 4 + 16);
    uint64_t *sums = reinterpret_cast<uint64_t *>(p);
    p += align256((nb + 2) * 8);
    int *bad = reinterpret_cast<int *>(p);

    SHGO_CUDA_TRY(cudaMemsetAsync(table, 0xff, slots * 8, st));
    SHGO_CUDA_TRY(cudaMemsetAsync(deg, 0, align256(n * 4) * 2, st));  // deg + cursor
    SHGO_CUDA_TRY(cudaMemsetAsync(bad, 0, sizeof(int), st));
    if (S > 0)
        edge_insert_kernel<<<grid_for(S, 256, 8), 256, 0, st>>>(simplices, S, dp1, n, table, slots - 1,
                                                                deg, bad);
    block_sum_kernel<uint32_t><<<(unsigned)nb, kScanThreads, 0, st>>>(deg, n, sums, 0);
    scan_sums_kernel<<<1, 1024, 0, st>>>(sums, nb, nnz);
    // n+1 outputs: one more block when n is a multiple of the block size
    rowptr_kernel<<<(unsigned)(n / kScanBlock + 1), kScanThreads, 0, st>>>(deg, n, sums, row_ptr);
    if (S > 0) {
        edge_fill_kernel<<<grid_for(slots, 256, 8), 256, 0, st>>>(table, slots, row_ptr, cursor, col_tmp);
        row_sort_kernel<<<grid_for(n * 32, 256, 8), 256, 0, st>>>(row_ptr, n, col_tmp, col);
        flag_bad_kernel<<<1, 1, 0, st>>>(bad, nnz);
    }
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

int shgo_b200_first_seen(const int32_t *simplices, uint64_t S, int dp1, uint64_t n, uint64_t *first_seen,
                         void *stream)
{
    SHGO_REQUIRE(first_seen, "null output");
    SHGO_REQUIRE(simplices || S == 0, "null simplices");
    SHGO_REQUIRE(dp1 >= 2, "simplices need at least two vertex slots (dp1 = %d)", dp1);
    cudaStream_t st = as_stream(stream);
    SHGO_CUDA_TRY(cudaMemsetAsync(first_seen, 0xff, n * sizeof(uint64_t), st));
    if (S == 0 || n == 0) return SHGO_B200_OK;
    first_seen_kernel<<<grid_for(S, 256, 8), 256, 0, st>>>(
        simplices, S, dp1, n, reinterpret_cast<unsigned long long *>(first_seen));
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

int shgo_b200_argmin(const double *f, const uint8_t *present, uint64_t n, int64_t *idx, double *val,
                     void *tmp, size_t tmp_bytes, void *stream)
{
    SHGO_REQUIRE(idx && val && tmp, "null pointer");
    SHGO_REQUIRE(f || n == 0, "null f");
    SHGO_REQUIRE(tmp_bytes >= 1024 * sizeof(MinPair), "argmin workspace: need %zu bytes",
                 1024 * sizeof(MinPair));
    cudaStream_t st = as_stream(stream);
    unsigned blocks = grid_for(n ? n : 1, 256, 4);
    if (blocks > 1024) blocks = 1024;
    MinPair *partial = static_cast<MinPair *>(tmp);
    argmin_partial_kernel<<<blocks, 256, 0, st>>>(f, present, n, partial);
    argmin_final_kernel<<<1, 1024, 0, st>>>(partial, (int)blocks, idx, val);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

int shgo_b200_count_nfev(const int64_t *row_ptr, const uint8_t *feasible, uint64_t n, uint64_t *count,
                         void *stream)
{
    SHGO_REQUIRE(row_ptr && count, "null pointer");
    cudaStream_t st = as_stream(stream);
    SHGO_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(uint64_t), st));
    if (n == 0) return SHGO_B200_OK;
    nfev_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(row_ptr, feasible, n,
                                                    reinterpret_cast<unsigned long long *>(count));
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

}  // extern "C"
