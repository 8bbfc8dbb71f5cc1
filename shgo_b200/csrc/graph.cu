// graph.cu -- K3 (vertex-face mesh -> CSR), pool compaction, arg-min and nfev bookkeeping (the CSR
// star test K4 lives in star.cu).  Integer / byte work bound by HBM; no tensor cores involved.
#include <cstring>

#include "common.cuh"

namespace shgo_b200 {

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

// one CTA: in-place exclusive scan of nb 64-bit block sums; grand total -> *total.  Every thread owns
// kSumItems consecutive sums per sweep (serial scan in registers, then a warp / CTA scan of the thread
// totals): 65 536 sums (n = 2^28) take 8 sweeps instead of 64 (69 -> ~12 us of the bench step).
constexpr int kSumItems = 8;
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint64_t *sums, uint64_t nb, uint64_t *total)
{
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t chunk_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t carry = 0;  // identical in every thread
    for (uint64_t base = 0; base < nb; base += 1024 * kSumItems) {
        const uint64_t k0 = base + (uint64_t)threadIdx.x * kSumItems;
        uint64_t v[kSumItems];
        uint64_t x = 0;
#pragma unroll
        for (int i = 0; i < kSumItems; ++i) {
            v[i] = k0 + i < nb ? sums[k0 + i] : 0;
            x += v[i];
        }
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
        uint64_t run = carry + wsum[w] + inc - x;
#pragma unroll
        for (int i = 0; i < kSumItems; ++i) {
            if (k0 + i < nb) sums[k0 + i] = run;
            run += v[i];
        }
        carry += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *total = carry;
        sums[nb] = carry;  // offset of a virtual block past the end (row_ptr[n] when n % 4096 == 0)
    }
}

// pool compaction, pass 3: ascending indices of the non-zero flags.  The block's indices are first
// laid out densely in shared memory and then written as one contiguous, coalesced run (every thread
// storing its own few indices straight to global memory ran at 1.7 TB/s on the bench pool).
__global__ void __launch_bounds__(kScanThreads)
compact_scatter_kernel(const uint8_t *__restrict__ flags, uint64_t n, const uint64_t *__restrict__ offs,
                       int64_t base_index, int64_t *__restrict__ idx_out, uint64_t cap, int match)
{
    __shared__ uint32_t stage[kScanBlock];  // block-local positions of the set flags
    uint32_t v[kScanItems];
    const uint64_t block0 = (uint64_t)blockIdx.x * kScanBlock;
    const uint32_t local0 = threadIdx.x * kScanItems;
    uint32_t s = load_items(flags, block0 + local0, n, v, match);
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total);
    if (total == 0) return;  // uniform
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (v[i]) stage[ex++] = local0 + i;
    __syncthreads();
    const uint64_t o = offs[blockIdx.x];
    for (uint32_t k = threadIdx.x; k < total; k += kScanThreads)
        if (o + k < cap) idx_out[o + k] = base_index + (int64_t)(block0 + stage[k]);
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
// (dp1 >= 3) or the pair itself (dp1 == 2).  Round 2: a BUCKET build instead of a global hash set --
//   count    every pair adds 1 to both endpoints' counters (duplicates included: an edge of a 2-D
//            Delaunay mesh is seen by two triangles)
//   scan     -> offsets of the duplicate-laden rows
//   fill     every pair drops each endpoint into the other's row (atomic cursor per row)
//   sort     every row is sorted and de-duplicated in place (a thread per row up to 32 entries, a CTA
//            radix sort for hubs) -> degree
//   scan + compact -> row_ptr, col (ascending, unique: deterministic whatever the atomics' order)
// The atomics now hit n 4-byte counters (L2-resident for any mesh qhull can produce) instead of a
// hash table of 2 x pairs 8-byte slots scattered over gigabytes: 2-D Delaunay of 2^24 points 21 -> ~3 ms.
// =============================================================================================
template <class Fn>
__device__ __forceinline__ void for_each_pair(const int32_t *__restrict__ simplices, uint64_t S, int dp1, uint64_t n,
                                              int *__restrict__ bad, Fn fn)
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
            const int32_t a = v[e == 2 ? 1 : 0], b = v[e == 0 ? 1 : 2];
            if (a == b) continue;  // connect(): `v is not self` (_vertex.py:123)
            if ((uint32_t)a >= n || (uint32_t)b >= n) {
                *bad = 1;
                continue;
            }
            fn((uint32_t)a, (uint32_t)b);
        }
    }
}

__global__ void __launch_bounds__(256)
edge_count_kernel(const int32_t *__restrict__ simplices, uint64_t S, int dp1, uint64_t n,
                  uint32_t *__restrict__ deg, int *__restrict__ bad)
{
    for_each_pair(simplices, S, dp1, n, bad, [&](uint32_t a, uint32_t b) {
        atomicAdd(deg + a, 1u);
        atomicAdd(deg + b, 1u);
    });
}

__global__ void __launch_bounds__(256)
edge_fill_kernel(const int32_t *__restrict__ simplices, uint64_t S, int dp1, uint64_t n,
                 const int64_t *__restrict__ rp_dup, uint32_t *__restrict__ cursor, int32_t *__restrict__ rows,
                 int *__restrict__ bad)
{
    for_each_pair(simplices, S, dp1, n, bad, [&](uint32_t a, uint32_t b) {
        rows[rp_dup[a] + atomicAdd(cursor + a, 1u)] = (int32_t)b;
        rows[rp_dup[b] + atomicAdd(cursor + b, 1u)] = (int32_t)a;
    });
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

// Rows (with duplicates) are sorted ascending and de-duplicated IN PLACE; deg[r] = distinct entries.
//   <= 32 entries  one thread per row: insertion sort in a private array, unique entries written back
//   33 .. 1024     listed, then one warp per row: bitonic sort in shared memory (row_sort_warp_kernel)
//   longer (hubs)  listed again, then one CTA per row: stable LSD radix sort (8 bits per pass, only the
//                  passes the vertex count needs) ping-ponging between the row and a scratch row in
//                  `col`, then a CTA-wide unique compaction back into the row -- O(len) per pass where
//                  round 1 ranked with O(len^2) comparisons on one warp
__global__ void __launch_bounds__(256)
row_sort_small_kernel(const int64_t *__restrict__ rp_dup, uint64_t n, int32_t *__restrict__ rows,
                      uint32_t *__restrict__ deg, uint32_t *__restrict__ big_list, int *__restrict__ big_count)
{
    // one THREAD per row: a mesh row has ~12 entries (duplicates included); a warp per row spent 32
    // shuffle steps on it whatever its length (7.2 of the 12 ms of a 2^24-vertex build, ncu)
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n;
         r += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t b = rp_dup[r];
        const int len = (int)min((int64_t)33, rp_dup[r + 1] - b);
        if (len == 0) {
            deg[r] = 0;
            continue;
        }
        if (len > 32) {
            big_list[atomicAdd(big_count, 1)] = (uint32_t)r;
            continue;
        }
        int32_t v[32];
        for (int i = 0; i < len; ++i) v[i] = rows[b + i];
        for (int i = 1; i < len; ++i) {  // insertion sort
            const int32_t x = v[i];
            int j = i - 1;
            while (j >= 0 && v[j] > x) {
                v[j + 1] = v[j];
                --j;
            }
            v[j + 1] = x;
        }
        int m = 0;
        for (int i = 0; i < len; ++i)
            if (i == 0 || v[i] != v[i - 1]) rows[b + m++] = v[i];
        deg[r] = (uint32_t)m;
    }
}

// rows of 33 .. kWarpRowMax entries (the dup-laden rows of 3-D and higher meshes: a vertex sits in
// dozens of simplices): one WARP per listed row, bitonic sort in a shared-memory line, unique entries
// compacted with ballots; longer rows are passed on to the CTA radix sort
constexpr int kWarpRowMax = 1024;
__global__ void __launch_bounds__(256)
row_sort_warp_kernel(const int64_t *__restrict__ rp_dup, int32_t *__restrict__ rows, uint32_t *__restrict__ deg,
                     const uint32_t *__restrict__ big_list, const int *__restrict__ big_count,
                     uint32_t *__restrict__ huge_list, int *__restrict__ huge_count)
{
    __shared__ int32_t line[8][kWarpRowMax];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    int32_t *sm = line[wl];
    const int nbig = *big_count;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nbig; q += warps) {
        const uint32_t r = big_list[q];
        const int64_t b = rp_dup[r];
        const int64_t len64 = rp_dup[r + 1] - b;
        if (len64 > kWarpRowMax) {
            if (lane == 0) huge_list[atomicAdd(huge_count, 1)] = r;
            continue;
        }
        const int len = (int)len64;
        int N = 64;
        while (N < len) N <<= 1;
        for (int i = lane; i < N; i += 32) sm[i] = i < len ? rows[b + i] : 0x7fffffff;
        __syncwarp();
        for (int k = 2; k <= N; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = lane; i < N; i += 32) {
                    const int x = i ^ j;
                    if (x > i) {
                        const int32_t va = sm[i], vb = sm[x];
                        if ((va > vb) == ((i & k) == 0)) {
                            sm[i] = vb;
                            sm[x] = va;
                        }
                    }
                }
                __syncwarp();
            }
        int run = 0;
        for (int i0 = 0; i0 < len; i0 += 32) {
            const int i = i0 + lane;
            const bool uniq = i < len && (i == 0 || sm[i] != sm[i - 1]);
            const unsigned m = __ballot_sync(0xffffffffu, uniq);
            if (uniq) rows[b + run + __popc(m & ((1u << lane) - 1u))] = sm[i];
            run += __popc(m);
        }
        if (lane == 0) deg[r] = (uint32_t)run;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kScanThreads)
row_sort_big_kernel(const int64_t *__restrict__ rp_dup, int32_t *__restrict__ rows, int32_t *__restrict__ scratch,
                    uint32_t *__restrict__ deg, const uint32_t *__restrict__ big_list,
                    const int *__restrict__ big_count, int npass)
{
    __shared__ uint32_t hist[256], base[256], running[256];
    __shared__ uint32_t cnt[kScanThreads / 32][256];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int nbig = *big_count;
    for (int q = blockIdx.x; q < nbig; q += gridDim.x) {
        const uint32_t r = big_list[q];
        const int64_t b = rp_dup[r];
        const uint32_t len = (uint32_t)(rp_dup[r + 1] - b);
        uint32_t *A = reinterpret_cast<uint32_t *>(rows + b), *B = reinterpret_cast<uint32_t *>(scratch + b);
        for (int p = 0; p < npass; ++p) {
            const uint32_t *in = (p & 1) ? B : A;
            uint32_t *out = (p & 1) ? A : B;
            const int shift = 8 * p;
            hist[t] = 0;
            __syncthreads();
            for (uint32_t i = t; i < len; i += kScanThreads) atomicAdd(&hist[(in[i] >> shift) & 0xffu], 1u);
            __syncthreads();
            uint32_t total;
            const uint32_t ex = block_exclusive_scan(hist[t], &total);
            base[t] = ex;
            running[t] = 0;
#pragma unroll
            for (int k = 0; k < kScanThreads / 32; ++k) cnt[k][t] = 0;
            __syncthreads();
            for (uint32_t r0 = 0; r0 < len; r0 += kScanThreads) {
                const uint32_t i = r0 + t;
                const bool live = i < len;
                const uint32_t key = live ? in[i] : 0u;
                const uint32_t dg = live ? ((key >> shift) & 0xffu) : 0x100u;
                const unsigned peers = __match_any_sync(0xffffffffu, dg);
                const uint32_t below = __popc(peers & ((1u << lane) - 1u));
                if (live && below == 0) cnt[w][dg] = __popc(peers);
                __syncthreads();
                if (live) {
                    uint32_t off = base[dg] + running[dg] + below;
                    for (int k = 0; k < w; ++k) off += cnt[k][dg];
                    out[off] = key;
                }
                __syncthreads();
                uint32_t sum = 0;
#pragma unroll
                for (int k = 0; k < kScanThreads / 32; ++k) {
                    sum += cnt[k][t];
                    cnt[k][t] = 0;
                }
                running[t] += sum;
                __syncthreads();
            }
        }
        // the sorted row sits in B after an odd number of passes, else in A; unique entries go to the
        // start of A (through B when the sorted row is in A: in-place compaction would race)
        uint32_t *src = (npass & 1) ? B : A;
        if (!(npass & 1)) {
            for (uint32_t i = t; i < len; i += kScanThreads) B[i] = A[i];
            __syncthreads();
            src = B;
        }
        uint32_t run = 0;
        for (uint32_t r0 = 0; r0 < len; r0 += kScanThreads) {
            const uint32_t i = r0 + t;
            const bool live = i < len;
            const uint32_t v = live ? src[i] : 0u;
            const bool uniq = live && (i == 0 || v != src[i - 1]);
            uint32_t total;
            const uint32_t ex = block_exclusive_scan(uniq ? 1u : 0u, &total);
            if (uniq) A[run + ex] = v;
            run += total;
        }
        if (t == 0) deg[r] = run;
        __syncthreads();
    }
}

// final CSR: row r = the first deg[r] entries of its duplicate-laden row (a thread per row; long rows
// are rare and only cost their own thread)
__global__ void __launch_bounds__(256)
row_compact_kernel(const int64_t *__restrict__ rp_dup, const int32_t *__restrict__ rows,
                   const int64_t *__restrict__ row_ptr, uint64_t n, int32_t *__restrict__ col)
{
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n;
         r += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t src = rp_dup[r], dst = row_ptr[r];
        const int64_t len = row_ptr[r + 1] - dst;
        for (int64_t i = 0; i < len; ++i) col[dst + i] = rows[src + i];
    }
}

// every undirected edge of a symmetric CSR once, as a degenerate triple [a, b, b] (a < b): how the edges
// of an earlier triangulation are carried into the next vf_to_vv call (the reference's connect() only
// ever adds, _vertex.py:116-131).  Order is irrelevant (the CSR build de-duplicates and sorts).
__global__ void __launch_bounds__(256)
csr_edges3_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col, uint64_t n,
                  int32_t *__restrict__ out, unsigned long long *__restrict__ cursor, uint64_t cap)
{
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n;
         r += (uint64_t)gridDim.x * blockDim.x)
        for (int64_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
            const int32_t c = col[k];
            if ((uint64_t)c > r) {
                const unsigned long long pos = atomicAdd(cursor, 1ull);
                if (pos < cap) {
                    out[3 * pos] = (int32_t)r;
                    out[3 * pos + 1] = c;
                    out[3 * pos + 2] = c;
                }
            }
        }
}

// present[r] = 1 iff row r is non-empty: the vertices that exist in the reference's cache
__global__ void __launch_bounds__(256)
row_present_kernel(const int64_t *__restrict__ row_ptr, uint64_t n, uint8_t *__restrict__ present)
{
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n;
         r += (uint64_t)gridDim.x * blockDim.x)
        present[r] = row_ptr[r + 1] > row_ptr[r] ? 1 : 0;
}

// =============================================================================================
// arg-min (find_lowest_vertex, reference _shgo.py:863-880) and nfev (reference _vertex.py:184,347)
// =============================================================================================
struct MinPair {
    double v;
    int64_t i;  // vertex index, -1 = none
    int64_t k;  // tie-break key: the caller's rank of the vertex, else its index
};
__device__ __forceinline__ MinPair min_pair(MinPair a, MinPair b)
{
    if (b.i >= 0 && (a.i < 0 || b.v < a.v || (b.v == a.v && b.k < a.k))) return b;
    return a;
}
__device__ __forceinline__ MinPair block_min(MinPair m)
{
    __shared__ double sv[32];
    __shared__ long long si[32], sk[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MinPair y{__shfl_down_sync(0xffffffffu, m.v, o), __shfl_down_sync(0xffffffffu, m.i, o),
                  __shfl_down_sync(0xffffffffu, m.k, o)};
        m = min_pair(m, y);
    }
    if (lane == 0) {
        sv[w] = m.v;
        si[w] = m.i;
        sk[w] = m.k;
    }
    __syncthreads();
    if (w == 0) {
        m = lane < nw ? MinPair{sv[lane], si[lane], sk[lane]} : MinPair{pos_inf(), -1, 0};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            MinPair y{__shfl_down_sync(0xffffffffu, m.v, o), __shfl_down_sync(0xffffffffu, m.i, o),
                      __shfl_down_sync(0xffffffffu, m.k, o)};
            m = min_pair(m, y);
        }
    }
    return m;  // valid in thread 0
}

__global__ void __launch_bounds__(256)
argmin_partial_kernel(const double *__restrict__ f, const uint8_t *__restrict__ present,
                      const int64_t *__restrict__ rank, uint64_t n, MinPair *__restrict__ partial)
{
    MinPair m{pos_inf(), -1, 0};
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        if (present && !present[i]) continue;
        const double v = ld_nc_f64(f + i);
        if (!(v < pos_inf())) continue;  // +inf and NaN never win
        const int64_t k = rank ? rank[i] : (int64_t)i;
        // ties: the lowest key wins (the reference keeps the first strictly lower vertex in cache
        // order, _shgo.py:863-880: among equal values that is the earliest inserted one)
        if (m.i < 0 || v < m.v || (v == m.v && k < m.k)) {
            m.v = v;
            m.i = (int64_t)i;
            m.k = k;
        }
    }
    m = block_min(m);
    if (threadIdx.x == 0) partial[blockIdx.x] = m;
}

__global__ void __launch_bounds__(1024)
argmin_final_kernel(const MinPair *__restrict__ partial, int nparts, int64_t *idx, double *val)
{
    MinPair m{pos_inf(), -1, 0};
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

int shgo_b200_copy_async(void *dst, const void *src, size_t bytes, void *stream)
{
    if (bytes == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(dst && src, "null pointer");
    SHGO_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, as_stream(stream)));
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
    const uint64_t pairs = dp1 >= 3 ? 3 * S : S;
    const uint64_t nb = (n + kScanBlock - 1) / kScanBlock;
    // duplicate-laden rows, their offsets, degree + cursor / long-row list, block sums, flags
    return align256(pairs * 2 * 4 + 16) + align256((n + 1) * 8) + 2 * align256(n * 4) + align256((nb + 2) * 8) +
           align256((pairs * 2 / kWarpRowMax + 64) * 4) + align256(16);
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
    const uint64_t nb = (n + kScanBlock - 1) / kScanBlock;
    unsigned char *p = static_cast<unsigned char *>(tmp);
    int32_t *rows = reinterpret_cast<int32_t *>(p);
    p += align256(pairs * 2 * 4 + 16);
    int64_t *rp_dup = reinterpret_cast<int64_t *>(p);
    p += align256((n + 1) * 8);
    uint32_t *deg = reinterpret_cast<uint32_t *>(p);
    p += align256(n * 4);
    uint32_t *cursor = reinterpret_cast<uint32_t *>(p);  // fill cursors, then the list of long rows
    p += align256(n * 4);
    uint64_t *sums = reinterpret_cast<uint64_t *>(p);
    p += align256((nb + 2) * 8);
    uint32_t *huge = reinterpret_cast<uint32_t *>(p);  // rows longer than kWarpRowMax (at most 2 pairs / 1024)
    p += align256((pairs * 2 / kWarpRowMax + 64) * 4);
    int *bad = reinterpret_cast<int *>(p);  // [0] out-of-range vertex seen, [1] rows > 32, [2] rows > kWarpRowMax

    SHGO_CUDA_TRY(cudaMemsetAsync(deg, 0, align256(n * 4) * 2, st));  // deg + cursor
    SHGO_CUDA_TRY(cudaMemsetAsync(bad, 0, 3 * sizeof(int), st));
    const unsigned rows_grid = (unsigned)(n / kScanBlock + 1);  // n+1 outputs
    if (S > 0) edge_count_kernel<<<grid_for(S, 256, 8), 256, 0, st>>>(simplices, S, dp1, n, deg, bad);
    block_sum_kernel<uint32_t><<<(unsigned)nb, kScanThreads, 0, st>>>(deg, n, sums, 0);
    scan_sums_kernel<<<1, 1024, 0, st>>>(sums, nb, nnz);
    rowptr_kernel<<<rows_grid, kScanThreads, 0, st>>>(deg, n, sums, rp_dup);
    if (S > 0) {
        edge_fill_kernel<<<grid_for(S, 256, 8), 256, 0, st>>>(simplices, S, dp1, n, rp_dup, cursor, rows, bad);
        // `cursor` is free after the fill: it takes the list of rows longer than a warp; deg becomes
        // the number of DISTINCT entries
        row_sort_small_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(rp_dup, n, rows, deg, cursor, bad + 1);
        int npass = 1;
        while (npass < 4 && (n - 1) >> (8 * npass)) ++npass;  // vertex ids < n: skip all-zero digits
        row_sort_warp_kernel<<<(unsigned)sm_count() * 4, 256, 0, st>>>(rp_dup, rows, deg, cursor, bad + 1, huge, bad + 2);
        row_sort_big_kernel<<<(unsigned)sm_count() * 2, kScanThreads, 0, st>>>(rp_dup, rows, col, deg, huge, bad + 2,
                                                                             npass);
    }
    block_sum_kernel<uint32_t><<<(unsigned)nb, kScanThreads, 0, st>>>(deg, n, sums, 0);
    scan_sums_kernel<<<1, 1024, 0, st>>>(sums, nb, nnz);
    rowptr_kernel<<<rows_grid, kScanThreads, 0, st>>>(deg, n, sums, row_ptr);
    if (S > 0) {
        row_compact_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(rp_dup, rows, row_ptr, n, col);
        flag_bad_kernel<<<1, 1, 0, st>>>(bad, nnz);
    }
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

int shgo_b200_csr_edges3(const int64_t *row_ptr, const int32_t *col, uint64_t n, int32_t *out, uint64_t cap,
                         uint64_t *count, void *stream)
{
    SHGO_REQUIRE(count, "null count");
    cudaStream_t st = as_stream(stream);
    SHGO_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(uint64_t), st));
    if (n == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(row_ptr && (out || cap == 0), "null pointer");
    csr_edges3_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(row_ptr, col, n, out,
                                                          reinterpret_cast<unsigned long long *>(count), cap);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

int shgo_b200_row_present(const int64_t *row_ptr, uint64_t n, uint8_t *present, void *stream)
{
    if (n == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(row_ptr && present, "null pointer");
    row_present_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(row_ptr, n, present);
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
    return shgo_b200_argmin_ranked(f, present, nullptr, n, idx, val, tmp, tmp_bytes, stream);
}

int shgo_b200_argmin_ranked(const double *f, const uint8_t *present, const int64_t *rank, uint64_t n,
                            int64_t *idx, double *val, void *tmp, size_t tmp_bytes, void *stream)
{
    SHGO_REQUIRE(idx && val && tmp, "null pointer");
    SHGO_REQUIRE(f || n == 0, "null f");
    SHGO_REQUIRE(tmp_bytes >= 1024 * sizeof(MinPair), "argmin workspace: need %zu bytes",
                 1024 * sizeof(MinPair));
    cudaStream_t st = as_stream(stream);
    unsigned blocks = grid_for(n ? n : 1, 256, 4);
    if (blocks > 1024) blocks = 1024;
    MinPair *partial = static_cast<MinPair *>(tmp);
    argmin_partial_kernel<<<blocks, 256, 0, st>>>(f, present, rank, n, partial);
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
