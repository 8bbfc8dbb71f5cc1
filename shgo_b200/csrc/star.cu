// star.cu -- K4: star test over an explicit CSR (reference _vertex.py:144-151, 422-426), single-GPU
// and sharded forms.  Integer / fp64-compare work bound by HBM; no tensor cores involved.
//
// Data flow of one tile (32 consecutive rows, owned by one WARP; warps never wait on one another):
//
//   stage   the tile's slice of `col` -- the contiguous range col[row_ptr[r0] .. row_ptr[r0+32]), the
//           bulk of the compulsory traffic -- is copied into the warp's shared-memory buffer
//             BULK  by ONE cp.async.bulk (the TMA engine's 1-D copy, SASS UBLKCP) that completes on an
//                   mbarrier; two buffers per warp, the copy of tile t+1 is issued before tile t is
//                   walked, so every resident warp always has a slice in flight;
//             else  by 16-byte cp.async (LDGSTS) of all lanes into a single buffer.
//           `col` and `row_ptr` pass through once (L2 evict_first), f is kept (evict_last).
//   round A one thread per row: the first 4, then the next 8 neighbours, gathered in two batches.
//           On a random field a row survives k neighbours with probability 1/(k+1): 1 lane in 13 is
//           still alive after this.
//   funnel  the survivors are finished by the WHOLE warp: their remaining neighbours are dealt out to
//           groups of lanes (up to 4 survivors: 8 lanes each, more: 4 each), one gather per lane, one
//           ballot per step -- up to 32 gathers in flight per instruction instead of 2-3 live lanes
//           dragging 16-wide rounds (round 1's kernel: 30 % of its stall samples).  Wider groups (16 / 32
//           lanes for 2 / 1 survivors) are compiled in (MAXLG) but measured slower on the bench graph:
//           8 neighbours per step test the near ones first and spare the row that fails there its
//           far gathers, each a 64-byte DRAM fill (7.40 -> 7.20 ms at 2^28).
//   remote  (sharded, FUSED) rows that passed every local neighbour and have neighbours on other GPUs
//           read those f values in place over NVLink (peer-mapped shards) from inside this kernel, as
//           soon as the owning GPU has published "my shard is written" (a flag per peer).  The loads
//           are issued by the row's own lane at the end of tile t and consumed at the end of tile t+1,
//           so the NVLink round trip hides behind the next tile.  What cannot be finished inline (peer
//           not ready yet, more than 4 remote neighbours) is listed for the phase-2 kernel.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace shgo_b200 {

constexpr int kMaxStarWarps = 16384;  // upper bound on the warps of one launch (candidate segments)
// Candidate workspace written by phase 1 (uint32 words):
//   [0] warps  [1] id-segment capacity  [2], [3] -
//   [4 .. 4+kMaxStarWarps)  candidates per warp
//   id segments    warps x seg_cap   row ids (relative to row0), dense per warp
// (Appending a 32-byte record {row, f, remote neighbour ids} per candidate, so that phase 2 touches
// no row_ptr / col / f of its own, was measured: phase 2 alone 0.18 -> 0.08 ms with all shards local,
// but phase 1 +0.07 ms and the real 8-GPU step 1.38 -> 1.44 ms -- phase 2 is bound by NVLink there.)
constexpr int kCandHeader = 4;
constexpr unsigned kFull = 0xffffffffu;

// ---- cache policies and flavoured loads ----------------------------------------------------------
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
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
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
// gather of a neighbour's f: kept in L2, and a miss fetches 64 instead of 128 bytes from DRAM
// (ncu, round 1: 55.0 -> 47.6 GB read on the bench graph; cudaLimitMaxL2FetchGranularity is ignored)
__device__ __forceinline__ double ld_gather_f64(const double *p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.L2::64B.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
// Peer memory (another GPU's shard, mapped over NVLink).  A plain weak load: a shard is only read
// after its owner's "written" word was seen, every value is read once, and L1 starts clean with the
// launch, so no stale line can be hit.  (System-scope strong loads -- ld.relaxed.sys / ld.acquire.sys,
// SASS LDG.E.STRONG.SYS -- were measured 34 % slower for the WHOLE kernel even when no peer was
// ready and every one of them went to local memory: tools/fused_probe.py, round 2.)
__device__ __forceinline__ double ld_peer_f64(const double *p)
{
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// a peer's "shard written" word: written remotely into LOCAL memory; read at L2 (.cg), where remote
// writes land, so that a line cached in L1 cannot keep answering "not yet".  A late answer only
// sends more candidates to the phase-2 kernel, never a wrong one.
__device__ __forceinline__ uint32_t ld_flag_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// ---- mbarrier + bulk copy (TMA 1-D) --------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(a),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar,
                                              uint64_t pol)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(d),
        "l"(gmem_src), "r"(bytes), "r"(b), "l"(pol)
        : "memory");
}

// ---- per-tile metadata -----------------------------------------------------------------------------
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

// where the slice of a tile sits in `col`: [s0a, s1), s0a = start aligned down to a 16-byte ADDRESS
struct Slice {
    int64_t s0a;
    int len;      // s1 - s0a
    bool staged;  // fits the warp's buffer
};
__device__ __forceinline__ Slice slice_of(const StarMeta &m, const int32_t *col, int cap_w)
{
    const int64_t s0 = __shfl_sync(kFull, m.b, 0), s1 = __shfl_sync(kFull, m.e, 31);
    Slice s;
    s.s0a = s0 - (int64_t)((reinterpret_cast<uintptr_t>(col + s0) >> 2) & 3);
    const int64_t len = s1 - s.s0a;
    s.staged = len <= cap_w;
    s.len = s.staged ? (int)len : 0;
    return s;
}

enum StarMode { kGlobal = 0, kLocal = 1, kLocalEnds = 2 };
constexpr int kMaxFusedShards = 16;

struct StarParams {
    const int64_t *row_ptr;
    const int32_t *col;
    const double *f;  // kGlobal: whole vector; kLocal*: the owned shard (row own0 at index 0)
    const uint8_t *feasible;
    uint64_t row0, rows, own0, own_rows;
    int cap_w;
    uint8_t *is_min;
    uint32_t *cand_ws;
    unsigned long long *nfev;
    // FUSED (sharded, remote neighbours read inline over NVLink)
    const double *const *shards;  // device array of nshards base pointers (own = local HBM); indexing
                                  // a by-value copy in the parameter bank per lane was measured slower
    const uint32_t *ready;        // ready[q] == epoch  <=>  shard q is completely written
    uint32_t epoch;
    uint32_t rows_per_shard;
    int nshards;
    int shard_shift;  // log2(rows_per_shard) when that is a power of two, else -1
};

// One round of the thread-per-row walk on the staged slice: the next W neighbours (32-bit local
// indices; the tail of a row is handled by clamping, so a short row re-tests its last neighbour
// instead of paying for per-entry predicates).  CHECK_OWNER (kLocal): f is this GPU's shard,
// neighbours owned by another GPU are skipped here and remembered in `remote`.
template <int W, bool CHECK_OWNER>
__device__ __forceinline__ bool star_round(const int32_t *__restrict__ sc, const double *__restrict__ f,
                                           double fv, int &k, int last, uint32_t r0, uint32_t nr,
                                           bool &remote, uint64_t keep)
{
    int32_t nb[W];
    double fn[W];
#pragma unroll
    for (int u = 0; u < W; ++u) nb[u] = sc[min(k + u, last)];
#pragma unroll
    for (int u = 0; u < W; ++u) {
        if (CHECK_OWNER) {
            const uint32_t l = (uint32_t)nb[u] - r0;
            const bool mine = l < nr;
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

// The funnel: rows still alive after round A are finished by the whole warp.  `fx` is f addressed
// by the ids found in the slice (kGlobal / kLocalEnds: global ids; kLocal: f - own0 is NOT used,
// the owner test rebases).  MAXLG caps the group width (2^MAXLG lanes per survivor).
template <bool CHECK_OWNER, int MAXLG>
__device__ __forceinline__ void star_funnel(const int32_t *__restrict__ sc, const double *__restrict__ f,
                                            double fv, int &k, int le, bool &ok, bool &remote, uint32_t r0,
                                            uint32_t nr, int lane, uint64_t keep)
{
    unsigned alive = __ballot_sync(kFull, ok && k < le);
    while (alive) {
        const int na = __popc(alive);
        int lg = na <= 1 ? 5 : (na <= 2 ? 4 : (na <= 4 ? 3 : 2));  // lanes per survivor = 2^lg
        if (lg > MAXLG) lg = MAXLG;
        const int groups = 32 >> lg;
        const int grp = lane >> lg, sub = lane & ((1 << lg) - 1);
        // source lane of my group: the grp-th set bit of `alive`
        unsigned m = alive;
        int src = -1;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            if (g < groups) {
                const int s = m ? __ffs(m) - 1 : -1;
                if (g == grp) src = s;
                m &= m - 1;
            }
        }
        const int sl = src < 0 ? 0 : src;
        const int kk = __shfl_sync(kFull, k, sl);
        const int ll = __shfl_sync(kFull, le, sl);
        const double fvv = __shfl_sync(kFull, fv, sl);
        const int idx = kk + sub;
        bool fail = false, rem = false;
        if (src >= 0 && idx < ll) {
            const int32_t nb = sc[idx];
            if (CHECK_OWNER) {
                const uint32_t l = (uint32_t)nb - r0;
                if (l < nr) fail = !(fvv < ld_gather_f64(f + l, keep)); else rem = true;
            } else {
                fail = !(fvv < ld_gather_f64(f + nb, keep));
            }
        }
        const unsigned fb = __ballot_sync(kFull, fail);
        const unsigned rb = CHECK_OWNER ? __ballot_sync(kFull, rem) : 0u;
        // my own row (if alive) was served by group number rank-among-survivors, when that is < groups
        const int rank = __popc(alive & ((1u << lane) - 1u));
        if (((alive >> lane) & 1u) && rank < groups) {
            const unsigned gm = lg == 5 ? kFull : (((1u << (1 << lg)) - 1u) << (rank << lg));
            if (fb & gm) ok = false;
            if (CHECK_OWNER && (rb & gm)) remote = true;
            k += 1 << lg;
        }
        alive = __ballot_sync(kFull, ok && k < le);
    }
}

// MODE kGlobal : `f` is the whole vector, flags are 0 / 1.
// MODE kLocal  : `f` is this GPU's shard (rows own0 .. own0+own_rows); a row that passed every LOCAL
//                neighbour but also has neighbours on other GPUs gets flag 2 ("candidate") and is
//                finished by the remote phase.
// MODE kLocalEnds: the caller promises that in every row the neighbours owned by other GPUs sit at
//                the two ends (ascending rows; the bench graph).  The row is trimmed to its local
//                middle part once and then walked exactly like a single-GPU row.
// BULK         : stage with cp.async.bulk + mbarrier, double buffered (see the head of the file).
// FUSED        : (kLocalEnds only) finish candidates inline over NVLink when their peers are ready.
constexpr int kBulkThreads = 256, kBulkMinBlocks = 5;    // double-buffered variant: same budget
constexpr int kPlainThreads = 256, kPlainMinBlocks = 5;  // <= 51 registers, 40 warps per SM
constexpr int kFusedMinBlocks = 5;
// FUSED: remote reads in flight.  A candidate's remote neighbours are fetched over NVLink by cp.async
// (LDGSTS) straight into a slot of a per-warp ring in shared memory -- no register is held while the
// request crosses the switch -- one commit group per tile, consumed kFusedDepth-1 tiles later
// (cp.async.wait_group).  8 slots per tile and warp (a tile has ~1.2 candidates on the bench graph).
constexpr int kFusedDepth = 4, kFusedSlots = 8, kInline = 4;
constexpr int kSlotDoubles = 2 + kInline;  // {row | -, f of the row, up to 4 neighbour values}
constexpr size_t kFusedRingBytes = (size_t)kFusedDepth * kFusedSlots * kSlotDoubles * sizeof(double);

template <int MODE, int STAGE, bool FUSED, int MAXLG>
__global__ void __launch_bounds__(STAGE == 1 ? kBulkThreads : kPlainThreads,
                                  STAGE == 1 ? kBulkMinBlocks : (FUSED ? kFusedMinBlocks : kPlainMinBlocks))
star_kernel(const StarParams p)
{
    extern __shared__ __align__(16) int32_t star_smem[];
    constexpr bool LOCAL = MODE != kGlobal;
    constexpr bool BULK = STAGE != 0;     // copy engine: cp.async.bulk + mbarrier, else per-lane cp.async
    constexpr bool DOUBLE = STAGE == 1;   // two buffers per warp, next tile's copy in flight during the walk
    constexpr int NBUF = DOUBLE ? 2 : 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int cap_w = p.cap_w;
    static_assert(!FUSED || (STAGE == 2 && MODE == kLocalEnds), "FUSED: bulk staging, remote_at_ends rows");
    int32_t *buf0 = star_smem + (size_t)wid * NBUF * cap_w;
    uint64_t *bars = reinterpret_cast<uint64_t *>(star_smem + (size_t)nwarp * NBUF * cap_w) + wid * 2;
    double *ring = reinterpret_cast<double *>(reinterpret_cast<uint64_t *>(star_smem + (size_t)nwarp * NBUF * cap_w) +
                                              nwarp * 2) + (size_t)wid * (kFusedRingBytes / sizeof(double));
    const int64_t *__restrict__ row_ptr = p.row_ptr;
    const int32_t *__restrict__ col = p.col;
    const double *__restrict__ f = p.f;
    const uint64_t row0 = p.row0, rows = p.rows;
    const uint64_t ntiles = (rows + 31) >> 5;
    const uint64_t gw = (uint64_t)blockIdx.x * nwarp + wid;
    const uint64_t nw = (uint64_t)gridDim.x * nwarp;
    const double *f_rows = LOCAL ? f - p.own0 : f;  // f addressed by global row id
    const uint32_t r0 = (uint32_t)p.own0, nr = (uint32_t)p.own_rows;
    unsigned present_feasible = 0;
    // LOCAL: every warp appends its candidates (flag 2) to a private, dense segment of cand_ws, so
    // that phase 2 needs no compaction pass: [0] = warps, [1] = segment capacity,
    // [2 .. 2+kMaxStarWarps) = per-warp counts, then the segments
    const uint32_t seg_cap = (uint32_t)(((ntiles + nw - 1) / nw) << 5);
    uint32_t *seg = (LOCAL && p.cand_ws) ? p.cand_ws + kCandHeader + kMaxStarWarps + gw * (uint64_t)seg_cap : nullptr;
    uint32_t ncand = 0;
    const uint64_t stream_pol = l2_evict_first_policy(), keep = l2_evict_last_policy();

    if (BULK) {
        if (lane == 0) {
            mbar_init(bars, 1);
            mbar_init(bars + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }
    unsigned phase_bits = 0;  // parity of the next wait on buffer 0 / 1 (bits 0 / 1)

    // issue the copy of one tile's slice into buffer `pb`
    // (always commits exactly one cp.async group, so that wait_group counts tiles)
    auto stage = [&](const Slice &s, int pb) {
        if (s.staged) {
            int32_t *dst = buf0 + pb * cap_w;
            const int32_t *src = col + s.s0a;
            const int nvec = s.len >> 2;  // whole 16-byte vectors
            if (BULK) {
                if (lane == 0 && nvec > 0) {
                    mbar_expect_tx(bars + pb, (unsigned)nvec << 4);
                    bulk_copy_g2s(dst, src, (unsigned)nvec << 4, bars + pb, stream_pol);
                }
            } else {
                for (int v = lane; v < nvec; v += 32) cp_async16(dst + (v << 2), src + (v << 2), stream_pol);
            }
            const int t = (nvec << 2) + lane;  // the <= 3 words behind the last whole vector
            if (FUSED) {  // cp.async groups count the remote reads there: plain load + store
                if (t < s.len) dst[t] = __ldg(src + t);
            } else if (t < s.len) {
                cp_async4(dst + t, src + t);
            }
        }
        if (!FUSED) cp_async_commit();
    };

    // FUSED: ring of pending remote reads, see kFusedDepth
    int head = 0;
    if (FUSED) {
        for (int i = lane; i < kFusedDepth * kFusedSlots; i += 32)
            reinterpret_cast<uint32_t *>(ring + i * kSlotDoubles)[0] = 0xffffffffu;
        __syncwarp();
    }
    auto resolve_stage = [&](int st) {  // the values of every slot of stage `st` have arrived
        if (lane < kFusedSlots) {
            double *sl = ring + (st * kFusedSlots + lane) * kSlotDoubles;
            const uint32_t row = reinterpret_cast<uint32_t *>(sl)[0];
            if (row != 0xffffffffu) {
                const double fvs = sl[1];
                bool good = true;
#pragma unroll
                for (int u = 0; u < kInline; ++u) good = good && (fvs < sl[2 + u]);
                p.is_min[row] = good ? 1 : 0;
                reinterpret_cast<uint32_t *>(sl)[0] = 0xffffffffu;
            }
        }
    };

    // FUSED: which peers have published this iteration's shard (bit q), re-read once per tile by
    // lanes q < nshards until every bit is up
    unsigned ready_mask = 0;
    const unsigned all_shards = p.nshards >= 32 ? kFull : ((1u << p.nshards) - 1u);

    StarMeta cur = star_meta(row_ptr, f_rows, row0, rows, gw, ntiles, lane, stream_pol, keep);
    StarMeta nxt;
    Slice cs = slice_of(cur, col, cap_w), ns;
    int pb = 0;
    if (DOUBLE) {
        stage(cs, 0);
        nxt = star_meta(row_ptr, f_rows, row0, rows, gw + nw, ntiles, lane, stream_pol, keep);
    }
    for (uint64_t tile = gw; tile < ntiles; tile += nw) {
        const uint64_t lr = (tile << 5) + lane;
        const int64_t b = cur.b, e = cur.e;
        const double fv = cur.fv;
        const bool valid = cur.valid;
        const Slice s = cs;
        StarMeta nn;
        if (DOUBLE) {
            // the next tile's slice goes into the other buffer before this one is walked
            ns = slice_of(nxt, col, cap_w);
            stage(ns, pb ^ 1);
            nn = star_meta(row_ptr, f_rows, row0, rows, tile + 2 * nw, ntiles, lane, stream_pol, keep);
        } else {
            // (fetching the metadata TWO tiles ahead was measured slower: 7.35 -> 8.1 ms at 2^28)
            stage(s, 0);
            nxt = star_meta(row_ptr, f_rows, row0, rows, tile + nw, ntiles, lane, stream_pol, keep);
        }
        uint32_t flag_word = 0;
        const bool poll = FUSED && ready_mask != all_shards;  // warp-uniform
        if (poll && lane < p.nshards) flag_word = ld_flag_u32(p.ready + lane);
        bool ok = valid && e > b;
        bool remote = false;
        if (p.nfev != nullptr && ok && (p.feasible == nullptr || p.feasible[lr])) ++present_feasible;
        int k0 = 0, kt = 0, let = 0, le0 = 0;  // FUSED: the trimmed-off ends [k0, kt) and [let, le0)
        const int32_t *sc = buf0 + pb * cap_w;
        if (s.staged) {
            if (DOUBLE) cp_async_wait<1>();  // this tile's tail words (the newest group is the next tile's)
            else if (!FUSED) cp_async_wait<0>();
            if (BULK && (s.len >> 2) > 0) {
                mbar_wait(bars + pb, (phase_bits >> pb) & 1u);
                phase_bits ^= 1u << pb;
            }
            __syncwarp();
            int k = (int)(b - s.s0a);
            int le = (int)(e - s.s0a);
            if (MODE == kLocalEnds) {
                k0 = k, le0 = le;
                while (k < le && (uint32_t)sc[k] - r0 >= nr) ++k;
                while (le > k && (uint32_t)sc[le - 1] - r0 >= nr) --le;
                kt = k, let = le;
                remote = k != k0 || le != le0;
            }
            const int last = le - 1;
            // kLocalEnds walks global ids on f_rows like the single-GPU kernel; kLocal rebases per entry
            const double *fx = MODE == kLocal ? f : f_rows;
            constexpr bool CHK = MODE == kLocal;
            if (ok && k < le) ok = star_round<4, CHK>(sc, fx, fv, k, last, r0, nr, remote, keep);
            if (ok && k < le) ok = star_round<8, CHK>(sc, fx, fv, k, last, r0, nr, remote, keep);
            star_funnel<CHK, MAXLG>(sc, fx, fv, k, le, ok, remote, r0, nr, lane, keep);
        } else {
            for (int64_t kk = b; kk < e && ok; ++kk) {
                const int32_t w = __ldg(col + kk);
                if (LOCAL) {
                    const uint32_t l = (uint32_t)w - r0;
                    if (l < nr) ok = fv < ld_nc_f64(f + l); else remote = true;
                } else {
                    ok = fv < ld_nc_f64(f + w);
                }
            }
        }
        bool candidate = LOCAL && valid && ok && remote;
        bool written_later = false;
        if (FUSED) {
            if (poll) ready_mask = __ballot_sync(kFull, lane < p.nshards && (int32_t)(flag_word - p.epoch) >= 0);
            const int nlo = kt - k0, nrem = nlo + (le0 - let);
            uint32_t w[kInline];
            bool rdy = candidate && s.staged && nrem <= kInline;
            if (rdy) {
#pragma unroll
                for (int u = 0; u < kInline; ++u) {
                    w[u] = u < nrem ? (uint32_t)sc[u < nlo ? k0 + u : let + (u - nlo)] : 0u;
                    const uint32_t q = p.shard_shift >= 0 ? (w[u] >> p.shard_shift) : (w[u] / p.rows_per_shard);
                    // (flag >= epoch: a peer may already be one iteration ahead; its buffer of THIS
                    // iteration is intact until it has passed the next barrier, which waits for this GPU)
                    rdy = rdy && (u >= nrem || ((ready_mask >> q) & 1u));
                }
            }
            const unsigned cb = __ballot_sync(kFull, rdy);
            const int slot = __popc(cb & ((1u << lane) - 1u));
            if (rdy && slot < kFusedSlots) {
                double *sl = ring + (head * kFusedSlots + slot) * kSlotDoubles;
                reinterpret_cast<uint32_t *>(sl)[0] = (uint32_t)lr;
                sl[1] = fv;
#pragma unroll
                for (int u = 0; u < kInline; ++u) {
                    if (u < nrem) {
                        const uint32_t q = p.shard_shift >= 0 ? (w[u] >> p.shard_shift) : (w[u] / p.rows_per_shard);
                        cp_async8(sl + 2 + u, p.shards[q] + (w[u] - q * p.rows_per_shard));
                    } else {
                        sl[2 + u] = pos_inf();
                    }
                }
                candidate = false;  // finished here, not by phase 2
                written_later = true;
            }
            cp_async_commit();
            cp_async_wait<kFusedDepth - 1>();  // the group committed kFusedDepth-1 tiles ago is complete
            __syncwarp();
            resolve_stage((head + 1) % kFusedDepth);
            head = (head + 1) % kFusedDepth;
        }
        if (valid && !written_later) p.is_min[lr] = ok ? (candidate ? 2 : 1) : 0;
        if (LOCAL && seg != nullptr) {
            const unsigned m = __ballot_sync(kFull, candidate);
            const uint32_t pos = ncand + __popc(m & ((1u << lane) - 1u));
            if (candidate) seg[pos] = (uint32_t)lr;
            ncand += __popc(m);
        }
        __syncwarp();  // the buffer is rewritten by a later tile
        if (DOUBLE) {
            cur = nxt;
            cs = ns;
            nxt = nn;
            pb ^= 1;
        } else {
            cur = nxt;
            cs = slice_of(cur, col, cap_w);
        }
    }
    if (DOUBLE) cp_async_wait<0>();
    if (FUSED) {
        cp_async_wait<0>();
        __syncwarp();
        for (int st = 0; st < kFusedDepth; ++st) resolve_stage(st);
    }
    if (LOCAL && seg != nullptr && lane == 0) {
        p.cand_ws[kCandHeader + gw] = ncand;
        if (gw == 0) {
            p.cand_ws[0] = (uint32_t)nw;
            p.cand_ws[1] = seg_cap;
        }
    }
    if (p.nfev != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) present_feasible += __shfl_down_sync(kFull, present_feasible, o);
        if (lane == 0 && present_feasible) atomicAdd(p.nfev, (unsigned long long)present_feasible);
    }
}

// ---- multi-GPU, phase 2 ---------------------------------------------------------------------------
// f lives in N equal shards, one per GPU, all mapped into this process over NVLink (CUDA IPC).
// One thread per CANDIDATE row (a few percent of the rows, fewer after the fused kernel) re-reads
// its neighbour list and tests only the neighbours owned by other GPUs, reading their f in place
// through the peer pointers.
struct ShardedF {
    const double *const *shards;  // device array of nshards base pointers (own shard = local HBM)
    uint32_t rows_per_shard;
    // optional conservative filter: block_min[w >> 5] = min of f over the 32 vertices around w, for
    // ALL shards, in LOCAL memory (gathered in bulk while phase 1 runs).  fv < block_min proves
    // fv < f[w] without the 8-byte read over NVLink; a candidate is the minimum of its ~25 local
    // neighbours, so it also undercuts the minimum of 32 unrelated values about 45 % of the time.
    const double *block_min;
    int plog;  // log2 of the evaluation kernel's points per thread: fixes which 32 vertices form a group
    __device__ __forceinline__ bool needs_read(double fv, int32_t w) const
    {
        if (block_min == nullptr) return true;
        const uint32_t u = (uint32_t)w, tl = 8 + plog;
        const uint32_t g = ((u >> tl) << (tl - 5)) + ((u & 255u) >> (5 - plog));
        return !(fv < __ldg(block_min + g));
    }
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
        // Promise of the caller: the neighbours owned by other GPUs sit at the two ENDS of the row.
        // Read four entries from each end together and fetch the remote ones' values over NVLink:
        // two latency round trips and two sectors of `col` per candidate instead of a walk over the
        // whole row.  A side whose four entries are all remote is continued inwards.
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
            for (int u = 0; u < 8; ++u) use[u] = use[u] && facc.needs_read(fv, w[u]);
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
            for (int u = 0; u < 8; ++u)
                fn[u] = ((uint32_t)w[u] - o0 < on || !facc.needs_read(fv, w[u])) ? pos_inf() : facc(w[u]);
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
        const uint32_t n = cand_ws[kCandHeader + sw];
        const uint32_t *seg = cand_ws + kCandHeader + kMaxStarWarps + sw * (uint64_t)seg_cap;
        for (uint32_t i = lane; i < n; i += 32)
            star_remote_one((uint64_t)seg[i], row_ptr, col, facc, f_own, own0, own_rows, row0, remote_at_ends,
                            is_min);
    }
}

// "my shard of f is written": one word per (peer, me), system-scope release so that a peer that
// observes the epoch also observes the f values written by the kernels that preceded this one
__global__ void publish_ready_kernel(uint32_t *const *flags, int npeers, int me, uint32_t epoch)
{
    const int q = threadIdx.x;
    if (q < npeers) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags[q] + me), "r"(epoch) : "memory");
    }
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace shgo_b200

using namespace shgo_b200;

// ---- launch -------------------------------------------------------------------------------------------
// Experiment switch (tools/star_variants.py): SHGO_B200_STAR = "<bulk 0/1>,<warps per CTA>,<maxlg>"
struct StarVariant {
    int bulk, warps, maxlg;
};
static StarVariant star_variant()
{
    static StarVariant v = [] {
        StarVariant d{2, 0, 3};  // bulk single-buffered staging, at most 8 lanes per funnel survivor
        if (const char *s = getenv("SHGO_B200_STAR")) {
            int a = d.bulk, b = d.warps, c = d.maxlg;
            sscanf(s, "%d,%d,%d", &a, &b, &c);
            d = StarVariant{a, b, c};
        }
        return d;
    }();
    return v;
}

static long star_cap(uint64_t rows, uint64_t nnz_hint)
{
    // staging slice per warp: the 32 rows of a tile, 1.125 x the mean degree (+ alignment slack)
    double mean_deg = nnz_hint ? (double)nnz_hint / (double)rows : 32.0;
    long cap = (long)(mean_deg * 32.0 * 1.125) + 16;
    if (cap < 256) cap = 256;
    if (cap > 6144) cap = 6144;  // 24 KB per warp at most
    return (cap + 3) & ~3L;
}

template <int MODE, int STAGE, bool FUSED, int MAXLG>
static int launch_star_t(const StarParams &p0, uint64_t nnz_hint, int warps_req, cudaStream_t st)
{
    StarParams p = p0;
    const long cap = star_cap(p.rows, nnz_hint);
    p.cap_w = (int)cap;
    constexpr int NBUF = STAGE == 1 ? 2 : 1;
    constexpr bool BULK = STAGE == 1;     // the launch-bounds class
    const size_t per_warp = (size_t)cap * NBUF * sizeof(int32_t) + 16 + (FUSED ? kFusedRingBytes : 0);  // + mbarriers
    auto kern = star_kernel<MODE, STAGE, FUSED, MAXLG>;
    // as many warps per SM as shared memory (228 KB, 1 KB of it reserved per CTA) and the register
    // budget of the launch bounds allow, in equal CTAs
    constexpr int kMaxWpc = (BULK ? kBulkThreads : kPlainThreads) / 32;
    constexpr int kMaxCtas = BULK ? kBulkMinBlocks : (FUSED ? kFusedMinBlocks : kPlainMinBlocks);
    int ctas = 1, wpc = 1;
    for (int c = 1; c <= kMaxCtas; ++c) {
        long w = (long)((228 * 1024 - c * 1024) / c / per_warp);
        if (w > kMaxWpc) w = kMaxWpc;
        if (w >= 1 && c * w > ctas * wpc) ctas = c, wpc = (int)w;
    }
    if (warps_req > 0 && warps_req < ctas * wpc) {  // experiment switch: fewer resident warps
        ctas = (warps_req + kMaxWpc - 1) / kMaxWpc;
        wpc = warps_req / ctas;
    }
    const size_t smem = per_warp * wpc;
    static bool attr_set[3][3][2][6] = {};
    if (!attr_set[MODE][STAGE][FUSED][MAXLG]) {
        SHGO_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[MODE][STAGE][FUSED][MAXLG] = true;
    }
    int occ = 0;
    SHGO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpc * 32, smem));
    if (occ < 1) occ = 1;
    if (occ > ctas) occ = ctas;
    const uint64_t ntiles = (p.rows + 31) / 32;
    uint64_t grid = (uint64_t)sm_count() * occ;
    const uint64_t need = (ntiles + wpc - 1) / wpc;
    if (grid > need) grid = need;
    if (grid * wpc > (uint64_t)kMaxStarWarps) grid = kMaxStarWarps / wpc;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, wpc * 32, smem, st>>>(p);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

// max_warps > 0: at most that many resident warps per SM (leaves room for a kernel on another stream)
template <int MODE, bool FUSED>
static int launch_star(const StarParams &p, uint64_t nnz_hint, cudaStream_t st, int max_warps = 0,
                       bool accumulate_nfev = false)
{
    if (p.nfev && !accumulate_nfev) SHGO_CUDA_TRY(cudaMemsetAsync(p.nfev, 0, sizeof(uint64_t), st));
    if (p.rows == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(p.row_ptr && p.f && p.is_min, "null pointer");
    SHGO_REQUIRE((reinterpret_cast<uintptr_t>(p.col) & 3) == 0, "col must be 4-byte aligned");
    StarVariant v = star_variant();
    if (FUSED) v.bulk = 2;  // the fused kernel's cp.async groups belong to the remote reads
    if (max_warps > 0 && (v.warps == 0 || max_warps < v.warps)) v.warps = max_warps;
#define SHGO_STAR_CASE(S)                                                                      \
    if (v.bulk == S) {                                                                         \
        if (v.maxlg >= 5) return launch_star_t<MODE, S, FUSED, 5>(p, nnz_hint, v.warps, st);   \
        if (v.maxlg == 4) return launch_star_t<MODE, S, FUSED, 4>(p, nnz_hint, v.warps, st);   \
        return launch_star_t<MODE, S, FUSED, 3>(p, nnz_hint, v.warps, st);                     \
    }
    if constexpr (!FUSED) {
        SHGO_STAR_CASE(0)
        SHGO_STAR_CASE(1)
    }
    SHGO_STAR_CASE(2)
#undef SHGO_STAR_CASE
    return fail(SHGO_B200_ERR_ARG, "SHGO_B200_STAR: unknown staging variant %d", v.bulk);
}

extern "C" {

int shgo_b200_star_csr(const int64_t *row_ptr, const int32_t *col, const double *f, uint64_t row0,
                       uint64_t rows, uint64_t nnz_hint, uint8_t *is_min, void *stream)
{
    StarParams p{};
    p.row_ptr = row_ptr, p.col = col, p.f = f, p.row0 = row0, p.rows = rows, p.is_min = is_min;
    return launch_star<kGlobal, false>(p, nnz_hint, as_stream(stream));
}

int shgo_b200_star_csr_nfev(const int64_t *row_ptr, const int32_t *col, const double *f,
                            const uint8_t *feasible, uint64_t row0, uint64_t rows, uint64_t nnz_hint,
                            uint8_t *is_min, uint64_t *nfev, void *stream)
{
    SHGO_REQUIRE(nfev, "null nfev");
    StarParams p{};
    p.row_ptr = row_ptr, p.col = col, p.f = f, p.feasible = feasible, p.row0 = row0, p.rows = rows;
    p.is_min = is_min, p.nfev = reinterpret_cast<unsigned long long *>(nfev);
    return launch_star<kGlobal, false>(p, nnz_hint, as_stream(stream));
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
    StarParams p{};
    p.row_ptr = row_ptr, p.col = col, p.f = f_own, p.feasible = feasible, p.row0 = row0, p.rows = rows;
    p.own0 = own0, p.own_rows = own_rows, p.is_min = flags, p.cand_ws = static_cast<uint32_t *>(cand_ws);
    p.nfev = reinterpret_cast<unsigned long long *>(nfev);
    const int max_warps = (remote_at_ends >> 8) & 0xff;
    const bool acc = (remote_at_ends & 2) != 0;
    if (remote_at_ends & 1) return launch_star<kLocalEnds, false>(p, nnz_hint, as_stream(stream), max_warps, acc);
    return launch_star<kLocal, false>(p, nnz_hint, as_stream(stream), max_warps, acc);
}

int shgo_b200_star_csr_fused(const int64_t *row_ptr, const int32_t *col, const double *const *f_shards,
                             const double *f_own, const uint8_t *feasible, int nshards,
                             uint64_t rows_per_shard, int my_shard, uint64_t row0, uint64_t rows,
                             uint64_t nnz_hint, const uint32_t *ready, uint32_t epoch, uint8_t *flags,
                             void *cand_ws, size_t cand_ws_bytes, uint64_t *nfev, void *stream)
{
    SHGO_REQUIRE(f_shards && f_own && ready && cand_ws, "null pointer");
    SHGO_REQUIRE(nshards >= 1 && nshards <= kMaxFusedShards && my_shard >= 0 && my_shard < nshards,
                 "bad shard index (the fused kernel takes at most %d shards)", kMaxFusedShards);
    SHGO_REQUIRE(rows_per_shard >= 1 && rows_per_shard * (uint64_t)nshards <= 0x7fffffffull, "bad shard size");
    const uint64_t own0 = (uint64_t)my_shard * rows_per_shard;
    SHGO_REQUIRE(row0 >= own0 && row0 + rows <= own0 + rows_per_shard, "classified rows outside the owned shard");
    if (cand_ws_bytes < shgo_b200_star_local_workspace(rows))
        return fail(SHGO_B200_ERR_WORKSPACE, "candidate workspace: need %zu bytes, got %zu",
                    shgo_b200_star_local_workspace(rows), cand_ws_bytes);
    StarParams p{};
    p.row_ptr = row_ptr, p.col = col, p.f = f_own, p.feasible = feasible, p.row0 = row0, p.rows = rows;
    p.own0 = own0, p.own_rows = rows_per_shard, p.is_min = flags, p.cand_ws = static_cast<uint32_t *>(cand_ws);
    p.nfev = reinterpret_cast<unsigned long long *>(nfev);
    p.shards = f_shards, p.ready = ready, p.epoch = epoch, p.rows_per_shard = (uint32_t)rows_per_shard;
    p.nshards = nshards;
    p.shard_shift = -1;
    for (int b = 0; b < 31; ++b)
        if (rows_per_shard == (1ull << b)) p.shard_shift = b;
    return launch_star<kLocalEnds, true>(p, nnz_hint, as_stream(stream));
}

int shgo_b200_publish_ready(uint32_t *const *peer_flags, int npeers, int me, uint32_t epoch, void *stream)
{
    SHGO_REQUIRE(peer_flags && npeers >= 1 && npeers <= 32 && me >= 0, "bad arguments");
    publish_ready_kernel<<<1, 32, 0, as_stream(stream)>>>(peer_flags, npeers, me, epoch);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

size_t shgo_b200_star_local_workspace(uint64_t rows)
{
    // header + per-warp counts + per-warp id segments (each rounded up to whole tiles of 32 rows)
    return align256((kCandHeader + (size_t)kMaxStarWarps + rows + 32ull * kMaxStarWarps) * sizeof(uint32_t));
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
    return shgo_b200_star_csr_remote_filtered(row_ptr, col, f_shards, f_own, nshards, rows_per_shard, my_shard,
                                              row0, rows, remote_at_ends, flags, cand_ws, cand_cap, tmp,
                                              tmp_bytes, nullptr, 0, stream);
}

int shgo_b200_star_csr_remote_filtered(const int64_t *row_ptr, const int32_t *col,
                                       const double *const *f_shards, const double *f_own, int nshards,
                                       uint64_t rows_per_shard, int my_shard, uint64_t row0, uint64_t rows,
                                       int remote_at_ends, uint8_t *flags, const void *cand_ws,
                                       uint64_t cand_cap, void *tmp, size_t tmp_bytes,
                                       const double *block_min, int layout_plog, void *stream)
{
    SHGO_REQUIRE(block_min == nullptr || (layout_plog >= 0 && layout_plog <= 5), "bad group layout");
    cudaStream_t st = as_stream(stream);
    if (rows == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(row_ptr && f_shards && f_own && flags && (tmp || cand_ws), "null pointer");
    SHGO_REQUIRE(nshards >= 1 && my_shard >= 0 && my_shard < nshards, "bad shard index");
    SHGO_REQUIRE(rows_per_shard >= 1 && rows_per_shard <= 0x7fffffffull, "bad shard size");
    const uint64_t own0 = (uint64_t)my_shard * rows_per_shard;
    SHGO_REQUIRE(row0 >= own0 && row0 + rows <= own0 + rows_per_shard, "classified rows outside the owned shard");
    if (cand_ws) {  // phase 1 already listed the candidates
        ShardedF fa{f_shards, (uint32_t)rows_per_shard, block_min, layout_plog};
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
    ShardedF facc{f_shards, (uint32_t)rows_per_shard, block_min, layout_plog};
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

}  // extern "C"
