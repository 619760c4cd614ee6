// query_kernels.cuh -- batch overlap queries on the device-resident superset index.
//
// Every query  [qs, qe]  is answered exactly as the reference's backward walk
// (superintervals.hpp:551-579 / 651-825, c_superintervals.h:575-608 / 729-756):
//     i = upper_bound(qe);  while i != NONE:  hit(ends[i] >= qs) ? emit, --i : i = branch[i]
// The walk visits blocks of consecutive intervals; inside a visited block EVERY
// hit is emitted (the reference's AVX2/NEON loop counts whole 32-blocks the same
// way, hpp:718-748), and the block's lowest element decides what comes next:
// a hit steps to the element below the block, a miss jumps through branch[].
// Elements skipped by a jump have ends < ends[miss] < qs, so the emitted set is
// { j <= ub(qe) : ends[j] >= qs } in strictly descending j -- bit-exact.
//
// Mapping to the GPU (one warp = one tile of 32 position-sorted queries):
//   phase A  lane-per-query: each lane runs the branch-free upper_bound for its
//            query (lanes of a sorted tile probe the same cache lines), then walks
//            with 128-bit loads of ends (4 intervals per step).
//   phase B  the few lanes whose walk is long are finished warp-cooperatively:
//            32 lanes x 128-bit loads = 128 ends per step, hits counted with
//            __ballot_sync/__popc (the AVX2 movemask/popcnt loop, 4x wider).
#pragma once

#include "common.cuh"
#include "partition.cuh"
#include <limits.h>

namespace sib {

constexpr int QK_THREADS = 256;
constexpr int QK_WARPS = QK_THREADS / 32;

// Ranks tabulated on a regular grid over the index span (build(), well-formed indexes only):
// tab_s[c] = #{ starts < lo + c * 2^shift },  tab_e[c] = #{ ends < lo + c * 2^shift },  c = 0..cells
// (cells + 1 entries; the last grid point lies beyond every start and end, so tab[cells] = n).
struct RankGrid {
    const uint32_t* tab_s;
    const uint32_t* tab_e;
    int32_t lo;
    uint32_t shift;
    uint32_t cells;
};

// Rank cells (build(), well-formed indexes only): the sorted array A (the starts, or all ends
// ascending) cut into cells of 2^shift coordinates; one 32-byte record -- one L2 sector -- per
// cell:   word 0 = #{ A < cell_lo }  (bit 31 set: the cell holds more values than fit, use A),
//         then the cell's values as ascending offsets from cell_lo, 28 x u8 (fmt 1, shift <= 8)
//         or 14 x u16 (fmt 2, shift <= 16), padded with all-ones (never below any offset).
// #{ A < x } is then ONE 256-bit load + a SWAR count, wherever x falls.
struct RankCells {
    const uint4* rec;   // 2 x uint4 per cell, cells + 1 records (the last is a sentinel with word 0 = n)
    int32_t lo;         // A[0]
    uint32_t span;      // A[n-1] - A[0]
    uint32_t shift;
    uint32_t fmt;       // 0 = not built
};

// Pair cells (build(), only when the rank cells outgrow L2): BOTH ranks of a coordinate cell in one 32-byte record,
//   word 0 = #{ starts - 1 < cell_lo }, word 1 = #{ eall < cell_lo }   (bit 31: that side holds more values than fit)
//   words 2-4 = the cell's (starts - 1) values, words 5-7 = its ends, as ascending offsets from cell_lo, padded with
//   all-ones: 24 x 4 bits per side (fmt 4, cells of 16 coordinates) or 12 x 8 bits (fmt 8, cells of <= 256).
// #{starts <= qe} = #{starts - 1 < qe} is read in the cell of qe, #{ends < qs} in the cell of qs: a stabbing or short
// query finds both in ONE sector. A table in HBM is charged per 32-byte sector (tools/hbm_gather.cu: an adjacent
// pair of sectors costs exactly two), so this halves the gather of such queries; range queries whose ends fall in
// different cells read two records, as with the separate tables.
struct PairCells {
    const uint4* rec;   // 2 x uint4 per cell, cells + 1 records (the last is a sentinel with words 0, 1 = n)
    int32_t lo;         // coordinate of cell 0
    uint32_t span;      // largest covered value - lo
    uint32_t shift;
    uint32_t fmt;       // 0 = not built, 4, 8
};

// Rank bits (build(), dense well-formed indexes only; stream_kernels.cuh): the same sorted array A as
// bit maps, one entry per 32 coordinates -- what the streaming count kernel stages in shared memory.
//   t[k]  = { #{A < lo + 32k} | flags (bits 31, 30),  coordinates of word k holding >= 1 value }
//   d2[k] = coordinates of word k holding >= 2 values
struct RankBits {
    const uint2* t;        // nullptr = not built
    const uint32_t* d2;
    int32_t lo;            // A[0]
    uint32_t span;         // A[n-1] - A[0]
    uint32_t nwords;       // ((span + 1) >> 5) + 1; both arrays are padded by >= 4 entries with prefix = n
};

// Stab lists (built on the first CSR fill of a well-formed index): a checkpoint every 2^kshift
// positions; list b holds every interval below position b << kshift that is still open just after
// the start at position (b << kshift) - 1,
//     L(b) = { j < b << kshift : ends[j] > starts[(b << kshift) - 1] },   L(0) empty,
// as records in descending position: 8-byte (position, end) on sparse data, where few records are
// read per query and only the hits fetch their payload, or 16-byte (position, end, value, -) when
// the lists are long (>= 16 records on average: nearly every record is a hit and a gather per hit
// would cost more than carrying the value). Every list is padded to a multiple of four records
// (pad: end = INT_MIN, never a hit), so it starts on a 32-byte boundary and four 8-byte records
// travel in one 256-bit load; one 16-byte header per list holds its first record and its length. Any query whose candidates end at
// top = #{starts < qs} with top >> kshift == b has qs > starts[(b << kshift) - 1], so its hits
// below the checkpoint are exactly the entries of L(b) with end >= qs: one sequential list read
// instead of the branch-array walk's chain of dependent loads.
struct StabLists {
    const uint4* hdr;      // per list: first record (64 bits: x | y << 32), record count (z); lists start on 32-byte boundaries
    const void* ent;       // int2 (position, end) or int4 (position, end, value, 0) records; nullptr = not built
    const void* entv;      // 8-byte lists only: the same records as (VALUE, end) -- search_values then needs no payload gather per hit
    uint32_t rec16;        // 1: 16-byte records
    uint32_t kshift;
    uint32_t nlists;       // (n >> kshift) + 1
};

__device__ __forceinline__ int4 ld_stab_record(const StabLists& st, uint64_t p) {
    if (st.rec16) return __ldg(reinterpret_cast<const int4*>(st.ent) + p);
    const int2 r = __ldg(reinterpret_cast<const int2*>(st.ent) + p);
    return make_int4(r.x, r.y, 0, 0);
}

struct IndexView {
    const int32_t* starts;
    const int32_t* ends;     // padded to a multiple of 128 entries (pad = INT_MIN)
    const int32_t* values;
    const uint32_t* branch;  // NONE32 = no previous interval reaches this far
    const int32_t* pmax32;   // pmax32[b] = max(ends[0 .. 32b-1]) (INT_MIN for b = 0), from build()
    const int32_t* esort;    // ends with every aligned 32-block sorted ascending (n_pad entries), from build()
    const int32_t* eall;     // all ends sorted ascending (n entries), from build(); only on a well-formed index
    RankGrid grid;           // rank tables for qk_count_rank_kernel; only on a well-formed index
    RankCells cells_s;       // rank cells over starts  } qk_count_cells_kernel; only on a well-formed
    RankCells cells_e;       // rank cells over eall    } index of fewer than 2^31 intervals
    PairCells pair;          // both ranks per record; built when the rank cells do not fit L2 (fmt == 0 otherwise)
    StabLists stab;          // qk_fill_runs_kernel's lists; ent == nullptr -> it walks instead
    RankBits bits_s;         // rank bits over starts  } sk_count_stream_kernel; only on a dense well-formed
    RankBits bits_e;         // rank bits over eall    } index (t == nullptr otherwise)
    uint32_t n;
    uint32_t wellformed;     // 1 when every stored interval has start <= end (checked by build())
    // closed-form count with a few malformed intervals: the rank tables (cells, eall) cover the well-formed
    // intervals only -- rstarts are their starts (== starts when n_mal == 0) -- and the malformed ones are tested per query
    const int32_t* rstarts;
    uint32_t n_mal;
    int32_t mal_s[8], mal_e[8];
};

// Number of starts <= v, by the reference's branch-free halving search (hpp:501-513).
// The trip count depends only on n, so a warp stays converged.
__device__ __forceinline__ uint32_t count_le(const int32_t* __restrict__ starts, uint32_t n, int32_t v) {
    uint32_t pos = 0, len = n;
    while (len > 1) {
        const uint32_t half = len >> 1;
        pos += (ld_nc(starts + pos + half) <= v) ? (len - half) : 0u;
        len = half;
    }
    return pos + ((ld_nc(starts + pos) <= v) ? 1u : 0u);
}

// Number of starts < v (same search, strict comparison).
__device__ __forceinline__ uint32_t count_lt(const int32_t* __restrict__ starts, uint32_t n, int32_t v) {
    uint32_t pos = 0, len = n;
    while (len > 1) {
        const uint32_t half = len >> 1;
        pos += (ld_nc(starts + pos + half) < v) ? (len - half) : 0u;
        len = half;
    }
    return pos + ((ld_nc(starts + pos) < v) ? 1u : 0u);
}

// phase-A policy: keep walking lane-parallel while enough lanes are busy
__device__ __forceinline__ bool keep_lane_phase(uint32_t active_mask, int iter) {
    const int a = __popc(active_mask);
    return a > 16 || (a > 3 && iter < 48);
}

// branch[] entries as ordered keys: NONE32 ("nothing further back can reach") must win
// a minimum, so shift by one: NONE32 -> 0, j -> j+1. min(key) - 1 is the jump target.
__device__ __forceinline__ uint32_t jump_key(uint32_t br) { return br + 1u; }

// One lane-parallel walk step over the aligned 4-block holding i (i != NONE32):
// counts every hit in [base, i]; returns the next index to visit. When the block's
// lowest interval misses, every miss m in the block rules out (branch[m], m), so the
// walk may jump to the smallest of those branch targets.
__device__ __forceinline__ uint32_t walk_step4(const IndexView& ix, uint32_t i, int32_t qs, uint32_t& c) {
    const uint32_t base = i & ~3u;
    const int4 e = ld_nc4(ix.ends + base);
    const uint32_t k = i - base;   // intervals base .. base+k are at or below i
    const bool hx = e.x >= qs, hy = e.y >= qs, hz = e.z >= qs, hw = e.w >= qs;
    c += (hx ? 1u : 0u) + ((k >= 1u && hy) ? 1u : 0u) + ((k >= 2u && hz) ? 1u : 0u) + ((k >= 3u && hw) ? 1u : 0u);
    if (hx) return base - 1u;      // base == 0 wraps to NONE32: walked off the front
    const uint4 br = __ldg(reinterpret_cast<const uint4*>(ix.branch + base));
    uint32_t key = jump_key(br.x);
    if (k >= 1u && !hy) key = min(key, jump_key(br.y));
    if (k >= 2u && !hz) key = min(key, jump_key(br.z));
    if (k >= 3u && !hw) key = min(key, jump_key(br.w));
    return key - 1u;
}

// Warp-cooperative walk of ONE query (warp-uniform bi, bqs): 32 lanes x 128-bit loads =
// 128 ends per step, hits counted with ballot/popc. Returns the hit count.
__device__ __forceinline__ uint32_t walk_warp128(const IndexView& ix, uint32_t bi, int32_t bqs, uint32_t lane) {
    uint32_t bc = 0;
    while (bi != NONE32) {
        const uint32_t base = bi & ~127u;
        const uint32_t j = base + lane * 4u;
        const bool vx = j <= bi, vy = j + 1u <= bi, vz = j + 2u <= bi, vw = j + 3u <= bi;
        int4 e = make_int4(0, 0, 0, 0);
        if (vx) e = ld_nc4(ix.ends + j);
        const bool hx = vx && e.x >= bqs, hy = vy && e.y >= bqs, hz = vz && e.z >= bqs, hw = vw && e.w >= bqs;
        const uint32_t bx = __ballot_sync(FULL_MASK, hx);
        bc += __popc(bx) + __popc(__ballot_sync(FULL_MASK, hy)) + __popc(__ballot_sync(FULL_MASK, hz)) +
              __popc(__ballot_sync(FULL_MASK, hw));
        if (bx & 1u) {             // lane 0's .x is the block's lowest interval: step below the block
            bi = base - 1u;
        } else {                   // smallest branch target over every miss in the block
            uint32_t key = 0xFFFFFFFFu;
            if (vx && !(hx && hy && hz && hw)) {
                const uint4 br = __ldg(reinterpret_cast<const uint4*>(ix.branch + j));
                if (!hx) key = jump_key(br.x);
                if (vy && !hy) key = min(key, jump_key(br.y));
                if (vz && !hz) key = min(key, jump_key(br.z));
                if (vw && !hw) key = min(key, jump_key(br.w));
            }
            bi = __reduce_min_sync(FULL_MASK, key) - 1u;
        }
    }
    return bc;
}

// Hits (ends >= qs) among the 32 intervals of one aligned block, from the block's
// ascending-sorted copy of ends: a 6-probe branch-free rank. The lanes of a tile read
// different words of the same 128-byte line, so each probe is one L1 wavefront.
__device__ __forceinline__ uint32_t block_hits_sorted(const int32_t* __restrict__ es, int32_t qs) {
    uint32_t c = (ld_nc(es + 15) < qs) ? 16u : 0u;
    c += (ld_nc(es + c + 7) < qs) ? 8u : 0u;
    c += (ld_nc(es + c + 3) < qs) ? 4u : 0u;
    c += (ld_nc(es + c + 1) < qs) ? 2u : 0u;
    c += (ld_nc(es + c) < qs) ? 1u : 0u;
    c += (ld_nc(es + c) < qs) ? 1u : 0u;   // c <= 31 here; moves only when all 32 are below qs
    return 32u - c;
}

// Finish the walks of one warp's queries: lane-parallel (4 intervals per step) while enough
// lanes are busy, then warp-cooperative for the stragglers, one query at a time. `i` is the
// lane's next interval to visit (NONE32 = nothing left); hits are added to `c`. Whole-warp call.
__device__ __forceinline__ void walk_tail(const IndexView& ix, uint32_t i, int32_t qs, uint32_t& c, uint32_t lane) {
    uint32_t active = __ballot_sync(FULL_MASK, i != NONE32);
    int iter = 0;
    while (keep_lane_phase(active, iter)) {
        if (i != NONE32) i = walk_step4(ix, i, qs, c);
        active = __ballot_sync(FULL_MASK, i != NONE32);
        ++iter;
    }
    while (active) {
        const int src = __ffs(active) - 1;
        active &= active - 1;
        const uint32_t bi = __shfl_sync(FULL_MASK, i, src);
        const int32_t bqs = __shfl_sync(FULL_MASK, qs, src);
        const uint32_t bc = walk_warp128(ix, bi, bqs, lane);
        if ((int)lane == src) c += bc;
    }
}

// The same, out of line: for kernels whose fast path never walks (inverted queries in the rank kernels are
// rare), so that the walk's registers do not count against the fast path. Returns the hits found.
__device__ __noinline__ uint32_t walk_tail_rare(const int32_t* __restrict__ ends, const uint32_t* __restrict__ branch,
                                                uint32_t i, int32_t qs) {
    IndexView v;
    v.ends = ends;
    v.branch = branch;
    uint32_t c = 0;
    walk_tail(v, i, qs, c, lane_id());
    return c;
}

constexpr uint32_t QK_DENSE_SPAN = 256;      // max spread of upper bounds inside a tile for the sweep
constexpr uint32_t QK_DENSE_MIN_HITS = 8;    // a 32-interval chunk must yield this many hits tile-wide
constexpr int QK_DENSE_MAX_CHUNKS = 256;     // then the sparse tail goes to the branch walk

// ---- count ---------------------------------------------------------------------------
// One warp = one tile of 32 queries (position-sorted, i.e. by start, when the caller or
// the radix sort made them so). Three stages:
//   search  per-lane branch-free upper_bound(qe) (+ lower bound of qs on a well-formed
//           index); a sorted tile probes the same lines.
//   sweep   the tile's queries overlap the same window of the index, so the warp walks
//           that window ONCE, top-down, in aligned 32-interval blocks -- the reference's
//           linear SIMD block count (hpp:718-748) shared by 32 queries. build() keeps a
//           copy of ends sorted inside each block, so a whole block costs each lane a
//           6-probe rank of its qs instead of 32 compares; the partial block at the top
//           is corrected with 128-bit loads of the plain ends. It stops when the prefix
//           maximum of ends says nothing further back can reach any query of the tile,
//           or when blocks stop producing hits.
//   walk    whatever remains (sparse, far-reaching containers) is finished by the
//           branch-array walk: lane-parallel first, warp-cooperative for stragglers.
// Queries arrive as records (partition.cuh): either the caller's arrays as they are
// (rec.idx == nullptr) or the partitioned copy, whose idx says where each count goes.
template <typename CountT>
__global__ void __launch_bounds__(QK_THREADS)
qk_count_kernel(IndexView ix, QueryRecords rec, uint32_t nq, CountT* __restrict__ counts) {
    const uint64_t t64 = (uint64_t)blockIdx.x * QK_THREADS + threadIdx.x;
    const bool live = t64 < nq;
    const uint32_t t = (uint32_t)t64;
    const uint32_t lane = lane_id();

    uint32_t q = t;
    int32_t qs = INT_MAX, qe = 0;
    if (live) {
        qs = ld_stream(rec.qs + t);
        qe = ld_stream(rec.qe + t);
        if (rec.idx) q = ld_stream(rec.idx + t);
    }

    // Candidates are the intervals [0, lim) with start <= qe. When every stored interval is
    // well formed (start <= end), those with start >= qs overlap for certain (end >= start
    // >= qs) -- the run the reference's count_large locates by search (hpp:834-847) --
    // so only [0, top) with start < qs needs its ends tested.
    uint32_t top = 0, c0 = 0;
    if (live && ix.n) {
        const uint32_t lim = count_le(ix.starts, ix.n, qe);
        top = lim;
        if (ix.wellformed) {
            top = min(lim, count_lt(ix.starts, ix.n, qs));
            c0 = lim - top;
        }
    }
    uint32_t c = 0;                              // hits found below top
    uint32_t i = top - 1u;                       // 0 - 1 wraps to NONE32

    // ---- sweep
    const uint32_t top_max = __reduce_max_sync(FULL_MASK, top);
    if (top_max == 0) {
        if (live) counts[q] = (CountT)c0;
        return;
    }
    const uint32_t top_min = __reduce_min_sync(FULL_MASK, top ? top : top_max);   // idle lanes do not widen the span
    if (top_max - top_min <= QK_DENSE_SPAN) {
        const int32_t qs_min = __reduce_min_sync(FULL_MASK, top ? qs : INT_MAX);
        uint32_t pos = top_max - 1u;             // inclusive top of the window, warp-uniform
        int chunks = 0;
        bool finished = false;
        while (true) {
            uint32_t cb = pos & ~31u;
            uint32_t hc, nblk = 1;
            if (pos == cb + 31u && cb + 32u <= top_min) {
                // whole blocks below every lane's limit: rank qs in each block's sorted copy.
                // Four independent searches per round keep loads in flight; blocks below the
                // termination point contribute exact zeros, so overshooting is harmless.
                if (cb >= 96u) {
                    hc = block_hits_sorted(ix.esort + cb, qs) + block_hits_sorted(ix.esort + cb - 32u, qs) +
                         block_hits_sorted(ix.esort + cb - 64u, qs) + block_hits_sorted(ix.esort + cb - 96u, qs);
                    cb -= 96u;
                    nblk = 4;
                } else {
                    hc = block_hits_sorted(ix.esort + cb, qs);
                }
            } else {
                // top of the window: the whole block from its sorted copy, minus the block's
                // intervals at or above this lane's limit (a lane whose limit is below the
                // block subtracts all of it and gets 0)
                const uint32_t lt = min(top, pos + 1u);
                const uint32_t lt_min = max(cb, __reduce_min_sync(FULL_MASK, lt));
                hc = block_hits_sorted(ix.esort + cb, qs);
                for (uint32_t j = lt_min & ~3u; j < cb + 32u; j += 4u) {
                    const int4 e = ld_nc4(ix.ends + j);    // padded to 128: in bounds
                    hc -= ((j >= lt && e.x >= qs) ? 1u : 0u) + ((j + 1u >= lt && e.y >= qs) ? 1u : 0u) +
                          ((j + 2u >= lt && e.z >= qs) ? 1u : 0u) + ((j + 3u >= lt && e.w >= qs) ? 1u : 0u);
                }
            }
            c += hc;
            if (cb == 0) { finished = true; break; }
            pos = cb - 1u;
            if (ld_nc(ix.pmax32 + (cb >> 5)) < qs_min) { finished = true; break; }   // nothing below reaches the tile
            chunks += (int)nblk;
            const uint32_t tot = __reduce_add_sync(FULL_MASK, top ? hc : 0u);
            if ((chunks >= 2 && tot < QK_DENSE_MIN_HITS * nblk) || chunks >= QK_DENSE_MAX_CHUNKS) break;
        }
        if (finished) {
            // lanes with top == 0 swept unmasked chunks: their c is meaningless
            if (live) counts[q] = (CountT)(c0 + (top ? c : 0u));
            return;
        }
        // hand the rest to the walk: everything above pos is already counted
        i = top ? min(i, pos) : NONE32;
        if (i != NONE32 && ld_nc(ix.pmax32 + ((pos + 1u) >> 5)) < qs) i = NONE32;
    }
    if (top == 0) c = 0;

    walk_tail(ix, i, qs, c, lane);
    if (live) counts[q] = (CountT)(c0 + c);
}

// ---- count by rank -----------------------------------------------------------------------
// On an index whose intervals are all well formed (start <= end, checked by build()) and for
// a query with qs <= qe, the walk's answer  #{ j : starts[j] <= qe, ends[j] >= qs }  has a
// closed form: an interval that ends before qs also starts before qs <= qe, so it is among
// the candidates, hence
//       count = #{ starts <= qe }  -  #{ ends < qs }
// -- the reference's upper_bound (hpp:501-513) on starts, and the same search on build()'s
// ascending copy of the ends (IndexView::eall). build() also tabulates both ranks on a regular
// grid over the index span (RankGrid, about 8 intervals per cell), so a query reads the two
// table entries that bracket each rank and finishes with a handful of branch-free halving
// steps inside the cell. Neighbouring queries (the partition, or a position-sorted caller,
// made them neighbours) hit the same table and index lines in L1. Queries with qs > qe
// (quirk Q6: the closed form does not hold) are answered by the branch-array walk in the
// same kernel. Bit-exact with qk_count_kernel on every input it accepts
// (tests/test_gpu_parity.py runs both).
constexpr int QR_THREADS = 256;
constexpr int QR_PER_THREAD = 2;
constexpr uint32_t QR_TILE = QR_THREADS * QR_PER_THREAD;

__device__ __forceinline__ uint32_t grid_cell(const RankGrid& g, int32_t v) {
    const uint32_t d = v > g.lo ? (uint32_t)v - (uint32_t)g.lo : 0u;
    return min(d >> g.shift, g.cells - 1u);
}

template <typename CountT>
__global__ void __launch_bounds__(QR_THREADS)
qk_count_rank_kernel(IndexView ix, QueryRecords rec, uint32_t nq, CountT* __restrict__ counts) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint64_t base = (uint64_t)blockIdx.x * QR_TILE;
    const RankGrid g = ix.grid;

    int32_t qs[QR_PER_THREAD], qe[QR_PER_THREAD];
    uint32_t q[QR_PER_THREAD];
    bool live[QR_PER_THREAD];
#pragma unroll
    for (int j = 0; j < QR_PER_THREAD; ++j) {
        const uint64_t t = base + (uint64_t)j * QR_THREADS + tid;
        live[j] = t < nq;
        qs[j] = live[j] ? ld_stream(rec.qs + t) : 0;
        qe[j] = live[j] ? ld_stream(rec.qe + t) : 0;
        q[j] = (live[j] && rec.idx) ? ld_stream(rec.idx + t) : (uint32_t)t;
    }
    // bracket both ranks from the grid tables: answer in [lo, lo + len]
    uint32_t elo[QR_PER_THREAD], elen[QR_PER_THREAD], slo[QR_PER_THREAD], slen[QR_PER_THREAD];
#pragma unroll
    for (int j = 0; j < QR_PER_THREAD; ++j) {
        const uint32_t ce = grid_cell(g, qs[j]), cs = grid_cell(g, qe[j]);
        elo[j] = ld_nc(g.tab_e + ce);
        elen[j] = ld_nc(g.tab_e + ce + 1u) - elo[j];
        slo[j] = ld_nc(g.tab_s + cs);
        slen[j] = ld_nc(g.tab_s + cs + 1u) - slo[j];
    }
    // halving steps inside the cells, all chains in one warp-uniform loop
    while (true) {
        uint32_t any = 0;
#pragma unroll
        for (int j = 0; j < QR_PER_THREAD; ++j) any |= elen[j] | slen[j];
        if (!__any_sync(FULL_MASK, any != 0)) break;
#pragma unroll
        for (int j = 0; j < QR_PER_THREAD; ++j) {
            if (elen[j]) {
                const uint32_t half = elen[j] >> 1;
                const bool below = ld_nc(ix.eall + elo[j] + half) < qs[j];
                elo[j] += below ? half + 1u : 0u;
                elen[j] = below ? elen[j] - half - 1u : half;
            }
            if (slen[j]) {
                const uint32_t half = slen[j] >> 1;
                const bool below = ld_nc(ix.starts + slo[j] + half) <= qe[j];
                slo[j] += below ? half + 1u : 0u;
                slen[j] = below ? slen[j] - half - 1u : half;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < QR_PER_THREAD; ++j) {
        uint32_t c = slo[j] - elo[j];
        const bool inverted = live[j] && qs[j] > qe[j];
        if (__any_sync(FULL_MASK, inverted)) {
            // qs > qe: the walk's own definition, #{ j <= ub(qe) : ends[j] >= qs }
            uint32_t cw = 0;
            const uint32_t i = inverted ? slo[j] - 1u : NONE32;   // slo = #{starts <= qe}; 0 - 1 wraps to NONE32
            walk_tail(ix, i, qs[j], cw, lane);
            if (inverted) c = cw;
        }
        if (live[j]) counts[q[j]] = (CountT)c;
    }
}

// ---- count by rank cells ---------------------------------------------------------------------
// The same closed form as qk_count_rank_kernel, count = #{starts <= qe} - #{ends < qs}, but each
// rank costs ONE 32-byte sector wherever the query falls (RankCells above), so a batch needs no
// locality at all: a shuffled batch is answered in the caller's order, the queries streamed in and
// the counts streamed out coalesced, the two sectors per query served by L2 (C2: 62 MB of cells).
// Per query: 8 B in, 4 B out, 2 x 32 B sector reads. Queries with qs > qe (quirk Q6) take the
// branch-array walk, as in qk_count_rank_kernel. Bit-exact with qk_count_kernel.
// 128 threads x 16 CTAs/SM measured 3.6 % faster than 256 x 8 on C2 (0.924 against 0.959 ms; 64 x 32 the same, 512 x 4 slower:
// tools/gpu_r02v.sh): the same 2048 resident threads, a finer grain for the block scheduler
#ifndef SIB_QC_THREADS
#define SIB_QC_THREADS 128
#endif
constexpr int QC_THREADS = SIB_QC_THREADS;
#ifndef SIB_QC_PER_THREAD
#define SIB_QC_PER_THREAD 1
#endif
constexpr int QC_PER_THREAD = SIB_QC_PER_THREAD;
constexpr uint32_t QC_TILE = QC_THREADS * QC_PER_THREAD;

struct __align__(32) CellRec { uint32_t w[8]; };

// Fused count + all-gather: besides its own output a count kernel stores every count into up to
// SI_FANOUT_MAX further arrays at the same index -- the other GPUs' copies of the gathered count vector, written over
// NVLink as peer stores while the kernel is still ranking (no separate collective pass). n == 0: plain kernel.
constexpr int SI_FANOUT_MAX = 15;
struct Fanout { uint32_t* p[SI_FANOUT_MAX]; int n; };
template <typename CountT>
__device__ __forceinline__ void fan_store(const Fanout& fan, uint64_t at, uint32_t c) {
    if (sizeof(CountT) == 4) {
        for (int k = 0; k < fan.n; ++k) st_stream(fan.p[k] + at, c);
    }
}

// L2 eviction policies (SIB_QC_HINTS): the rank cells should stay in L2 (evict_last) while the
// query / count streams pass through (evict_first).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ int32_t ld_stream_hint(const int32_t* p, uint64_t pol) {
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_stream_hint(uint32_t* p, uint32_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream_hint(uint64_t* p, uint64_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}

__device__ __forceinline__ CellRec ld_cell(const uint4* p) {
    CellRec r;
#ifdef SIB_QC_NOALLOC
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ CellRec ld_cell_hint(const uint4* p, uint64_t pol) {
    CellRec r;
    asm volatile("ld.global.nc.L2::cache_hint.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p), "l"(pol));
    return r;
}

// where x falls: cell number and offset inside the cell. x is clamped to [lo, hi + 1], so
// everything below the array ranks 0 (cell 0, offset 0) and everything above ranks n.
__device__ __forceinline__ void cell_of(const RankCells& rc, int64_t x, uint32_t& cell, uint32_t& off) {
    int64_t d = x - (int64_t)rc.lo;
    d = d < 0 ? 0 : d;
    const int64_t dmax = (int64_t)rc.span + 1;
    d = d > dmax ? dmax : d;
    cell = (uint32_t)((uint64_t)d >> rc.shift);
    off = (uint32_t)d & ((1u << rc.shift) - 1u);
}

// #{ offsets in the record < off }, SWAR over the 7 payload words. Per lane (byte or half-word) of a
// word a and the replicated threshold t:  a < t  <=>  (~a & t) | (~(a ^ t) & ~d)  at the lane's top bit,
// where d = (a | H) - (t & ~H) never borrows across lanes and has its top bit set iff low(a) >= low(t).
// The seven words' top bits are shifted to distinct positions and counted by ONE popc.
__device__ __forceinline__ uint32_t cell_below(const CellRec& r, uint32_t off, uint32_t fmt) {
#ifdef SIB_QC_NOSWAR   // timing experiment only (wrong counts): what the kernel costs without the in-cell rank
    return (r.w[1] ^ r.w[7]) & 1u;
#endif
#ifdef SIB_QC_SWAR_OLD
    uint32_t acc = 0;
    if (fmt == 1u) {
        const uint32_t t = off * 0x01010101u;
#pragma unroll
        for (int k = 1; k < 8; ++k) acc += __vcmpltu4(r.w[k], t) & 0x01010101u;
        return (acc * 0x01010101u) >> 24;
    }
    const uint32_t t = off * 0x00010001u;
#pragma unroll
    for (int k = 1; k < 8; ++k) acc += __vcmpltu2(r.w[k], t) & 0x00010001u;
    return (acc + (acc >> 16)) & 0xFFFFu;
#else
    const uint32_t H = fmt == 1u ? 0x80808080u : 0x80008000u;
    const uint32_t t = off * (fmt == 1u ? 0x01010101u : 0x00010001u);
    const uint32_t tl = t & ~H;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        const uint32_t a = r.w[k];
        const uint32_t d = (a | H) - tl;
        const uint32_t lt = (~a & t) | (~(a ^ t) & ~d);
        acc |= (lt & H) >> (k - 1);
    }
    return __popc(acc);
#endif
}

// an over-full cell (bit 31 of word 0): #{ A < x } by halving search inside the cell's run of A
__device__ __noinline__ uint32_t cell_overflow_rank(const RankCells& rc, const int32_t* __restrict__ A, uint32_t cell,
                                                    uint32_t base, int64_t x) {
    const uint32_t next = __ldg(reinterpret_cast<const uint32_t*>(rc.rec + 2 * ((size_t)cell + 1))) & 0x7FFFFFFFu;
    uint32_t lo = base, len = next - base;
    while (len) {
        const uint32_t half = len >> 1;
        const bool below = (int64_t)ld_nc(A + lo + half) < x;
        lo += below ? half + 1u : 0u;
        len = below ? len - half - 1u : half;
    }
    return lo;
}

// #{ A < x } from an already loaded record
__device__ __forceinline__ uint32_t cell_rank(const RankCells& rc, const int32_t* __restrict__ A, const CellRec& r,
                                              uint32_t cell, uint32_t off, int64_t x) {
    if (r.w[0] & 0x80000000u) return cell_overflow_rank(rc, A, cell, r.w[0] & 0x7FFFFFFFu, x);
    return r.w[0] + cell_below(r, off, rc.fmt);
}

// #{ A < x } (one dependent load); used by the fill kernel for upper_bound(qe)
__device__ __forceinline__ uint32_t cells_rank_lt(const RankCells& rc, const int32_t* __restrict__ A, int64_t x) {
    uint32_t cell, off;
    cell_of(rc, x, cell, off);
    const CellRec r = ld_cell(rc.rec + 2 * (size_t)cell);
    return cell_rank(rc, A, r, cell, off, x);
}

// ---- pair cells -------------------------------------------------------------------------------------------
__device__ __forceinline__ void pair_cell_of(const PairCells& pc, int32_t x, uint32_t& cell, uint32_t& off) {
    int64_t d = (int64_t)x - (int64_t)pc.lo;
    d = d < 0 ? 0 : d;
    const int64_t dmax = (int64_t)pc.span + 1;
    d = d > dmax ? dmax : d;
    cell = (uint32_t)((uint64_t)d >> pc.shift);
    off = (uint32_t)d & ((1u << pc.shift) - 1u);
}
// #{ offsets of one side < off }: the SWAR compare of cell_below over three words, 4- or 8-bit lanes
template <int FIRST>
__device__ __forceinline__ uint32_t pair_below(const CellRec& r, uint32_t off, uint32_t fmt) {
    const uint32_t H = fmt == 4u ? 0x88888888u : 0x80808080u;
    const uint32_t t = off * (fmt == 4u ? 0x11111111u : 0x01010101u);
    const uint32_t tl = t & ~H;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t a = r.w[FIRST + k];
        const uint32_t d = (a | H) - tl;
        const uint32_t lt = (~a & t) | (~(a ^ t) & ~d);
        acc |= (lt & H) >> k;
    }
    return __popc(acc);
}
// an over-full side: halving search inside the cell's run of A; SUB = 1 ranks A - 1 (the starts side)
template <int SIDE>
__device__ __noinline__ uint32_t pair_overflow_rank(const PairCells& pc, const int32_t* __restrict__ A, uint32_t cell, uint32_t base, int32_t x) {
    const uint32_t next = __ldg(reinterpret_cast<const uint32_t*>(pc.rec + 2 * ((size_t)cell + 1)) + SIDE) & 0x7FFFFFFFu;
    uint32_t lo = base, len = next - base;
    while (len) {
        const uint32_t half = len >> 1;
        const bool below = SIDE == 0 ? (int64_t)ld_nc(A + lo + half) - 1 < (int64_t)x : (int64_t)ld_nc(A + lo + half) < (int64_t)x;
        lo += below ? half + 1u : 0u;
        len = below ? len - half - 1u : half;
    }
    return lo;
}
// SIDE 0: #{ starts <= x } from the record of x's cell; SIDE 1: #{ eall < x }
template <int SIDE>
__device__ __forceinline__ uint32_t pair_rank(const PairCells& pc, const int32_t* __restrict__ A, const CellRec& r, uint32_t cell, uint32_t off, int32_t x) {
    const uint32_t b = r.w[SIDE];
    if (b & 0x80000000u) return pair_overflow_rank<SIDE>(pc, A, cell, b & 0x7FFFFFFFu, x);
    return b + pair_below<SIDE == 0 ? 2 : 5>(r, off, pc.fmt);
}

// ---- pair cells: both ranks of a coordinate cell in one record (layout: PairCells in query_kernels.cuh) ----
// One thread per cell; the ranks at the cell borders come from the rank cells built just before (one sector each,
// neighbouring threads share them). Side 0 stores starts - 1 (so that "< x" reads #{starts <= x}), side 1 the ends.
// FMT 4: 24 four-bit offsets per side, FMT 8: 12 one-byte offsets. overfull[0,1] count the sides that did not fit.
constexpr int PC_BUILD_THREADS = 256;
template <int FMT>
__global__ void __launch_bounds__(PC_BUILD_THREADS)
bk_pair_cells_kernel(RankCells cs, RankCells ce, const int32_t* __restrict__ S, const int32_t* __restrict__ E, uint32_t n,
                     int32_t lo, uint32_t shift, uint32_t cells, uint4* __restrict__ rec, unsigned long long* __restrict__ overfull) {
    constexpr uint32_t SLOTS = FMT == 4 ? 24u : 12u;
    constexpr uint32_t BITS = FMT == 4 ? 4u : 8u;
    constexpr uint32_t PER_WORD = 32u / BITS;
    constexpr uint32_t LANE = (1u << BITS) - 1u;
    const uint64_t stride = (uint64_t)gridDim.x * PC_BUILD_THREADS;
    for (uint64_t c = (uint64_t)blockIdx.x * PC_BUILD_THREADS + threadIdx.x; c <= cells; c += stride) {
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = 0xFFFFFFFFu;
        if (c == cells) {                       // sentinel: the next-base words of the last cell
            w[0] = n; w[1] = n;
        } else {
            const int64_t v0 = (int64_t)lo + (int64_t)(c << shift);
            const int64_t v1 = v0 + ((int64_t)1 << shift);
            // #{S - 1 < v} = #{S < v + 1};  #{E < v}
            const uint32_t s0 = cells_rank_lt(cs, S, v0 + 1), s1 = cells_rank_lt(cs, S, v1 + 1);
            const uint32_t e0 = cells_rank_lt(ce, E, v0), e1 = cells_rank_lt(ce, E, v1);
            w[0] = s0; w[1] = e0;
            if (s1 - s0 > SLOTS) { w[0] |= 0x80000000u; atomicAdd(overfull, 1ull); }
            else
                for (uint32_t k = 0; k < s1 - s0; ++k) {
                    const uint32_t off = (uint32_t)((int64_t)S[s0 + k] - 1 - v0), sh = (k % PER_WORD) * BITS;
                    const uint32_t word = 2u + k / PER_WORD;
#pragma unroll
                    for (uint32_t q = 2; q < 5; ++q) if (q == word) w[q] = (w[q] & ~(LANE << sh)) | (off << sh);
                }
            if (e1 - e0 > SLOTS) { w[1] |= 0x80000000u; atomicAdd(overfull + 1, 1ull); }
            else
                for (uint32_t k = 0; k < e1 - e0; ++k) {
                    const uint32_t off = (uint32_t)((int64_t)E[e0 + k] - v0), sh = (k % PER_WORD) * BITS;
                    const uint32_t word = 5u + k / PER_WORD;
#pragma unroll
                    for (uint32_t q = 5; q < 8; ++q) if (q == word) w[q] = (w[q] & ~(LANE << sh)) | (off << sh);
                }
        }
        rec[2 * c] = make_uint4(w[0], w[1], w[2], w[3]);
        rec[2 * c + 1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// one tile of QC_TILE queries starting at `base` (whole CTA)
template <typename CountT, bool PAIRED>
__device__ __forceinline__ void count_cells_tile(const IndexView& ix, const QueryRecords& rec, uint64_t base, uint32_t nq,
                                                 CountT* __restrict__ counts, const Fanout& fan) {
    const uint32_t tid = threadIdx.x;
    const RankCells cs = ix.cells_s, ce = ix.cells_e;
#ifdef SIB_QC_HINTS
    const uint64_t pol_keep = l2_policy_evict_last(), pol_pass = l2_policy_evict_first();
#endif

    int32_t qs[QC_PER_THREAD], qe[QC_PER_THREAD];
    bool live[QC_PER_THREAD];
#pragma unroll
    for (int j = 0; j < QC_PER_THREAD; ++j) {
        const uint64_t t = base + (uint64_t)j * QC_THREADS + tid;
        live[j] = t < nq;
#ifdef SIB_QC_HINTS
        qs[j] = live[j] ? ld_stream_hint(rec.qs + t, pol_pass) : 0;
        qe[j] = live[j] ? ld_stream_hint(rec.qe + t, pol_pass) : 0;
#else
        qs[j] = live[j] ? ld_stream(rec.qs + t) : 0;
        qe[j] = live[j] ? ld_stream(rec.qe + t) : 0;
#endif
    }
    // all sector loads of the thread's queries in flight together
    uint32_t cell_s[QC_PER_THREAD], off_s[QC_PER_THREAD], cell_e[QC_PER_THREAD], off_e[QC_PER_THREAD];
    CellRec rs[QC_PER_THREAD], re[QC_PER_THREAD];
    const PairCells pc = ix.pair;
    constexpr bool paired = PAIRED;        // the launch is compiled for one table kind
#pragma unroll
    for (int j = 0; j < QC_PER_THREAD; ++j) {
        if (paired) {     // both ranks from the pair table: one sector when qs and qe share a cell
            pair_cell_of(pc, qe[j], cell_s[j], off_s[j]);
            pair_cell_of(pc, qs[j], cell_e[j], off_e[j]);
            rs[j] = ld_cell(pc.rec + 2 * (size_t)cell_s[j]);
            re[j] = rs[j];
            if (cell_e[j] != cell_s[j]) re[j] = ld_cell(pc.rec + 2 * (size_t)cell_e[j]);
            continue;
        }
        cell_of(cs, (int64_t)qe[j] + 1, cell_s[j], off_s[j]);   // #{starts <= qe} = #{starts < qe + 1}
        cell_of(ce, (int64_t)qs[j], cell_e[j], off_e[j]);       // #{ends < qs}
#ifdef SIB_QC_HINTS
        rs[j] = ld_cell_hint(cs.rec + 2 * (size_t)cell_s[j], pol_keep);
        re[j] = ld_cell_hint(ce.rec + 2 * (size_t)cell_e[j], pol_keep);
#else
        rs[j] = ld_cell(cs.rec + 2 * (size_t)cell_s[j]);
        re[j] = ld_cell(ce.rec + 2 * (size_t)cell_e[j]);
#endif
    }
#pragma unroll
    for (int j = 0; j < QC_PER_THREAD; ++j) {
        const uint32_t ns = paired ? pair_rank<0>(pc, ix.rstarts, rs[j], cell_s[j], off_s[j], qe[j])
                                   : cell_rank(cs, ix.rstarts, rs[j], cell_s[j], off_s[j], (int64_t)qe[j] + 1);
        const uint32_t ne = paired ? pair_rank<1>(pc, ix.eall, re[j], cell_e[j], off_e[j], qs[j])
                                   : cell_rank(ce, ix.eall, re[j], cell_e[j], off_e[j], (int64_t)qs[j]);
        uint32_t c = ns - ne;
        uint32_t mal_before = 0;    // malformed intervals with start <= qe: they are candidates too
        for (uint32_t k = 0; k < ix.n_mal; ++k) {           // warp-uniform trip count, usually zero
            const bool cand = ix.mal_s[k] <= qe[j];
            mal_before += cand ? 1u : 0u;
            c += (cand && ix.mal_e[k] >= qs[j]) ? 1u : 0u;
        }
        const bool inverted = live[j] && qs[j] > qe[j];
        if (__any_sync(FULL_MASK, inverted)) {
            // qs > qe: the walk's own definition, #{ j <= ub(qe) : ends[j] >= qs }, over ALL stored intervals
            const uint32_t i = inverted ? ns + mal_before - 1u : NONE32;   // #{starts <= qe} - 1; 0 - 1 wraps to NONE32
            const uint32_t cw = walk_tail_rare(ix.ends, ix.branch, i, qs[j]);
            if (inverted) c = cw;
        }
        const uint64_t t = base + (uint64_t)j * QC_THREADS + tid;
        if (live[j]) {
            if (rec.idx) {
                const uint32_t at = ld_stream(rec.idx + t);
                counts[at] = (CountT)c;
                fan_store<CountT>(fan, at, c);
            } else {
#ifdef SIB_QC_HINTS
                st_stream_hint(counts + t, (CountT)c, pol_pass);
#else
                st_stream(counts + t, (CountT)c);
#endif
                fan_store<CountT>(fan, t, c);
            }
        }
    }
}

#ifndef SIB_QC_MINBLOCKS
#define SIB_QC_MINBLOCKS 16
#endif
template <typename CountT, bool PAIRED = false>
__global__ void __launch_bounds__(QC_THREADS, SIB_QC_MINBLOCKS)
qk_count_cells_kernel(IndexView ix, QueryRecords rec, uint32_t nq, CountT* __restrict__ counts, const __grid_constant__ Fanout fan) {
    // one tile per CTA. Persistent CTAs taking tiles grid-stride (16 per SM) measured slower on C2 shuffled, 1.016 against 0.953 ms
    // (tools/gpu_r02zl.sh), and short-lived CTAs are also what HBM serves random sectors to fastest (tools/hbm_gather.cu)
    count_cells_tile<CountT, PAIRED>(ix, rec, (uint64_t)blockIdx.x * QC_TILE, nq, counts, fan);
}

// ---- count of a MIXED batch over several indexes (mode B: one index per contig) -----------------------
// The reference's callers keep one map per chromosome and send each query to the map of its contig
// (examples/bed-intersect-si.rs:100-123). Here the whole mixed batch (contig id, start, end per query) is answered
// by ONE launch in the caller's order: a thread reads its query's contig id, takes that contig's rank-cell
// descriptors from a table staged in shared memory, and does the same two 32-byte sector gathers as
// qk_count_cells_kernel -- no grouping of the batch by contig, no scatter of the counts back.
// Per query: 12 B in, 4 B out, 2 x 32 B sectors. Ids outside [0, n_contigs) and contigs without an index count 0.
struct MixedEntry {
    RankCells cs, ce;
    PairCells pc;             // fmt != 0: both ranks from the contig's pair table
    const int32_t* rstarts;   // sorted array under cs (starts of the well-formed intervals)
    const int32_t* eall;      // sorted array under ce
    const int32_t* ends;      // for the rare qs > qe walk (quirk Q6)
    const uint32_t* branch;
    uint32_t n;               // 0: no index for this contig
    uint32_t n_mal;
    int32_t mal_s[8], mal_e[8];
};
constexpr int QM_SMEM_ENTRIES = 256;    // tables up to this many contigs are staged in shared memory
constexpr uint32_t QM_PLAIN = 0u;       // every query of the batch is answered here (0 when its contig has no index)
constexpr uint32_t QM_PEER_HOME = 1u;   // this GPU's own slice of a contig-partitioned batch: foreign contigs are skipped
constexpr uint32_t QM_PEER_AWAY = 2u;   // another GPU's slice (peer memory): only the contigs indexed here are answered
constexpr uint32_t QM_FOREIGN = 0xFFFFFFFFu;   // MixedEntry.n_mal of a contig (n == 0) that another GPU answers
constexpr int QM_THREADS = 256;

// the reference's element walk (hpp:551-579) by one lane: only for inverted queries, which well-formed callers never send
__device__ __noinline__ uint32_t walk_scalar(const int32_t* __restrict__ ends, const uint32_t* __restrict__ branch, uint32_t i, int32_t qs) {
    uint32_t c = 0;
    while (i != NONE32) {
        if (ld_nc(ends + i) >= qs) { ++c; i -= 1u; }      // 0 - 1 wraps to NONE32
        else i = __ldg(branch + i);
    }
    return c;
}

// count of one query on the index of its contig (e.n != 0): the two rank gathers of qk_count_cells_kernel in that contig's tables
__device__ __forceinline__ uint32_t mixed_answer(const MixedEntry& e, int32_t qs, int32_t qe) {
    uint32_t cell_s, off_s, cell_e, off_e, ns, ne;
    if (e.pc.fmt != 0u) {
        const PairCells pc = e.pc;
        pair_cell_of(pc, qe, cell_s, off_s);
        pair_cell_of(pc, qs, cell_e, off_e);
        const CellRec rs = ld_cell(pc.rec + 2 * (size_t)cell_s);
        CellRec re = rs;
        if (cell_e != cell_s) re = ld_cell(pc.rec + 2 * (size_t)cell_e);
        ns = pair_rank<0>(pc, e.rstarts, rs, cell_s, off_s, qe);
        ne = pair_rank<1>(pc, e.eall, re, cell_e, off_e, qs);
    } else {
        const RankCells cs = e.cs, ce = e.ce;
        cell_of(cs, (int64_t)qe + 1, cell_s, off_s);
        cell_of(ce, (int64_t)qs, cell_e, off_e);
        const CellRec rs = ld_cell(cs.rec + 2 * (size_t)cell_s);
        const CellRec re = ld_cell(ce.rec + 2 * (size_t)cell_e);
        ns = cell_rank(cs, e.rstarts, rs, cell_s, off_s, (int64_t)qe + 1);
        ne = cell_rank(ce, e.eall, re, cell_e, off_e, (int64_t)qs);
    }
    uint32_t c = ns - ne;
    uint32_t mal_before = 0;
    for (uint32_t k = 0; k < e.n_mal; ++k) {
        const bool cand = e.mal_s[k] <= qe;
        mal_before += cand ? 1u : 0u;
        c += (cand && e.mal_e[k] >= qs) ? 1u : 0u;
    }
    if (qs > qe) c = walk_scalar(e.ends, e.branch, ns + mal_before - 1u, qs);
    return c;
}

// The descriptor table is staged in shared memory once per CTA; per-contig hit totals (the CSR bases of mode B) are summed in
// shared memory and flushed once per CTA. rounds == 0: persistent CTAs (grid = what the device holds), tiles taken grid-stride;
// rounds > 0: a CTA owns `rounds` consecutive tiles and leaves -- gathers that miss L2 are served faster to short-lived CTAs
// (tools/hbm_gather.cu: 47 G against 37 G sectors/s).
#ifndef SIB_QM_MINBLOCKS
#define SIB_QM_MINBLOCKS 5
#endif
constexpr uint32_t QM_MAX_ROUNDS = 8;    // tiles per short-lived CTA at most (the compaction list of the peer modes holds them)
template <typename CountT>
__global__ void __launch_bounds__(QM_THREADS, SIB_QM_MINBLOCKS)
qk_count_mixed_kernel(const MixedEntry* __restrict__ table, uint32_t n_contigs, const int32_t* __restrict__ contig,
                      const int32_t* __restrict__ qs_in, const int32_t* __restrict__ qe_in, uint32_t nq, CountT* __restrict__ counts,
                      unsigned long long* __restrict__ totals, uint32_t rounds, uint32_t peer_mode) {
    extern __shared__ __align__(16) unsigned char qm_smem[];
    const MixedEntry* tab = table;
    unsigned long long* s_tot = nullptr;
    if (n_contigs <= (uint32_t)QM_SMEM_ENTRIES) {
        const uint32_t words = n_contigs * (uint32_t)(sizeof(MixedEntry) / 4);
        uint32_t* dst = reinterpret_cast<uint32_t*>(qm_smem);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(table);
        for (uint32_t k = threadIdx.x; k < words; k += QM_THREADS) dst[k] = __ldg(src + k);
        tab = reinterpret_cast<const MixedEntry*>(qm_smem);
        if (totals) {
            s_tot = reinterpret_cast<unsigned long long*>(qm_smem + (((size_t)words * 4 + 15) & ~(size_t)15));
            for (uint32_t k = threadIdx.x; k < n_contigs; k += QM_THREADS) s_tot[k] = 0ull;
        }
        __syncthreads();
    }
    auto add_total = [&](uint32_t cid, uint32_t c) {
        if (totals && c) {
            if (s_tot) atomicAdd(s_tot + cid, (unsigned long long)c);
            else atomicAdd(totals + cid, (unsigned long long)c);
        }
    };
    if (peer_mode != QM_PLAIN && rounds != 0u) {
        // Contig-partitioned batch read in place (siCountMixedPeerDevice): of this slice only the queries whose contig is indexed
        // HERE are answered -- one in N on N GPUs. A thread per query would leave most lanes idle during the gathers, so the CTA
        // first reads the contig ids of its tiles (4 B per query: all that crosses NVLink for the queries of other GPUs' contigs),
        // packs the positions of its own queries into a shared list, and then answers the list with every lane busy. A query of
        // a contig nobody indexes (or with an id outside the table) gets its 0 from the slice's home GPU.
        __shared__ uint32_t s_list[QM_MAX_ROUNDS * QM_THREADS];     // (contig id << 11) | position inside the CTA's tiles
        __shared__ uint32_t s_n;
        if (threadIdx.x == 0) s_n = 0u;
        __syncthreads();
        const uint64_t base = (uint64_t)blockIdx.x * rounds * QM_THREADS;
        const uint32_t lane = threadIdx.x & 31u;
        // all contig ids of the thread first: up to eight independent loads in flight (a remote slice answers after an NVLink round trip)
        uint32_t cidv[QM_MAX_ROUNDS];
#pragma unroll
        for (uint32_t r = 0; r < QM_MAX_ROUNDS; ++r) {
            const uint64_t t = base + r * QM_THREADS + threadIdx.x;
            cidv[r] = (r < rounds && t < nq) ? (uint32_t)ld_stream(contig + t) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (uint32_t r = 0; r < QM_MAX_ROUNDS; ++r) {
            if (r >= rounds) break;
            const uint32_t i = r * QM_THREADS + threadIdx.x;
            const uint64_t t = base + i;
            const bool live = t < nq;
            const uint32_t cid = cidv[r];
            const bool known = cid < n_contigs;
            const bool mine = live && known && tab[known ? cid : 0u].n != 0u;
            const bool foreign = live && known && tab[known ? cid : 0u].n == 0u && tab[known ? cid : 0u].n_mal == QM_FOREIGN;
            const uint32_t m = __ballot_sync(FULL_MASK, mine);
            if (m) {
                uint32_t at = 0;
                if (lane == (uint32_t)(__ffs(m) - 1)) at = atomicAdd(&s_n, (uint32_t)__popc(m));
                at = __shfl_sync(FULL_MASK, at, __ffs(m) - 1);
                if (mine) s_list[at + __popc(m & lanemask_lt())] = (cid << 11) | i;
            }
            if (live && !mine && !foreign && peer_mode == QM_PEER_HOME) st_stream(counts + t, (CountT)0);
        }
        __syncthreads();
        const uint32_t n_mine = s_n;
        // two list entries per thread and step: the coordinates of both (remote for another GPU's slice) are requested before either is used
        for (uint32_t k = threadIdx.x; k < n_mine; k += 2 * QM_THREADS) {
            const bool two = k + QM_THREADS < n_mine;
            const uint32_t rec0 = s_list[k], rec1 = two ? s_list[k + QM_THREADS] : rec0;
            const uint64_t t0 = base + (rec0 & 2047u), t1 = base + (rec1 & 2047u);
            const int32_t qs0 = ld_stream(qs_in + t0), qe0 = ld_stream(qe_in + t0);
            const int32_t qs1 = ld_stream(qs_in + t1), qe1 = ld_stream(qe_in + t1);
            const uint32_t c0 = mixed_answer(tab[rec0 >> 11], qs0, qe0);
            add_total(rec0 >> 11, c0);
            st_stream(counts + t0, (CountT)c0);
            if (two) {
                const uint32_t c1 = mixed_answer(tab[rec1 >> 11], qs1, qe1);
                add_total(rec1 >> 11, c1);
                st_stream(counts + t1, (CountT)c1);
            }
        }
    } else {
        const uint64_t stride = rounds ? (uint64_t)QM_THREADS : (uint64_t)gridDim.x * QM_THREADS;
        const uint64_t first = rounds ? (uint64_t)blockIdx.x * rounds * QM_THREADS + threadIdx.x : (uint64_t)blockIdx.x * QM_THREADS + threadIdx.x;
        const uint64_t last = rounds ? min((uint64_t)nq, ((uint64_t)blockIdx.x + 1) * rounds * QM_THREADS) : (uint64_t)nq;
        for (uint64_t t = first; t < last; t += stride) {
            const uint32_t cid = (uint32_t)ld_stream(contig + t);
            const bool mine = cid < n_contigs && tab[cid < n_contigs ? cid : 0u].n != 0u;
            if (peer_mode != QM_PLAIN) {   // tables that fit L2 (persistent CTAs): a thread per query, the others' queries skipped
                const bool foreign = cid < n_contigs && tab[cid].n == 0u && tab[cid].n_mal == QM_FOREIGN;
                if (!mine && (foreign || peer_mode == QM_PEER_AWAY)) continue;
            }
            uint32_t c = 0;
            if (mine) {
                const int32_t qs = ld_stream(qs_in + t), qe = ld_stream(qe_in + t);
                c = mixed_answer(tab[cid], qs, qe);
                add_total(cid, c);
            }
            st_stream(counts + t, (CountT)c);
        }
    }
    if (s_tot) {
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < n_contigs; k += QM_THREADS)
            if (s_tot[k]) atomicAdd(totals + k, s_tot[k]);
    }
}

// ---- has_overlaps: tests ONLY the last candidate (hpp:865-871, quirk Q1) ----------------
__global__ void __launch_bounds__(QK_THREADS)
qk_any_kernel(IndexView ix, const int32_t* __restrict__ qs_in, const int32_t* __restrict__ qe_in,
              uint32_t nq, uint8_t* __restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * QK_THREADS + threadIdx.x;
    if (t >= nq) return;
    uint8_t r = 0;
    if (ix.n) {
        // ub(qe) = #{starts <= qe} - 1: one rank-cell sector where the cells cover every stored interval, else the halving search
        const int32_t qe = qe_in[t];
        const uint32_t i = (ix.cells_s.fmt != 0u && ix.n_mal == 0u ? cells_rank_lt(ix.cells_s, ix.starts, (int64_t)qe + 1)
                                                                    : count_le(ix.starts, ix.n, qe)) - 1u;
        r = (i != NONE32 && qs_in[t] <= ld_nc(ix.ends + i)) ? 1 : 0;
    }
    out[t] = r;
}

// ---- upperBound for the C ABI cursor (c_superintervals.h:537-568) ------------------------

// ---- coverage: hit count + sum of clipped lengths (hpp:979-1006, c.h:758-792) -----------
// int32 arithmetic that wraps exactly like the reference's `S` accumulator.
__global__ void __launch_bounds__(QK_THREADS)
qk_coverage_kernel(IndexView ix, const int32_t* __restrict__ qs_in, const int32_t* __restrict__ qe_in,
                   uint32_t nq, uint32_t* __restrict__ counts, int32_t* __restrict__ cov) {
    const uint64_t t = (uint64_t)blockIdx.x * QK_THREADS + threadIdx.x;
    if (t >= nq) return;
    const int32_t qs = qs_in[t], qe = qe_in[t];
    uint32_t c = 0, sum = 0;
    uint32_t i = ix.n ? count_le(ix.starts, ix.n, qe) - 1u : NONE32;
    while (i != NONE32) {
        const uint32_t base = i & ~3u;
        const int4 e = ld_nc4(ix.ends + base);
        const int4 s = ld_nc4(ix.starts + base);
        const uint32_t k = i - base;
#define SIB_COV(EE, SS, OK)                                                        \
        if ((OK) && (EE) >= qs) {                                                  \
            ++c;                                                                   \
            sum += (uint32_t)min((EE), qe) - (uint32_t)max((SS), qs);              \
        }
        SIB_COV(e.x, s.x, true)
        SIB_COV(e.y, s.y, k >= 1u)
        SIB_COV(e.z, s.z, k >= 2u)
        SIB_COV(e.w, s.w, k >= 3u)
#undef SIB_COV
        i = (e.x >= qs) ? base - 1u : ld_nc(ix.branch + base);
    }
    counts[t] = c;
    cov[t] = (int32_t)sum;
}

// ---- fill: CSR values in the reference's descending order -------------------------------
enum FillMode : int { FILL_VALUES = 0, FILL_IDXS = 1, FILL_KEYS = 2, FILL_ITEMS = 3 };

template <int MODE> struct FillOut;
template <> struct FillOut<FILL_VALUES> { using T = int32_t; };
template <> struct FillOut<FILL_IDXS> { using T = uint32_t; };
template <> struct FillOut<FILL_KEYS> { using T = int2; };
struct __align__(4) Item3 { int32_t start, end, data; };
template <> struct FillOut<FILL_ITEMS> { using T = Item3; };

template <int MODE>
__device__ __forceinline__ void emit(const IndexView& ix, typename FillOut<MODE>::T* __restrict__ out,
                                     uint64_t pos, uint32_t j, int32_t end_j) {
    if (MODE == FILL_VALUES) {
        reinterpret_cast<int32_t*>(out)[pos] = ld_nc(ix.values + j);
    } else if (MODE == FILL_IDXS) {
        reinterpret_cast<uint32_t*>(out)[pos] = j;
    } else if (MODE == FILL_KEYS) {
        reinterpret_cast<int2*>(out)[pos] = make_int2(ld_nc(ix.starts + j), end_j);
    } else {
        Item3 it;
        it.start = ld_nc(ix.starts + j);
        it.end = end_j;
        it.data = ld_nc(ix.values + j);
        reinterpret_cast<Item3*>(out)[pos] = it;
    }
}

constexpr uint32_t QK_FILL_LANE_MAX = 48;   // queries with more hits go straight to the warp phase

template <int MODE>
__global__ void __launch_bounds__(QK_THREADS)
qk_fill_kernel(IndexView ix, QueryRecords rec, uint32_t nq, const uint64_t* __restrict__ offsets,
               typename FillOut<MODE>::T* __restrict__ out) {
    const uint64_t t64 = (uint64_t)blockIdx.x * QK_THREADS + threadIdx.x;
    const bool live = t64 < nq;
    const uint32_t t = (uint32_t)t64;
    const uint32_t lane = lane_id();

    uint32_t q = t;
    int32_t qs = 0, qe = 0;
    uint64_t o = 0, o_end = 0;
    if (live) {
        qs = ld_stream(rec.qs + t);
        qe = ld_stream(rec.qe + t);
        if (rec.idx) q = ld_stream(rec.idx + t);
        o = ld_stream(offsets + q);
        o_end = ld_stream(offsets + q + 1);
    }
    const bool any_hits = live && o_end > o;
    uint32_t i = NONE32;
    if (any_hits) i = count_le(ix.starts, ix.n, qe) - 1u;   // no hits -> nothing to walk

    // ---- phase A: short result lists, lane-per-query
    const bool small = any_hits && (o_end - o) <= QK_FILL_LANE_MAX;
    if (small) {
        while (i != NONE32) {
            const uint32_t base = i & ~3u;
            const int4 e = ld_nc4(ix.ends + base);
            const uint32_t k = i - base;
            if (k >= 3u && e.w >= qs) emit<MODE>(ix, out, o++, base + 3u, e.w);
            if (k >= 2u && e.z >= qs) emit<MODE>(ix, out, o++, base + 2u, e.z);
            if (k >= 1u && e.y >= qs) emit<MODE>(ix, out, o++, base + 1u, e.y);
            if (e.x >= qs) emit<MODE>(ix, out, o++, base, e.x);
            if (o == o_end) break;   // every hit written: the rest of the walk only skips
            i = (e.x >= qs) ? base - 1u : ld_nc(ix.branch + base);
        }
        i = NONE32;
    }
    // ---- phase B: long result lists, whole warp per query, coalesced stores
    uint32_t pending = __ballot_sync(FULL_MASK, any_hits && !small);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        uint32_t bi = __shfl_sync(FULL_MASK, i, src);
        const int32_t bqs = __shfl_sync(FULL_MASK, qs, src);
        uint64_t bo = __shfl_sync(FULL_MASK, o, src);
        const uint64_t bo_end = __shfl_sync(FULL_MASK, o_end, src);
        while (bi != NONE32 && bo < bo_end) {
            // lane l looks at interval bi - l: 32 consecutive intervals, descending
            const bool inb = lane <= bi;
            const uint32_t j = bi - lane;
            const int32_t e = inb ? ld_nc(ix.ends + j) : INT_MIN;
            const bool hit = inb && e >= bqs;
            const uint32_t hm = __ballot_sync(FULL_MASK, hit);
            if (hit) emit<MODE>(ix, out, bo + __popc(hm & lanemask_lt()), j, e);
            bo += __popc(hm);
            if (bi < 32u) break;                     // the block reached interval 0
            const uint32_t low = bi - 31u;           // lowest interval of the block (lane 31)
            bi = (hm >> 31) ? low - 1u : ld_nc(ix.branch + low);
        }
    }
}

// ---- ONE query, one warp: the reference's single-query calls (hpp:551-579, 879-971; c.h:575-727) ----------
// The C ABI is one query per call, and every reference driver loops over it (test/bench.cpp:219-222,240-242). A
// call is then latency, not throughput: one launch of this kernel, whose query arrives as kernel parameters and
// whose result goes straight into a mapped pinned mailbox on the host (no copy in, no copy out, one stream
// synchronise). The warp finds upper_bound(qe) by a 32-ary search (5 dependent loads for 2^25 intervals instead
// of 25) and walks the branch array 32 intervals at a time, hits emitted in the reference's descending order.
// mailbox[0] = hits found (all of them, also those past `cap` that were not written: the host then takes the batch path).
__device__ __forceinline__ uint32_t warp_count_le(const int32_t* __restrict__ starts, uint32_t n, int32_t v, uint32_t lane) {
    uint32_t lo = 0, len = n;                       // every index < lo holds a start <= v; none at or past lo + len does
    while (len) {
        const uint32_t width = (len + 31u) >> 5;
        const uint32_t p = lo + (lane + 1u) * width - 1u;
        const bool le = p < lo + len && ld_nc(starts + p) <= v;
        const uint32_t c = __popc(__ballot_sync(FULL_MASK, le));
        const uint32_t nlo = lo + c * width;
        const uint32_t rest = lo + len - nlo;       // c == 32 leaves nothing or the tail past the last probe
        len = min(width - 1u, rest);
        if (c == 32u) len = rest;
        lo = nlo;
    }
    return lo;
}

// Every single-query kernel ends by publishing the call's sequence number in the mailbox (after a system-scope fence):
// the host spins on that word in pinned memory instead of paying a stream synchronise.
__device__ __forceinline__ void single_done(uint32_t* __restrict__ done, uint32_t seq) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(done) = seq;
}

// op 0: upper_bound(a) (hpp:501-516) -> out32; op 1: has_overlaps(a, b), the last candidate only (hpp:865-871, Q1) -> out32;
// op 2: count(a, b) -> out64: the closed form from the rank cells where the index carries them and a <= b (two sector
// loads by lane 0), else the reference's walk, 32 intervals per step.
__device__ __forceinline__ void single_scalar_body(const IndexView& ix, int op, int32_t a, int32_t b, uint32_t* __restrict__ out32,
                                                   unsigned long long* __restrict__ out64, uint32_t lane) {
    if (op == 0) {
        const uint32_t c = warp_count_le(ix.starts, ix.n, a, lane);
        if (lane == 0) *out32 = c - 1u;                                            // 0 - 1 wraps to NONE32
    } else if (op == 1) {
        const uint32_t i = warp_count_le(ix.starts, ix.n, b, lane) - 1u;
        if (lane == 0) *out32 = (i != NONE32 && a <= ld_nc(ix.ends + i)) ? 1u : 0u;
    } else if (ix.cells_s.fmt && ix.cells_e.fmt && a <= b) {
        if (lane == 0) {
            uint32_t c = cells_rank_lt(ix.cells_s, ix.rstarts, (int64_t)b + 1) - cells_rank_lt(ix.cells_e, ix.eall, (int64_t)a);
            for (uint32_t k = 0; k < ix.n_mal; ++k) c += (ix.mal_s[k] <= b && ix.mal_e[k] >= a) ? 1u : 0u;
            *out64 = c;
        }
    } else {
        uint32_t bi = warp_count_le(ix.starts, ix.n, b, lane) - 1u, bo = 0;
        while (bi != NONE32) {
            const bool inb = lane <= bi;
            const bool hit = inb && ld_nc(ix.ends + (bi - lane)) >= a;
            const uint32_t hm = __ballot_sync(FULL_MASK, hit);
            bo += __popc(hm);
            if (bi < 32u) break;
            const uint32_t low = bi - 31u;
            bi = (hm >> 31) ? low - 1u : ld_nc(ix.branch + low);
        }
        if (lane == 0) *out64 = bo;
    }
}

__global__ void __launch_bounds__(32)
qk_single_scalar_kernel(IndexView ix, int op, int32_t a, int32_t b, uint32_t* __restrict__ out32,
                        unsigned long long* __restrict__ out64, uint32_t* __restrict__ done, uint32_t seq) {
    const uint32_t lane = lane_id();
    single_scalar_body(ix, op, a, b, out32, out64, lane);
    if (lane == 0) single_done(done, seq);
}

// One query, one warp, 128 intervals per step (32 lanes x one 128-bit load of ends: the step of walk_warp128): a call is
// a chain of dependent loads, so the fewer steps the better -- C2's lists (138 hits spread over ~400 positions, ~100
// jumps of the element walk) take 4-6 steps instead of 25. Hits leave in the reference's descending order: a lane's
// rank is the number of hits in the lanes above it (suffix sum by shuffles) plus those among its own higher slots.
template <int MODE>
__device__ __forceinline__ void single_search_body(const IndexView& ix, int32_t qs, int32_t qe, uint32_t cap,
                                                   unsigned long long* __restrict__ found, typename FillOut<MODE>::T* __restrict__ out,
                                                   uint32_t lane) {
    uint32_t bi = warp_count_le(ix.starts, ix.n, qe, lane) - 1u;      // 0 - 1 wraps to NONE32
    uint32_t bo = 0;
    while (bi != NONE32) {
        const uint32_t base = bi & ~127u;
        const uint32_t j = base + lane * 4u;
        const bool v0 = j <= bi, v1 = j + 1u <= bi, v2 = j + 2u <= bi, v3 = j + 3u <= bi;
        int4 e = make_int4(INT_MIN, INT_MIN, INT_MIN, INT_MIN);
        if (v0) e = ld_nc4(ix.ends + j);                               // ends is padded to a multiple of 128 entries
        const bool h0 = v0 && e.x >= qs, h1 = v1 && e.y >= qs, h2 = v2 && e.z >= qs, h3 = v3 && e.w >= qs;
        const uint32_t mine = (h0 ? 1u : 0u) + (h1 ? 1u : 0u) + (h2 ? 1u : 0u) + (h3 ? 1u : 0u);
        uint32_t incl = mine;                                          // hits in this lane and every lane above it
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_down_sync(FULL_MASK, incl, off);
            if (lane + (uint32_t)off < 32u) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL_MASK, incl, 0);
        uint32_t p = bo + incl - mine;
        if (h3) { if (p < cap) emit<MODE>(ix, out, p, j + 3u, e.w); ++p; }
        if (h2) { if (p < cap) emit<MODE>(ix, out, p, j + 2u, e.z); ++p; }
        if (h1) { if (p < cap) emit<MODE>(ix, out, p, j + 1u, e.y); ++p; }
        if (h0) { if (p < cap) emit<MODE>(ix, out, p, j, e.x); }
        bo += total;
        if (__shfl_sync(FULL_MASK, h0 ? 1u : 0u, 0)) {                 // the block's lowest interval hits: step below the block
            bi = base - 1u;                                            // base == 0 wraps to NONE32
        } else {                                                       // smallest branch target over every miss of the block
            uint32_t key = 0xFFFFFFFFu;
            if (v0 && !(h0 && h1 && h2 && h3)) {
                const uint4 br = __ldg(reinterpret_cast<const uint4*>(ix.branch + j));
                if (!h0) key = jump_key(br.x);
                if (v1 && !h1) key = min(key, jump_key(br.y));
                if (v2 && !h2) key = min(key, jump_key(br.z));
                if (v3 && !h3) key = min(key, jump_key(br.w));
            }
            bi = __reduce_min_sync(FULL_MASK, key) - 1u;
        }
    }
    __syncwarp();
    if (lane == 0) *found = bo;
}

template <int MODE>
__global__ void __launch_bounds__(32)
qk_single_search_kernel(IndexView ix, int32_t qs, int32_t qe, uint32_t cap, unsigned long long* __restrict__ found,
                        typename FillOut<MODE>::T* __restrict__ out, uint32_t* __restrict__ done, uint32_t seq) {
    const uint32_t lane = lane_id();
    single_search_body<MODE>(ix, qs, qe, cap, found, out, lane);
    if (lane == 0) single_done(done, seq);
}

// ---- resident single-query kernel (SI_OPT_RESIDENT_QUERIES) --------------------------------------------------
// A loop of single-query C calls pays a kernel launch per call (~11 us). With this option ONE warp stays resident and
// polls the request block of the mapped pinned mailbox: the host writes the query and then its sequence number, the warp
// answers with the same code as the per-call kernels and publishes the sequence number after the answer. The kernel
// leaves by itself after `idle_ns` without a request and at the latest `life_ns` after its launch (so that a device-wide
// synchronise elsewhere in the process -- cudaFree, cudaDeviceSynchronize -- is never held up for longer), or when the host
// raises `stop`; the host relaunches it on demand. Before leaving it lowers `alive` and looks once more for a request.
__device__ __forceinline__ uint32_t ld_sys_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint4 ld_sys_v4(const void* p) {
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

__global__ void __launch_bounds__(32)
qk_single_server_kernel(IndexView ix, SingleReq* __restrict__ req, uint32_t* __restrict__ out32, unsigned long long* __restrict__ out64,
                        void* __restrict__ out, uint32_t* __restrict__ done, uint32_t last, unsigned long long idle_ns,
                        unsigned long long life_ns) {
    const uint32_t lane = lane_id();
    const unsigned long long t_start = global_ns();
    unsigned long long t_idle = t_start;
    while (true) {
        uint4 rec = make_uint4(0, 0, 0, 0);
        if (lane == 0) rec = ld_sys_v4(req);                              // { seq, a, b, op << 28 | cap }: one PCIe read per poll
        const uint32_t r = __shfl_sync(FULL_MASK, rec.x, 0);
        if (r != last) {
            const int32_t a = (int32_t)__shfl_sync(FULL_MASK, rec.y, 0), b = (int32_t)__shfl_sync(FULL_MASK, rec.z, 0);
            const uint32_t opcap = __shfl_sync(FULL_MASK, rec.w, 0);
            const int op = (int)(opcap >> 28);
            const uint32_t cap = opcap & 0x0FFFFFFFu;
            if (op == 15) {                                               // the host asks the kernel to leave
                if (lane == 0) { *reinterpret_cast<volatile uint32_t*>(&req->alive) = 0u; __threadfence_system(); }
                return;
            }
            switch (op) {
                case 3 + FILL_VALUES: single_search_body<FILL_VALUES>(ix, a, b, cap, out64, reinterpret_cast<int32_t*>(out), lane); break;
                case 3 + FILL_IDXS: single_search_body<FILL_IDXS>(ix, a, b, cap, out64, reinterpret_cast<uint32_t*>(out), lane); break;
                case 3 + FILL_KEYS: single_search_body<FILL_KEYS>(ix, a, b, cap, out64, reinterpret_cast<int2*>(out), lane); break;
                case 3 + FILL_ITEMS: single_search_body<FILL_ITEMS>(ix, a, b, cap, out64, reinterpret_cast<Item3*>(out), lane); break;
                default: single_scalar_body(ix, op, a, b, out32, out64, lane); break;
            }
            __syncwarp();
            if (lane == 0) single_done(done, r);
            last = r;
            t_idle = global_ns();
            continue;
        }
        const unsigned long long now = global_ns();
        const bool expired = now - t_start > life_ns;
        if (expired || now - t_idle > idle_ns) {
            uint32_t again = 0;
            if (lane == 0) {
                *reinterpret_cast<volatile uint32_t*>(&req->alive) = 0u;
                __threadfence_system();
                again = (!expired && ld_sys_u32(&req->seq) != last) ? 1u : 0u;   // a request slipped in: stay
                if (again) *reinterpret_cast<volatile uint32_t*>(&req->alive) = 1u;
            }
            again = __shfl_sync(FULL_MASK, again, 0);
            if (!again) return;
            t_idle = now;
        }
    }
}

// ---- stab lists: one branch-array walk per checkpoint, at build time ------------------------
// FILL = false counts |L(b)|, FILL = true writes the entries at off[b]. The walk is the
// reference's (hpp:551-579) with the threshold x = starts[(b << kshift) - 1] + 1, started at the
// checkpoint's last position; a miss jumps to the smallest branch target over the block's misses.
template <bool FILL, bool REC16>
__global__ void __launch_bounds__(QK_THREADS)
qk_stab_lists_kernel(IndexView ix, uint32_t kshift, uint32_t nlists, uint32_t* __restrict__ counts,
                     const uint64_t* __restrict__ off, void* __restrict__ ent_out, uint4* __restrict__ hdr, int2* __restrict__ entv_out) {
    const uint64_t t64 = (uint64_t)blockIdx.x * QK_THREADS + threadIdx.x;
    if (t64 >= nlists) return;
    const uint32_t b = (uint32_t)t64;
    uint32_t c = 0;
    const uint64_t o0 = FILL ? off[b] : 0;
    uint64_t o = o0;
    if (b > 0) {
        uint32_t i = (b << kshift) - 1u;
        const int32_t sv = ld_nc(ix.starts + i);
        if (sv != INT_MAX) {                 // nothing ends beyond INT_MAX
            const int32_t x = sv + 1;
            while (i != NONE32) {
                const uint32_t base = i & ~3u;
                const int4 e = ld_nc4(ix.ends + base);
                const uint32_t k = i - base;
                const bool hx = e.x >= x, hy = k >= 1u && e.y >= x, hz = k >= 2u && e.z >= x, hw = k >= 3u && e.w >= x;
                if (FILL) {
#define SIB_PUT(J, E)                                                                                     \
                    do {                                                                                  \
                        if (REC16) reinterpret_cast<int4*>(ent_out)[o++] = make_int4((int)(J), (E), ld_nc(ix.values + (J)), 0); \
                        else {                                                                            \
                            if (entv_out) entv_out[o] = make_int2(ld_nc(ix.values + (J)), (E));           \
                            reinterpret_cast<int2*>(ent_out)[o++] = make_int2((int)(J), (E));             \
                        }                                                                                 \
                    } while (0)
                    if (hw) SIB_PUT(base + 3u, e.w);
                    if (hz) SIB_PUT(base + 2u, e.z);
                    if (hy) SIB_PUT(base + 1u, e.y);
                    if (hx) SIB_PUT(base, e.x);
#undef SIB_PUT
                } else {
                    c += (hx ? 1u : 0u) + (hy ? 1u : 0u) + (hz ? 1u : 0u) + (hw ? 1u : 0u);
                }
                if (hx) { i = base - 1u; continue; }
                const uint4 br = __ldg(reinterpret_cast<const uint4*>(ix.branch + base));
                uint32_t key = jump_key(br.x);
                if (k >= 1u && !hy) key = min(key, jump_key(br.y));
                if (k >= 2u && !hz) key = min(key, jump_key(br.z));
                if (k >= 3u && !hw) key = min(key, jump_key(br.w));
                i = key - 1u;
            }
        }
    }
    if (!FILL) {
        counts[b] = (c + 3u) & ~3u;          // padded: every list starts on a 32-byte boundary
    } else {
        const uint32_t len = (uint32_t)(o - o0);
        for (uint64_t pad = o; pad < o0 + ((len + 3u) & ~3u); ++pad) {
            if (REC16) reinterpret_cast<int4*>(ent_out)[pad] = make_int4(0, INT_MIN, 0, 0);
            else {
                reinterpret_cast<int2*>(ent_out)[pad] = make_int2(0, INT_MIN);
                if (entv_out) entv_out[pad] = make_int2(0, INT_MIN);
            }
        }
        hdr[b] = make_uint4((uint32_t)o0, (uint32_t)(o0 >> 32), len, 0u);
    }
}

// totals[k] = sum of counts[b] over the checkpoints a spacing of 2^(QK_STAB_SHIFT0 + k) keeps (b % 2^k == 0)
constexpr uint32_t QK_STAB_SHIFT0 = 3;   // finest checkpoint spacing: 8 positions (two 128-bit loads of ends at most)
constexpr int QK_STAB_SPACINGS = 8;      // coarsest: 1024
__global__ void __launch_bounds__(QK_THREADS)
qk_stab_totals_kernel(const uint32_t* __restrict__ counts, uint32_t nlists, unsigned long long* __restrict__ totals) {
    unsigned long long acc[QK_STAB_SPACINGS];
#pragma unroll
    for (int k = 0; k < QK_STAB_SPACINGS; ++k) acc[k] = 0;
    const uint64_t stride = (uint64_t)gridDim.x * QK_THREADS;
    for (uint64_t b = (uint64_t)blockIdx.x * QK_THREADS + threadIdx.x; b < nlists; b += stride) {
        const uint32_t c = counts[b];
#pragma unroll
        for (int k = 0; k < QK_STAB_SPACINGS; ++k)
            if ((b & ((1u << k) - 1u)) == 0) acc[k] += c;
    }
#pragma unroll
    for (int k = 0; k < QK_STAB_SPACINGS; ++k) {
        unsigned long long v = acc[k];
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
        if (lane_id() == 0 && v) atomicAdd(totals + k, v);
    }
}

// ---- fill by run + stab (well-formed index with rank cells) --------------------------------
// On an index whose intervals all have start <= end, the walk's hit list
//     { j <= ub(qe) : ends[j] >= qs }   in descending j
// splits at  top = #{ starts < qs }  (clamped to ub(qe) + 1):
//   run   j in [top, ub(qe)]: start >= qs, so end >= start >= qs -- EVERY one is a hit (the run the
//         reference's search_values_large locates by search, hpp:588-615). No ends are tested:
//         the list's head is a reversed copy of values[top .. ub(qe)].
//   stab  j < top with ends[j] >= qs: the intervals that begin before the query and reach into
//         it. With stab lists (StabLists above) the candidates of a query are the positions of
//         its partial checkpoint block [top & ~(2^kshift - 1), top), descending, followed by the
//         checkpoint's list; the candidates of a warp's 32 queries are pooled like the runs, every
//         lane tests one candidate per step (neighbouring lanes read neighbouring ends / list
//         records), hits are ranked inside their query's lane segment with ballot/popc and stored
//         next to each other -- independent coalesced loads, no pointer chase.
//         Without them (an index nested too deeply for the lists' memory budget): the
//         reference's own branch-array walk from top - 1. Either way it stops as soon as the CSR
//         slot is full (the offsets say how many hits exist).
// Both ranks come from the rank cells over starts (one sector each, no binary search), so the
// batch needs no locality and is filled in the caller's order, unpartitioned.
// The runs of a warp's 32 queries are pooled and copied by all lanes together (lane k of the
// pool finds its query by a 5-step search over the warp's prefix sums in shared memory): the
// copy is load-balanced whatever the distribution of run lengths. Runs longer than
// QF_LONG_RUN are streamed by the whole warp, one query at a time. Inverted queries
// (qs > qe, quirk Q6) get top = ub(qe) + 1: empty run, plain walk.
constexpr int QF_THREADS = 256;
constexpr int QF_WARPS = QF_THREADS / 32;
constexpr uint32_t QF_LONG_RUN = 1024;
constexpr uint32_t QF_POOL_MIN = 32 * 24;   // stab hits owed by a warp from which its candidates are pooled

// one element of a run: position j is a hit for certain, only the payload is read
template <int MODE>
__device__ __forceinline__ void emit_run(const IndexView& ix, typename FillOut<MODE>::T* __restrict__ out,
                                         uint64_t pos, uint32_t j) {
    if (MODE == FILL_VALUES) {
        __stcs(reinterpret_cast<int32_t*>(out) + pos, ld_nc(ix.values + j));
    } else if (MODE == FILL_IDXS) {
        __stcs(reinterpret_cast<uint32_t*>(out) + pos, j);
    } else {
        emit<MODE>(ix, out, pos, j, ld_nc(ix.ends + j));
    }
}

template <int MODE>
__global__ void __launch_bounds__(QF_THREADS)
qk_fill_runs_kernel(IndexView ix, QueryRecords rec, uint32_t nq, const uint64_t* __restrict__ offsets,
                    typename FillOut<MODE>::T* __restrict__ out) {
    // 4 KB per CTA: six resident CTAs stay inside the smallest shared-memory carve-out, the rest of the
    // SM's 256 KB remains L1 for the scattered index reads (9 KB per CTA cost 20 % of the kernel)
    __shared__ uint32_t s_pre[QF_WARPS][32];   // inclusive prefix sums of the warp's pooled run (then candidate) counts
    __shared__ uint32_t s_top[QF_WARPS][32];   // highest position of each run (= ub(qe)); then hits written per query
    __shared__ uint64_t s_out[QF_WARPS][32];   // where each run's first element goes
    const uint64_t t64 = (uint64_t)blockIdx.x * QF_THREADS + threadIdx.x;
    const bool live = t64 < nq;
    const uint32_t t = (uint32_t)t64;
    const uint32_t lane = lane_id(), w = threadIdx.x >> 5;

    uint32_t q = t;
    int32_t qs = 0, qe = 0;
    uint64_t o = 0, o_end = 0;
    if (live) {
        qs = ld_stream(rec.qs + t);
        qe = ld_stream(rec.qe + t);
        if (rec.idx) q = ld_stream(rec.idx + t);
        o = ld_stream(offsets + q);
        o_end = ld_stream(offsets + q + 1);
    }
    const bool any_hits = live && o_end > o;
    uint32_t lim = 0, top = 0;   // lim = #{starts <= qe}, top = min(#{starts < qs}, lim)
    if (any_hits) {
        uint32_t c1, f1, c2, f2;
        cell_of(ix.cells_s, (int64_t)qe + 1, c1, f1);
        cell_of(ix.cells_s, (int64_t)qs, c2, f2);
        const CellRec r1 = ld_cell(ix.cells_s.rec + 2 * (size_t)c1);   // both sectors in flight
        const CellRec r2 = ld_cell(ix.cells_s.rec + 2 * (size_t)c2);
        lim = cell_rank(ix.cells_s, ix.starts, r1, c1, f1, (int64_t)qe + 1);
        top = min(cell_rank(ix.cells_s, ix.starts, r2, c2, f2, (int64_t)qs), lim);
    }
    const uint32_t run = lim - top;

    // ---- run: long ones by the whole warp, one query at a time
    uint32_t longm = __ballot_sync(FULL_MASK, run > QF_LONG_RUN);
    while (longm) {
        const int src = __ffs(longm) - 1;
        longm &= longm - 1;
        const uint32_t blen = __shfl_sync(FULL_MASK, run, src);
        const uint32_t bhi = __shfl_sync(FULL_MASK, lim, src) - 1u;
        const uint64_t bo = __shfl_sync(FULL_MASK, o, src);
#pragma unroll 4
        for (uint32_t k = lane; k < blen; k += 32u) emit_run<MODE>(ix, out, bo + k, bhi - k);
    }
    // ---- run: the rest pooled over the warp
    const uint32_t pooled = run > QF_LONG_RUN ? 0u : run;
    uint32_t incl = pooled;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(FULL_MASK, incl, d);
        if ((int)lane >= d) incl += x;
    }
    const uint32_t total = __shfl_sync(FULL_MASK, incl, 31);
    if (total) {
        s_pre[w][lane] = incl;
        s_top[w][lane] = lim - 1u;
        s_out[w][lane] = o;
        __syncwarp();
        for (uint32_t k = lane; k < total; k += 32u) {
            uint32_t own = 0;                       // number of queries whose pooled runs end at or before k
#pragma unroll
            for (uint32_t step = 16; step; step >>= 1) own += (s_pre[w][own + step - 1u] <= k) ? step : 0u;
            const uint32_t within = k - (own ? s_pre[w][own - 1u] : 0u);
            emit_run<MODE>(ix, out, s_out[w][own] + within, s_top[w][own] - within);
        }
        __syncwarp();
    }
    o += run;

    // ---- stab: below top, until the slot is full
    const bool more = any_hits && o < o_end;
    if (ix.stab.ent) {
        // candidates of this query: plen positions top-1 .. cb (descending), then llen list records
        uint32_t plen = 0, llen = 0;
        uint64_t l0 = 0;
        if (more && top) {
            const uint32_t b = top >> ix.stab.kshift;
            plen = top - (b << ix.stab.kshift);
            const uint4 h = __ldg(ix.stab.hdr + b);                // first record and length in one load
            l0 = (uint64_t)h.x | ((uint64_t)h.y << 32);
            llen = h.z;
        }
        const uint32_t topm1 = top - 1u;
        const bool rec16 = ix.stab.rec16 != 0;   // list records carry their value
        const bool vlist = MODE == FILL_VALUES && !rec16 && ix.stab.entv != nullptr;   // 8-byte (value, end) lists: no payload gather
        // one candidate: c < plen -> position topm1 - c read from ends[], else list record c - plen
#define SIB_STAB_CAND(C, PLEN, TOPM1, L0, POS, END, VAL, FROM_LIST)                                   \
        if ((C) < (PLEN)) { POS = (TOPM1) - (C); END = ld_nc(ix.ends + POS); VAL = 0; FROM_LIST = false; } \
        else if (vlist) { const int2 r_ = __ldg(reinterpret_cast<const int2*>(ix.stab.entv) + ((L0) + ((C) - (PLEN)))); POS = 0; END = r_.y; VAL = r_.x; FROM_LIST = true; } \
        else { const int4 r_ = ld_stab_record(ix.stab, (L0) + ((C) - (PLEN))); POS = (uint32_t)r_.x; END = r_.y; VAL = r_.z; FROM_LIST = rec16; }
#define SIB_STAB_EMIT(AT, POS, END, VAL, FROM_LIST)                                                   \
        if (MODE == FILL_VALUES) reinterpret_cast<int32_t*>(out)[AT] = (FROM_LIST) ? (VAL) : ld_nc(ix.values + (POS)); /* FROM_LIST: VAL is valid */ \
        else emit<MODE>(ix, out, (AT), (POS), (END));
        // few stab hits in the whole warp (sparse data): pooling costs more instructions than the
        // scattered accesses it saves, so each lane sweeps its own candidates. Decided from the hits
        // still owed -- known from the offsets -- so that nothing waits for the list bounds here.
        const uint32_t owed = more ? (uint32_t)min(o_end - o, (uint64_t)0x3FFFFFFu) : 0u;
        const uint32_t wsum = __reduce_add_sync(FULL_MASK, owed);
        if (wsum < QF_POOL_MIN) {
            if (more && top) {
                if (plen) {
                    uint32_t base = topm1 & ~3u;
                    uint32_t k = topm1 - base;
                    const uint32_t cb = top - plen;
                    while (true) {
                        const int4 e = ld_nc4(ix.ends + base);
                        const bool hw = k >= 3u && e.w >= qs, hz = k >= 2u && e.z >= qs, hy = k >= 1u && e.y >= qs, hx = e.x >= qs;
                        if (MODE == FILL_VALUES) {
                            // the four payloads are neighbours: one 128-bit load instead of a gather per hit
                            if (hw || hz || hy || hx) {
                                const int4 pv = ld_nc4(ix.values + base);   // values[] is padded like ends[]
                                int32_t* o32 = reinterpret_cast<int32_t*>(out);
                                if (hw) o32[o++] = pv.w;
                                if (hz) o32[o++] = pv.z;
                                if (hy) o32[o++] = pv.y;
                                if (hx) o32[o++] = pv.x;
                            }
                        } else {
                            if (hw) emit<MODE>(ix, out, o++, base + 3u, e.w);
                            if (hz) emit<MODE>(ix, out, o++, base + 2u, e.z);
                            if (hy) emit<MODE>(ix, out, o++, base + 1u, e.y);
                            if (hx) emit<MODE>(ix, out, o++, base, e.x);
                        }
                        if (base == cb || o == o_end) break;
                        base -= 4u;
                        k = 3u;
                    }
                }
                const uint64_t l1 = l0 + llen;
                if (!rec16) {
                    const int2* __restrict__ ent2 = reinterpret_cast<const int2*>(vlist ? ix.stab.entv : ix.stab.ent);
                    for (uint64_t p = l0; p < l1 && o < o_end; p += 4) {   // four records = one 32-byte sector per step
                        const CellRec r = ld_cell(reinterpret_cast<const uint4*>(ent2 + p));   // lists are padded to 4
                        int2 v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = make_int2((int)r.w[2 * u], (int)r.w[2 * u + 1]);
                        bool h[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) h[u] = v[u].y >= qs;   // pad records end at INT_MIN
                        if (MODE == FILL_VALUES) {
                            int32_t pv[4];                       // the hits' payload gathers leave together
#pragma unroll
                            for (int u = 0; u < 4; ++u) pv[u] = vlist ? v[u].x : (h[u] ? ld_nc(ix.values + (uint32_t)v[u].x) : 0);   // value lists: no gather
                            int32_t* o32 = reinterpret_cast<int32_t*>(out);
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (h[u]) o32[o++] = pv[u];
                        } else {
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (h[u]) emit<MODE>(ix, out, o++, (uint32_t)v[u].x, v[u].y);
                        }
                    }
                } else {
                    const int4* __restrict__ ent4 = reinterpret_cast<const int4*>(ix.stab.ent);
                    for (uint64_t p = l0; p < l1 && o < o_end; p += 2) {
                        const int4 r0 = __ldg(ent4 + p);
                        const int4 r1 = (p + 1 < l1) ? __ldg(ent4 + p + 1) : make_int4(0, INT_MIN, 0, 0);
                        if (r0.y >= qs) { SIB_STAB_EMIT(o, (uint32_t)r0.x, r0.y, r0.z, true) ++o; }
                        if (p + 1 < l1 && r1.y >= qs) { SIB_STAB_EMIT(o, (uint32_t)r1.x, r1.y, r1.z, true) ++o; }
                    }
                }
            }
            return;
        }
        const uint64_t clen64 = (uint64_t)plen + llen;
        // queries with very many candidates: the whole warp sweeps one query at a time
        uint32_t longm = __ballot_sync(FULL_MASK, clen64 > QF_LONG_RUN);
        while (longm) {
            const int src = __ffs(longm) - 1;
            longm &= longm - 1;
            const uint32_t bpl = __shfl_sync(FULL_MASK, plen, src), bll = __shfl_sync(FULL_MASK, llen, src);
            const uint32_t btm = __shfl_sync(FULL_MASK, topm1, src);
            const uint64_t bl0 = __shfl_sync(FULL_MASK, l0, src);
            const int32_t bqs = __shfl_sync(FULL_MASK, qs, src);
            uint64_t bo = __shfl_sync(FULL_MASK, o, src);
            const uint64_t bo_end = __shfl_sync(FULL_MASK, o_end, src);
            const uint64_t bn = (uint64_t)bpl + bll;
            for (uint64_t cb = 0; cb < bn && bo < bo_end; cb += 32) {
                const uint64_t c = cb + lane;
                uint32_t pos = 0; int32_t end = INT_MIN, val = 0; bool fl = false;
                if (c < bn) {
                    if (c < bpl) { pos = btm - (uint32_t)c; end = ld_nc(ix.ends + pos); }
                    else if (vlist) { const int2 r = __ldg(reinterpret_cast<const int2*>(ix.stab.entv) + (bl0 + (c - bpl))); end = r.y; val = r.x; fl = true; }
                    else { const int4 r = ld_stab_record(ix.stab, bl0 + (c - bpl)); pos = (uint32_t)r.x; end = r.y; val = r.z; fl = rec16; }
                }
                const bool hit = c < bn && end >= bqs;
                const uint32_t hm = __ballot_sync(FULL_MASK, hit);
                if (hit) { SIB_STAB_EMIT(bo + __popc(hm & lanemask_lt()), pos, end, val, fl) }
                bo += __popc(hm);
            }
        }
        // everything else pooled over the warp
        const uint32_t pooled_c = clen64 > QF_LONG_RUN ? 0u : (uint32_t)clen64;
        uint32_t cincl = pooled_c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(FULL_MASK, cincl, d);
            if ((int)lane >= d) cincl += x;
        }
        const uint32_t ctotal = __shfl_sync(FULL_MASK, cincl, 31);
        if (ctotal) {
            s_pre[w][lane] = cincl;
            s_top[w][lane] = 0;               // hits of each query written so far
            __syncwarp();
            for (uint32_t kb = 0; kb < ctotal; kb += 32u) {
                const uint32_t k = kb + lane;
                const bool valid = k < ctotal;
                uint32_t own = 0;
                if (valid) {
#pragma unroll
                    for (uint32_t step = 16; step; step >>= 1) own += (s_pre[w][own + step - 1u] <= k) ? step : 0u;
                }
                const uint32_t excl = own ? s_pre[w][own - 1u] : 0u;
                // the owner's query lives in lane `own`: read it from there
                const uint32_t o_pl = __shfl_sync(FULL_MASK, plen, own), o_tm = __shfl_sync(FULL_MASK, topm1, own);
                const uint64_t o_l0 = __shfl_sync(FULL_MASK, l0, own), o_at = __shfl_sync(FULL_MASK, o, own);
                const int32_t o_qs = __shfl_sync(FULL_MASK, qs, own);
                uint32_t pos = 0; int32_t end = INT_MIN, val = 0; bool fl = false;
                bool hit = false;
                if (valid) {
                    const uint32_t c = k - excl;
                    SIB_STAB_CAND(c, o_pl, o_tm, o_l0, pos, end, val, fl)
                    hit = end >= o_qs;
                }
                const uint32_t hm = __ballot_sync(FULL_MASK, hit);
                // lanes of one query are neighbours: its segment starts at lane seg0 of this step
                const uint32_t seg0 = excl > kb ? excl - kb : 0u;
                const uint32_t seg = hm & ~((1u << seg0) - 1u);
                const uint32_t carry = valid ? s_top[w][own] : 0u;
                if (hit) { SIB_STAB_EMIT(o_at + carry + __popc(seg & lanemask_lt()), pos, end, val, fl) }
                __syncwarp();
                const uint32_t next_own = __shfl_down_sync(FULL_MASK, own, 1);
                const bool last = valid && (lane == 31u || k + 1u >= ctotal || next_own != own);
                if (last) s_top[w][own] = carry + __popc(seg & (lanemask_lt() | (1u << lane)));
                __syncwarp();
            }
        }
#undef SIB_STAB_CAND
#undef SIB_STAB_EMIT
        return;
    }
    uint32_t i = (more && top) ? top - 1u : NONE32;
    const bool small = more && (o_end - o) <= QK_FILL_LANE_MAX;
    if (small) {
        while (i != NONE32) {
            const uint32_t base = i & ~3u;
            const int4 e = ld_nc4(ix.ends + base);
            const uint32_t br = ld_nc(ix.branch + base);   // requested with the ends: a miss does not wait twice
            const uint32_t k = i - base;
            if (k >= 3u && e.w >= qs) emit<MODE>(ix, out, o++, base + 3u, e.w);
            if (k >= 2u && e.z >= qs) emit<MODE>(ix, out, o++, base + 2u, e.z);
            if (k >= 1u && e.y >= qs) emit<MODE>(ix, out, o++, base + 1u, e.y);
            if (e.x >= qs) emit<MODE>(ix, out, o++, base, e.x);
            if (o == o_end) break;
            i = (e.x >= qs) ? base - 1u : br;
        }
        i = NONE32;
    }
    uint32_t pending = __ballot_sync(FULL_MASK, more && !small);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        uint32_t bi = __shfl_sync(FULL_MASK, i, src);
        const int32_t bqs = __shfl_sync(FULL_MASK, qs, src);
        uint64_t bo = __shfl_sync(FULL_MASK, o, src);
        const uint64_t bo_end = __shfl_sync(FULL_MASK, o_end, src);
        while (bi != NONE32 && bo < bo_end) {
            const bool inb = lane <= bi;
            const uint32_t j = bi - lane;
            const int32_t e = inb ? ld_nc(ix.ends + j) : INT_MIN;
            const bool hit = inb && e >= bqs;
            const uint32_t hm = __ballot_sync(FULL_MASK, hit);
            if (hit) emit<MODE>(ix, out, bo + __popc(hm & lanemask_lt()), j, e);
            bo += __popc(hm);
            if (bi < 32u) break;
            const uint32_t low = bi - 31u;
            bi = (hm >> 31) ? low - 1u : ld_nc(ix.branch + low);
        }
    }
}

// ---- exclusive scan of counts -> 64-bit CSR offsets (single pass, decoupled look-back) --
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 16;
constexpr uint32_t SC_TILE = SC_THREADS * SC_ITEMS;
constexpr unsigned long long SC_FLAG_AGG = 1ull << 62;
constexpr unsigned long long SC_FLAG_INC = 2ull << 62;
constexpr unsigned long long SC_VAL_MASK = (1ull << 62) - 1;

// offsets has nq+1 entries; offsets[0] = 0, offsets[nq] = total hits.
__global__ void __launch_bounds__(SC_THREADS)
qk_scan_kernel(const uint32_t* __restrict__ counts, uint32_t nq, uint64_t* __restrict__ offsets,
               unsigned long long* __restrict__ status, uint32_t* __restrict__ ticket) {
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_warp[SC_THREADS / 32];
    __shared__ uint64_t s_prefix;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * SC_TILE + (uint64_t)tid * SC_ITEMS;

    uint32_t v[SC_ITEMS];
    if (base + SC_ITEMS <= nq) {
        const uint4* p = reinterpret_cast<const uint4*>(counts + base);
#pragma unroll
        for (int k = 0; k < SC_ITEMS / 4; ++k) {
            uint4 x = p[k];
            v[4 * k] = x.x; v[4 * k + 1] = x.y; v[4 * k + 2] = x.z; v[4 * k + 3] = x.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SC_ITEMS; ++k) v[k] = (base + k < nq) ? counts[base + k] : 0u;
    }
    uint64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) tsum += v[k];

    uint64_t incl = tsum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint64_t x = __shfl_up_sync(FULL_MASK, incl, off);
        if (lane >= (uint32_t)off) incl += x;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t wpre = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; ++w) {
        uint64_t x = s_warp[w];
        if (w < (int)warp) wpre += x;
        tile_total += x;
    }

    if (warp == 0) {
        // publish, then look back 32 tiles at a time
        if (lane == 0)
            *(volatile unsigned long long*)(status + tile) = (tile == 0 ? SC_FLAG_INC : SC_FLAG_AGG) | tile_total;
        uint64_t excl = 0;
        if (tile > 0) {
            int64_t look = (int64_t)tile - 1;
            while (true) {
                const int64_t mine = look - lane;
                unsigned long long w = SC_FLAG_INC;   // tiles before 0 count as an inclusive 0
                if (mine >= 0) {
                    do { w = *(volatile unsigned long long*)(status + mine); } while ((w & ~SC_VAL_MASK) == 0);
                }
                const uint32_t inc_mask = __ballot_sync(FULL_MASK, (w & ~SC_VAL_MASK) == SC_FLAG_INC);
                // nearest tile holding a full prefix; none in this window -> take all 32 aggregates
                const int first_inc = inc_mask ? __ffs(inc_mask) - 1 : 31;
                uint64_t contrib = ((int)lane <= first_inc) ? (w & SC_VAL_MASK) : 0ull;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) contrib += __shfl_xor_sync(FULL_MASK, contrib, off);
                excl += contrib;
                if (inc_mask) break;
                look -= 32;
            }
            if (lane == 0)
                *(volatile unsigned long long*)(status + tile) = SC_FLAG_INC | (excl + tile_total);
        }
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    uint64_t run = s_prefix + wpre + incl - tsum;   // exclusive prefix of this thread's first item
    if (base + SC_ITEMS <= nq) {
        uint64_t outv[SC_ITEMS];
#pragma unroll
        for (int k = 0; k < SC_ITEMS; ++k) { outv[k] = run; run += v[k]; }
        ulonglong2* po = reinterpret_cast<ulonglong2*>(offsets + base);
#pragma unroll
        for (int k = 0; k < SC_ITEMS / 2; ++k) po[k] = make_ulonglong2(outv[2 * k], outv[2 * k + 1]);
        if (base + SC_ITEMS == nq) offsets[nq] = run;
    } else {
#pragma unroll
        for (int k = 0; k < SC_ITEMS; ++k) {
            if (base + k < nq) { offsets[base + k] = run; run += v[k]; }
            if (base + k == nq) offsets[nq] = run;   // first slot past the end carries the total
        }
    }
}

// ---- query ordering helpers --------------------------------------------------------------
// *flag (init 1) is cleared when the query starts are not non-decreasing.
__global__ void __launch_bounds__(QK_THREADS)
qk_check_sorted_kernel(const int32_t* __restrict__ qe, uint32_t nq, uint32_t* __restrict__ flag) {
    uint32_t bad = 0;
    const uint64_t stride = (uint64_t)gridDim.x * QK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * QK_THREADS + threadIdx.x + 1; i < nq; i += stride)
        bad |= (qe[i] < qe[i - 1]) ? 1u : 0u;
    bad = __reduce_or_sync(FULL_MASK, bad);
    if (lane_id() == 0 && bad) *flag = 0;
}

__global__ void __launch_bounds__(QK_THREADS)
qk_widen_kernel(const uint32_t* __restrict__ in, uint32_t n, uint64_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * QK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * QK_THREADS + threadIdx.x; i < n; i += stride) out[i] = in[i];
}

}  // namespace sib
