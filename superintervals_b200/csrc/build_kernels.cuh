// build_kernels.cuh -- device-side build() for the superset index (sm_100a).
//
//   reference build()  = sort_intervals() (superintervals.hpp:1396-1439)
//                      + sequential monotonic-stack branch loop (hpp:117-129)
//   here               = key packing -> radix_sort.cuh -> gather
//                      + a parallel "all nearest previous end >= mine" (ANSV) pass:
//                        in-warp resolution by shuffles, a prefix-max early-out for
//                        roots, and a 32-ary max-tree descent for the far tail.
#pragma once

#include "common.cuh"
#include <limits.h>

namespace sib {

constexpr int BK_THREADS = 256;

// ---- sortedness, as add() tracks it (hpp:96-101) ------------------------------------
// flags bit0 = start_sorted, bit1 = end_sorted, bit2 = every interval has start <= end;
// caller initialises *flags = 7.
__global__ void __launch_bounds__(BK_THREADS)
bk_check_sorted_kernel(const int32_t* __restrict__ s, const int32_t* __restrict__ e, uint32_t n,
                       uint32_t* __restrict__ flags) {
    uint32_t bad = 0;
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) {
        const int32_t s1 = s[i], e1 = e[i];
        if (s1 > e1) bad |= 4u;
        if (i == 0) continue;
        const int32_t s0 = s[i - 1];
        if (s1 < s0) bad |= 1u;
        else if (s1 == s0 && e1 > e[i - 1]) bad |= 2u;
    }
    bad = __reduce_or_sync(FULL_MASK, bad);
    if (lane_id() == 0 && bad) atomicAnd(flags, ~bad);
}

// ---- stable grouping of records by a small non-negative key (BED ingest: contig id) -----
__global__ void __launch_bounds__(BK_THREADS)
bk_iota_key_kernel(const int32_t* __restrict__ key, uint32_t n, uint32_t* __restrict__ k, uint32_t* __restrict__ v) {
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) {
        k[i] = (uint32_t)key[i];
        v[i] = (uint32_t)i;
    }
}
__global__ void __launch_bounds__(BK_THREADS)
bk_take_perm_kernel(const uint32_t* __restrict__ vA, const uint32_t* __restrict__ vB, const uint32_t* __restrict__ final_sel,
                    uint32_t n, uint32_t* __restrict__ perm) {
    const uint32_t* __restrict__ v = *final_sel ? vB : vA;
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) perm[i] = v[i];
}

// ---- mode B routing: a mixed query batch grouped by contig id -------------------------------
// routed columns through the stable order of the contig ids; the routed ids stay for the offsets search
__global__ void __launch_bounds__(BK_THREADS)
bk_route_gather_kernel(const uint32_t* __restrict__ perm, const int32_t* __restrict__ c, const int32_t* __restrict__ qs,
                       const int32_t* __restrict__ qe, uint32_t n, uint32_t contigs, int32_t* __restrict__ gc,
                       int32_t* __restrict__ gs, int32_t* __restrict__ ge, unsigned long long* __restrict__ bad) {
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) {
        const uint32_t j = perm[i];
        const int32_t cj = c[j];
        if ((uint32_t)cj >= contigs) *bad = 1ull;     // id outside [0, contigs): the host reports it
        gc[i] = cj;
        gs[i] = qs[j];
        ge[i] = qe[j];
    }
}
// offsets[k] = first routed query of contig k (k = 0..contigs; offsets[contigs] = n)
__global__ void bk_key_offsets_kernel(const int32_t* __restrict__ gc, uint32_t n, uint32_t contigs,
                                      unsigned long long* __restrict__ offsets) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > contigs) return;
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if ((uint32_t)gc[mid] < k) lo = mid + 1; else hi = mid;
    }
    offsets[k] = lo;
}
// out[perm[i]] = v[i]: the routed counts back in the caller's order
__global__ void __launch_bounds__(BK_THREADS)
bk_scatter_u32_kernel(const uint32_t* __restrict__ v, const uint32_t* __restrict__ perm, uint32_t n, uint32_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) out[perm[i]] = v[i];
}

// uint32 counts -> the size_t of the C ABI, on the device (the copy back then lands in the caller's array directly)
__global__ void __launch_bounds__(BK_THREADS)
bk_widen_kernel(const uint32_t* __restrict__ in, uint64_t n, unsigned long long* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) out[i] = in[i];
}

// ---- key packing: (start asc, end DESC, insertion idx) ------------------------------
__global__ void __launch_bounds__(BK_THREADS)
bk_make_keys_kernel(const int32_t* __restrict__ s, const int32_t* __restrict__ e, uint32_t n,
                    uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) {
        uint32_t hs = flip_i32(s[i]);
        uint32_t he = ~flip_i32(e[i]);   // complement -> descending end under an ascending sort
        keys[i] = ((uint64_t)hs << 32) | he;
        idx[i] = (uint32_t)i;
    }
}

// ---- gather: sorted keys carry start and end; only the payload needs the permutation -
__global__ void __launch_bounds__(BK_THREADS)
bk_gather_kernel(const uint64_t* __restrict__ kA, const uint64_t* __restrict__ kB,
                 const uint32_t* __restrict__ vA, const uint32_t* __restrict__ vB,
                 const uint32_t* __restrict__ final_sel, const int32_t* __restrict__ values_in,
                 uint32_t n, int32_t* __restrict__ starts, int32_t* __restrict__ ends,
                 int32_t* __restrict__ values, uint32_t* __restrict__ perm) {
    const bool useB = *final_sel != 0;
    const uint64_t* __restrict__ keys = useB ? kB : kA;
    const uint32_t* __restrict__ idx = useB ? vB : vA;
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) {
        uint64_t k = keys[i];
        uint32_t j = idx[i];
        starts[i] = unflip_i32((uint32_t)(k >> 32));
        ends[i] = unflip_i32(~(uint32_t)k);
        values[i] = values_in ? values_in[j] : (int32_t)j;
        perm[i] = j;
    }
}

// ---- narrow sort: by START only (32-bit keys, half the passes of the composite key and 16 instead of 24 bytes per
// record and pass), then the ties among equal starts put in end-DESCENDING order in place. Equal starts are rare on
// genomic data and their runs short; a run longer than BK_TIE_MAX raises *overflow and the host redoes the build with
// the 64-bit composite key. The order is the composite sort's exactly: the radix sort is stable (equal starts arrive in
// insertion order) and the in-run insertion sort moves an element only past strictly smaller ends.
constexpr uint32_t BK_TIE_MAX = 16;

__global__ void __launch_bounds__(BK_THREADS)
bk_make_start_keys_kernel(const int32_t* __restrict__ s, uint32_t n, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) {
        keys[i] = flip_i32(s[i]);
        idx[i] = (uint32_t)i;
    }
}

__global__ void __launch_bounds__(BK_THREADS)
bk_gather_narrow_kernel(const uint32_t* __restrict__ kA, const uint32_t* __restrict__ kB, const uint32_t* __restrict__ vA,
                        const uint32_t* __restrict__ vB, const uint32_t* __restrict__ sel, const int32_t* __restrict__ ends_in,
                        const int32_t* __restrict__ values_in, uint32_t n, int32_t* __restrict__ starts, int32_t* __restrict__ ends,
                        int32_t* __restrict__ values, uint32_t* __restrict__ perm) {
    const bool useB = *sel != 0;
    const uint32_t* __restrict__ keys = useB ? kB : kA;
    const uint32_t* __restrict__ idx = useB ? vB : vA;
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) {
        const uint32_t j = idx[i];
        starts[i] = unflip_i32(keys[i]);
        ends[i] = ends_in[j];
        values[i] = values_in ? values_in[j] : (int32_t)j;
        perm[i] = j;
    }
}

__global__ void __launch_bounds__(BK_THREADS)
bk_fix_ties_kernel(const int32_t* __restrict__ starts, int32_t* __restrict__ ends, int32_t* __restrict__ values,
                   uint32_t* __restrict__ perm, uint32_t n, uint32_t* __restrict__ overflow) {
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i + 1 < n; i += stride) {
        const int32_t s = starts[i];
        if (starts[i + 1] != s || (i > 0 && starts[i - 1] == s)) continue;    // not the head of a run of >= 2
        uint64_t r = i + 2;
        while (r < n && r - i <= BK_TIE_MAX && starts[r] == s) ++r;
        if (r - i > BK_TIE_MAX) { *overflow = 1u; continue; }
        for (uint64_t a = i + 1; a < r; ++a) {                                  // stable insertion sort, end descending
            const int32_t e = ends[a], v = values[a];
            const uint32_t p = perm[a];
            uint64_t b = a;
            while (b > i && ends[b - 1] < e) {
                ends[b] = ends[b - 1]; values[b] = values[b - 1]; perm[b] = perm[b - 1];
                --b;
            }
            ends[b] = e; values[b] = v; perm[b] = p;
        }
    }
}

// already (start asc, end desc): the reference performs no sort (hpp:1416,1421)
__global__ void __launch_bounds__(BK_THREADS)
bk_identity_kernel(const int32_t* __restrict__ s, const int32_t* __restrict__ e,
                   const int32_t* __restrict__ values_in, uint32_t n, int32_t* __restrict__ starts,
                   int32_t* __restrict__ ends, int32_t* __restrict__ values, uint32_t* __restrict__ perm) {
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) {
        starts[i] = s[i];
        ends[i] = e[i];
        values[i] = values_in ? values_in[i] : (int32_t)i;
        perm[i] = (uint32_t)i;
    }
}

// pad tail [n, n_padded) so 128-bit loads past the end read harmless data
__global__ void bk_pad_kernel(int32_t* __restrict__ starts, int32_t* __restrict__ ends, int32_t* __restrict__ values,
                              uint32_t* __restrict__ branch, uint32_t n, uint32_t n_padded) {
    uint32_t i = n + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_padded) {
        starts[i] = INT_MAX;
        ends[i] = INT_MIN;
        values[i] = 0;        // read (never used) by the fill's aligned 128-bit payload loads
        branch[i] = NONE32;
    }
}

// ---- esort: ends with every aligned 32-block sorted ascending (bitonic network in a warp) ----
// Lets the count sweep rank a query start inside a block with 6 probes (query_kernels.cuh).
__global__ void __launch_bounds__(BK_THREADS)
bk_sort_blocks_kernel(const int32_t* __restrict__ ends, uint32_t n_padded, int32_t* __restrict__ esort) {
    const uint32_t lane = lane_id();
    const uint64_t warps_total = (uint64_t)gridDim.x * (BK_THREADS / 32);
    const uint64_t nblocks = n_padded >> 5;
    for (uint64_t b = (uint64_t)blockIdx.x * (BK_THREADS / 32) + (threadIdx.x >> 5); b < nblocks; b += warps_total) {
        int32_t v = ends[b * 32u + lane];
#pragma unroll
        for (uint32_t k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                const int32_t o = __shfl_xor_sync(FULL_MASK, v, j);
                const bool asc = (lane & k) == 0;
                const bool low = (lane & j) == 0;
                v = (low == asc) ? min(v, o) : max(v, o);
            }
        }
        esort[b * 32u + lane] = v;
    }
}

// ---- eall: every end, ascending (count by rank, query_kernels.cuh) ----------------------
// Malformed intervals (start > end) take the largest key: they sort behind every well-formed end and are
// cut off (the closed-form rank tables cover the well-formed intervals only; index.cu).
__global__ void __launch_bounds__(BK_THREADS)
bk_end_keys_kernel(const int32_t* __restrict__ starts, const int32_t* __restrict__ ends, uint32_t n, uint32_t* __restrict__ keys) {
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride)
        keys[i] = starts[i] > ends[i] ? 0xFFFFFFFFu : flip_i32(ends[i]);
}

// ---- the (few) malformed intervals of a built index: positions, starts and ends, at most `cap` of them ----
// out[0] = how many exist (may exceed cap); out[1 + 3k ..] = (position, start, end) of the first cap found
// (in no particular order: the host sorts them by position).
__global__ void __launch_bounds__(BK_THREADS)
bk_find_malformed_kernel(const int32_t* __restrict__ starts, const int32_t* __restrict__ ends, uint32_t n, uint32_t cap,
                         uint32_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) {
        const int32_t s = starts[i], e = ends[i];
        if (s > e) {
            const uint32_t k = atomicAdd(out, 1u);
            if (k < cap) { out[1 + 3 * k] = (uint32_t)i; out[2 + 3 * k] = (uint32_t)s; out[3 + 3 * k] = (uint32_t)e; }
        }
    }
}

constexpr int BK_MAX_MALFORMED = 8;
struct MalformedList { uint32_t n; uint32_t pos[BK_MAX_MALFORMED]; };
// starts of the well-formed intervals only, order kept: position i moves down by the malformed positions below it
__global__ void __launch_bounds__(BK_THREADS)
bk_compact_wellformed_kernel(const int32_t* __restrict__ starts, uint32_t n, MalformedList mal, uint32_t n_padded_out,
                             int32_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += stride) {
        uint32_t below = 0;
        bool is_mal = false;
        for (uint32_t k = 0; k < mal.n; ++k) { below += mal.pos[k] < i ? 1u : 0u; is_mal |= mal.pos[k] == i; }
        if (!is_mal) out[i - below] = starts[i];
    }
    // pad like the index arrays (128-bit loads never leave the array)
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x + (n - mal.n); i < n_padded_out; i += stride) out[i] = INT_MAX;
}

__global__ void __launch_bounds__(BK_THREADS)
bk_sorted_ends_kernel(const uint32_t* __restrict__ kA, const uint32_t* __restrict__ kB,
                      const uint32_t* __restrict__ final_sel, uint32_t n, uint32_t n_padded,
                      int32_t* __restrict__ eall) {
    const uint32_t* __restrict__ keys = *final_sel ? kB : kA;
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; i < n_padded; i += stride)
        eall[i] = i < n ? unflip_i32(keys[i]) : INT_MAX;
}

// ---- rank grid: #{starts < v}, #{ends < v} at the points v = lo + c * 2^shift, c = 0..cells ----
__global__ void __launch_bounds__(BK_THREADS)
bk_rank_grid_kernel(const int32_t* __restrict__ starts, const int32_t* __restrict__ eall, uint32_t n, int32_t lo,
                    uint32_t shift, uint32_t cells, uint32_t* __restrict__ tab_s, uint32_t* __restrict__ tab_e) {
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t c = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; c <= cells; c += stride) {
        const int64_t v = (int64_t)lo + (int64_t)(c << shift);
        uint32_t rs = n, re = n;
        if (v <= (int64_t)INT_MAX) {
            // branch-free halving search for #{a < v} (same shape as hpp:501-513)
            uint32_t ps = 0, pe = 0, len = n;
            while (len > 1) {
                const uint32_t half = len >> 1;
                ps += (starts[ps + half] < (int32_t)v) ? (len - half) : 0u;
                pe += (eall[pe + half] < (int32_t)v) ? (len - half) : 0u;
                len = half;
            }
            rs = ps + ((starts[ps] < (int32_t)v) ? 1u : 0u);
            re = pe + ((eall[pe] < (int32_t)v) ? 1u : 0u);
        }
        tab_s[c] = rs;
        tab_e[c] = re;
    }
}

// ---- rank cells: one 32-byte record per 2^shift coordinates of a sorted array A -------------
// (layout: RankCells in query_kernels.cuh). FMT 1: 28 one-byte offsets, FMT 2: 14 two-byte
// offsets. Records 0..cells; the last one lies beyond A[n-1] and carries word 0 = n.
// *overfull counts the cells whose values do not fit (they are answered from A itself).
template <int FMT>
__global__ void __launch_bounds__(BK_THREADS)
bk_rank_cells_kernel(const int32_t* __restrict__ A, uint32_t n, int32_t lo, uint32_t shift, uint32_t cells,
                     uint4* __restrict__ rec, unsigned long long* __restrict__ overfull) {
    constexpr uint32_t SLOTS = FMT == 1 ? 28u : 14u;
    const uint64_t stride = (uint64_t)gridDim.x * BK_THREADS;
    for (uint64_t c = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x; c <= cells; c += stride) {
        const int64_t v0 = (int64_t)lo + (int64_t)(c << shift);
        const int64_t v1 = v0 + ((int64_t)1 << shift);
        // #{A < v0}, #{A < v1}: halving searches compared in 64 bits (v1 may exceed INT_MAX)
        uint32_t p0 = 0, p1 = 0, len = n;
        while (len > 1) {
            const uint32_t half = len >> 1;
            p0 += ((int64_t)A[p0 + half] < v0) ? (len - half) : 0u;
            p1 += ((int64_t)A[p1 + half] < v1) ? (len - half) : 0u;
            len = half;
        }
        p0 += ((int64_t)A[p0] < v0) ? 1u : 0u;
        p1 += ((int64_t)A[p1] < v1) ? 1u : 0u;
        const uint32_t cnt = p1 - p0;
        uint32_t w[8];
        w[0] = p0;
#pragma unroll
        for (int k = 1; k < 8; ++k) w[k] = 0xFFFFFFFFu;
        if (cnt > SLOTS) {
            w[0] |= 0x80000000u;
            atomicAdd(overfull, 1ull);
        } else {
#pragma unroll
            for (uint32_t k = 0; k < SLOTS; ++k) {
                if (k < cnt) {
                    const uint32_t off = (uint32_t)((int64_t)A[p0 + k] - v0);
                    if (FMT == 1) {
                        const uint32_t sh = (k & 3u) * 8u;
                        w[1 + (k >> 2)] = (w[1 + (k >> 2)] & ~(0xFFu << sh)) | (off << sh);
                    } else {
                        const uint32_t sh = (k & 1u) * 16u;
                        w[1 + (k >> 1)] = (w[1 + (k >> 1)] & ~(0xFFFFu << sh)) | (off << sh);
                    }
                }
            }
        }
        rec[2 * c] = make_uint4(w[0], w[1], w[2], w[3]);
        rec[2 * c + 1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// ---- 32-ary max tree over ends -------------------------------------------------------
// level 0 = ends; level L entry k = max of level L-1 entries [32k, 32k+32).
// One launch produces two levels: a CTA of 1024 threads folds 1024 inputs into
// 32 level-(L+1) entries and 1 level-(L+2) entry.
__global__ void __launch_bounds__(1024)
bk_max2_kernel(const int32_t* __restrict__ in, uint32_t n_in, int32_t* __restrict__ out1,
               int32_t* __restrict__ out2) {
    __shared__ int32_t s_m[32];
    const uint64_t i = (uint64_t)blockIdx.x * 1024u + threadIdx.x;
    int32_t v = i < n_in ? in[i] : INT_MIN;
    v = __reduce_max_sync(FULL_MASK, v);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (lane == 0) {
        s_m[warp] = v;
        uint64_t o = (uint64_t)blockIdx.x * 32u + warp;
        if (o * 32u < n_in) out1[o] = v;
    }
    __syncthreads();
    if (warp == 0) {
        int32_t m = __reduce_max_sync(FULL_MASK, s_m[lane]);
        if (lane == 0) out2[blockIdx.x] = m;
    }
}

// Exclusive prefix max per entry of one level, top-down:
//   P_L[k] = max(P_{L+1}[k/32], max of M_L over the left siblings of k)
// parentP == nullptr at the top level (a single group).
__global__ void __launch_bounds__(BK_THREADS)
bk_prefix_level_kernel(const int32_t* __restrict__ M, uint32_t n_level,
                       const int32_t* __restrict__ parentP, int32_t* __restrict__ P) {
    const uint64_t k = (uint64_t)blockIdx.x * BK_THREADS + threadIdx.x;
    const uint32_t lane = lane_id();
    int32_t m = k < n_level ? M[k] : INT_MIN;
    int32_t incl = m;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int32_t t = __shfl_up_sync(FULL_MASK, incl, off);
        if (lane >= (uint32_t)off) incl = max(incl, t);
    }
    int32_t excl = __shfl_up_sync(FULL_MASK, incl, 1);
    if (lane == 0) excl = INT_MIN;
    if (k < n_level) {
        int32_t up = parentP ? parentP[k >> 5] : INT_MIN;
        P[k] = max(up, excl);
    }
}

constexpr int BK_MAX_LEVELS = 7;   // 32^7 > 2^32
struct MaxTree {
    const int32_t* M[BK_MAX_LEVELS + 1];   // M[0] = ends
    uint32_t n[BK_MAX_LEVELS + 1];
    const int32_t* P1;                     // exclusive prefix max per 32-block of ends
    int top;                               // highest level with n[top] >= 1
};

// ---- branch[i] = nearest j < i with ends[j] >= ends[i], else NONE (hpp:117-129) -------
// One warp per 32 consecutive intervals.
__global__ void __launch_bounds__(BK_THREADS)
bk_branch_kernel(MaxTree t, uint32_t n, uint32_t* __restrict__ branch) {
    const uint32_t lane = lane_id();
    const uint64_t warps_total = (uint64_t)gridDim.x * (BK_THREADS / 32);
    const uint32_t nblocks = t.n[1];
    for (uint64_t b = (uint64_t)blockIdx.x * (BK_THREADS / 32) + (threadIdx.x >> 5); b < nblocks; b += warps_total) {
        const uint64_t i = b * 32u + lane;
        const bool live = i < n;
        const int32_t e = live ? t.M[0][i] : INT_MIN;

        // exclusive prefix max inside the block: lanes above it have no in-block answer
        int32_t incl = e;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int32_t v = __shfl_up_sync(FULL_MASK, incl, off);
            if (lane >= (uint32_t)off) incl = max(incl, v);
        }
        int32_t pm = __shfl_up_sync(FULL_MASK, incl, 1);
        const bool in_block = live && lane > 0 && pm >= e;

        uint32_t res = NONE32;
        bool found = !in_block;
        // nearest previous lane with end >= mine; distances are short on real data
        for (int k = 1; k < 32; ++k) {
            if (__all_sync(FULL_MASK, found)) break;
            int32_t v = __shfl_up_sync(FULL_MASK, e, k);
            if (!found && lane >= (uint32_t)k && v >= e) {
                found = true;
                res = (uint32_t)(i - k);
            }
        }

        if (live && !in_block && b > 0 && e <= t.P1[b]) {
            // some earlier interval reaches at least as far: climb the max tree
            int level = 1;
            uint64_t k = b;
            int64_t hit = -1;
            while (true) {
                const uint64_t g0 = k & ~(uint64_t)31;
                for (int64_t kk = (int64_t)k - 1; kk >= (int64_t)g0; --kk) {
                    if (ld_nc(t.M[level] + kk) >= e) { hit = kk; break; }
                }
                if (hit >= 0 || g0 == 0 || level == t.top) break;
                k >>= 5;
                ++level;
            }
            if (hit >= 0) {
                while (level > 0) {
                    const uint64_t lo = (uint64_t)hit * 32u;
                    uint64_t c = lo + 31u;
                    const uint64_t last = (uint64_t)t.n[level - 1] - 1;
                    if (c > last) c = last;
                    // the subtree max is >= e, so a child >= e exists: take the last one
                    while (c > lo && ld_nc(t.M[level - 1] + c) < e) --c;
                    hit = (int64_t)c;
                    --level;
                }
                res = (uint32_t)hit;
            }
        }
        if (live) branch[i] = res;
    }
}

}  // namespace sib
