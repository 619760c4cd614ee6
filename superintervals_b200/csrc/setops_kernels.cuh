// setops_kernels.cuh -- device side of the set-algebra callers of the query path
// (c_superintervals.h:823-1064; C++ twins superintervals.hpp:1037-1390).
//
// The reference runs these as sequential host loops around its queries. Here the
// interval lists live on the device and every operation is one of three shapes:
//   * sorted sweep (merge_overlaps / unique / gaps): the list in (start asc, end desc)
//     order -- build()'s radix sort -- then "does element i open a new output interval?"
//     from the exclusive prefix maximum of the ends (build()'s max tree), a scan of those
//     head flags, and a scatter;
//   * CSR consumer (intersection / difference): the stored intervals of A are a query
//     batch against B's index (count -> scan -> fill, query_kernels.cuh); one lane per
//     query then turns its hit list into output pieces (count pass, scan, write pass);
//   * elementwise (expand / flank / span): count, scan, scatter in stored order.
// Outputs are written in exactly the reference's emission order.
#pragma once

#include "common.cuh"
#include <limits.h>

namespace sib {

constexpr int SO_THREADS = 256;

struct Item3i { int32_t start, end, data; };   // == Interval (c_superintervals.h:61-65)

// ---- sorted sweep -------------------------------------------------------------------------
// merge_overlaps on a WELL-FORMED sorted list (c.h:854-881): the running end of the open
// cluster equals the maximum of ALL earlier ends (earlier clusters end before this one
// starts), so element i opens a cluster iff starts[i] > max(ends[0..i-1]).
// One warp per aligned 32-block; pmax32[b] = max(ends[0..32b-1]) comes from build().
// runmax[i] = max(ends[0..i]) is the cluster's end when i is its last element.
__global__ void __launch_bounds__(SO_THREADS)
so_merge_heads_kernel(const int32_t* __restrict__ starts, const int32_t* __restrict__ ends,
                      const int32_t* __restrict__ pmax32, uint32_t n, uint32_t* __restrict__ head,
                      int32_t* __restrict__ runmax) {
    const uint64_t i64 = (uint64_t)blockIdx.x * SO_THREADS + threadIdx.x;
    const uint32_t lane = lane_id();
    const bool live = i64 < n;
    const uint32_t i = (uint32_t)i64;
    if (__ballot_sync(FULL_MASK, live) == 0) return;
    const int32_t e = live ? ends[i] : INT_MIN;
    int32_t incl = e;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int32_t x = __shfl_up_sync(FULL_MASK, incl, d);
        if ((int)lane >= d) incl = max(incl, x);
    }
    const int32_t before = pmax32[i64 >> 5];          // warp-uniform; block 0: INT_MIN
    int32_t excl = __shfl_up_sync(FULL_MASK, incl, 1);
    excl = lane == 0 ? before : max(before, excl);
    if (live) {
        head[i] = (i == 0 || starts[i] > excl) ? 1u : 0u;
        runmax[i] = max(before, incl);
    }
}

// The same recurrence on ANY list (stored start > end allowed: the running end then is not a
// prefix maximum, it restarts at every flush). One thread runs the recurrence in order;
// only malformed inputs come here.
__global__ void so_merge_heads_seq_kernel(const int32_t* __restrict__ starts, const int32_t* __restrict__ ends,
                                          uint32_t n, uint32_t* __restrict__ head, int32_t* __restrict__ runmax) {
    if (blockIdx.x || threadIdx.x || n == 0) return;
    int32_t cur = ends[0];
    head[0] = 1;
    runmax[0] = cur;
    for (uint32_t i = 1; i < n; ++i) {
        const int32_t e = ends[i];
        if (starts[i] <= cur) { head[i] = 0; cur = max(cur, e); }
        else { head[i] = 1; cur = e; }
        runmax[i] = cur;
    }
}

// unique (c.h:1042-1064): exact (start, end) duplicates are neighbours in the sorted list
__global__ void __launch_bounds__(SO_THREADS)
so_unique_heads_kernel(const int32_t* __restrict__ starts, const int32_t* __restrict__ ends, uint32_t n,
                       uint32_t* __restrict__ head) {
    const uint64_t i = (uint64_t)blockIdx.x * SO_THREADS + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || starts[i] != starts[i - 1] || ends[i] != ends[i - 1]) ? 1u : 0u;
}

// off = exclusive scan of head (off[n] = clusters). A head writes its cluster's start and
// data (the first element's: "keep first"), the cluster's last element writes its end.
__global__ void __launch_bounds__(SO_THREADS)
so_cluster_scatter_kernel(const int32_t* __restrict__ starts, const int32_t* __restrict__ ends,
                          const int32_t* __restrict__ values, const uint32_t* __restrict__ head,
                          const uint64_t* __restrict__ off, const int32_t* __restrict__ runmax, uint32_t n,
                          int32_t* __restrict__ out_s, int32_t* __restrict__ out_e, int32_t* __restrict__ out_d) {
    const uint64_t i = (uint64_t)blockIdx.x * SO_THREADS + threadIdx.x;
    if (i >= n) return;
    const uint32_t h = head[i];
    const uint64_t c = off[i] + h - 1;          // cluster of element i
    if (h) { out_s[c] = starts[i]; out_d[c] = values[i]; }
    if (i + 1 == n || head[i + 1]) out_e[c] = runmax ? runmax[i] : ends[i];
}

// gaps over a merged list from a well-formed set (disjoint, starts and ends ascending;
// c.h:883-905). The entries inside [lo, hi] are one contiguous range; the cursor before an
// in-span entry is lo for the first of them and the previous end + 1 afterwards. Each
// in-span entry emits the gap before it; the last one also emits the trailing gap.
// WRITE = false: cnt[k] = pieces of entry k. WRITE = true: write them at poff[k].
// *any_in is set when some entry lies in the span (otherwise the host emits [lo, hi]).
template <bool WRITE>
__global__ void __launch_bounds__(SO_THREADS)
so_gaps_kernel(const int32_t* __restrict__ ms, const int32_t* __restrict__ me, uint32_t m, int32_t lo, int32_t hi,
               int32_t fill, uint32_t* __restrict__ cnt, const uint64_t* __restrict__ poff, uint32_t* __restrict__ any_in,
               int32_t* __restrict__ out_s, int32_t* __restrict__ out_e, int32_t* __restrict__ out_d) {
    const uint64_t k = (uint64_t)blockIdx.x * SO_THREADS + threadIdx.x;
    if (k >= m) return;
    const int32_t s = ms[k], e = me[k];
    const bool in = !(e < lo || s > hi);
    uint32_t c = 0;
    uint64_t o = WRITE ? poff[k] : 0;
    if (in) {
        const bool prev_in = k > 0 && !(me[k - 1] < lo || ms[k - 1] > hi);
        const int64_t cursor = prev_in ? max((int64_t)lo, (int64_t)me[k - 1] + 1) : (int64_t)lo;
        if ((int64_t)s > cursor) {
            if (WRITE) { out_s[o] = (int32_t)cursor; out_e[o] = s - 1; out_d[o] = fill; ++o; }
            ++c;
        }
        const bool next_in = k + 1 < m && !(me[k + 1] < lo || ms[k + 1] > hi);
        if (!next_in) {
            const int64_t after = max(cursor, (int64_t)e + 1);
            if (after <= (int64_t)hi) {
                if (WRITE) { out_s[o] = (int32_t)after; out_e[o] = hi; out_d[o] = fill; ++o; }
                ++c;
            }
        }
        if (!WRITE) *any_in = 1u;
    }
    if (!WRITE) cnt[k] = c;
}

// gaps over ANY merged list: the same cursor recurrence run by one thread (malformed sets only).
// out_* hold at most m + 1 pieces; *count_out receives how many were written.
__global__ void so_gaps_seq_kernel(const int32_t* __restrict__ ms, const int32_t* __restrict__ me, uint32_t m,
                                   int32_t lo, int32_t hi, int32_t fill, int32_t* __restrict__ out_s,
                                   int32_t* __restrict__ out_e, int32_t* __restrict__ out_d,
                                   unsigned long long* __restrict__ count_out) {
    if (blockIdx.x || threadIdx.x) return;
    int64_t cursor = lo;
    unsigned long long o = 0;
    for (uint32_t k = 0; k < m; ++k) {
        const int32_t s = ms[k], e = me[k];
        if (e < lo || s > hi) continue;
        if ((int64_t)s > cursor) { out_s[o] = (int32_t)cursor; out_e[o] = s - 1; out_d[o] = fill; ++o; }
        if ((int64_t)e + 1 > cursor) cursor = (int64_t)e + 1;
    }
    if (cursor <= (int64_t)hi) { out_s[o] = (int32_t)cursor; out_e[o] = hi; out_d[o] = fill; ++o; }
    *count_out = o;
}

// ---- CSR consumers --------------------------------------------------------------------------
// intersection (c.h:921-939): query k = stored interval k of A; items = B's (start, end, data)
// per hit in descending position (FILL_ITEMS). A hit contributes the clipped piece when it is
// non-empty. out_da / out_db carry both payloads: the host applies `combine` (a callback) or
// keeps out_da.
template <bool WRITE>
__global__ void __launch_bounds__(SO_THREADS)
so_intersection_kernel(const int32_t* __restrict__ qs, const int32_t* __restrict__ qe, const int32_t* __restrict__ qd,
                       uint32_t nq, const uint64_t* __restrict__ offsets, const Item3i* __restrict__ items,
                       uint32_t* __restrict__ cnt, const uint64_t* __restrict__ poff, int32_t* __restrict__ out_s,
                       int32_t* __restrict__ out_e, int32_t* __restrict__ out_da, int32_t* __restrict__ out_db) {
    const uint64_t k = (uint64_t)blockIdx.x * SO_THREADS + threadIdx.x;
    if (k >= nq) return;
    const int32_t s = qs[k], e = qe[k];
    const int32_t d = WRITE ? qd[k] : 0;
    uint32_t c = 0;
    uint64_t o = WRITE ? poff[k] : 0;
    const uint64_t p1 = offsets[k + 1];
    for (uint64_t p = offsets[k]; p < p1; ++p) {
        const Item3i it = items[p];
        const int32_t ps = max(s, it.start), pe = min(e, it.end);
        if (ps <= pe) {
            if (WRITE) { out_s[o] = ps; out_e[o] = pe; out_da[o] = d; out_db[o] = it.data; ++o; }
            ++c;
        }
    }
    if (!WRITE) cnt[k] = c;
}

// difference (c.h:950-974): keys = B's (start, end) per hit in descending position
// (FILL_KEYS). The reference visits them by (start asc, end asc): that is ascending
// position with every run of equal starts reversed (inside a run the position order is end
// DESC). A cursor sweeps [s, e]; the stretches no pair covers are emitted with A's data.
template <bool WRITE>
__global__ void __launch_bounds__(SO_THREADS)
so_difference_kernel(const int32_t* __restrict__ qs, const int32_t* __restrict__ qe, const int32_t* __restrict__ qd,
                     uint32_t nq, const uint64_t* __restrict__ offsets, const int2* __restrict__ keys,
                     uint32_t* __restrict__ cnt, const uint64_t* __restrict__ poff, int32_t* __restrict__ out_s,
                     int32_t* __restrict__ out_e, int32_t* __restrict__ out_d) {
    const uint64_t k = (uint64_t)blockIdx.x * SO_THREADS + threadIdx.x;
    if (k >= nq) return;
    const int32_t s = qs[k], e = qe[k];
    const int32_t d = WRITE ? qd[k] : 0;
    uint32_t c = 0;
    uint64_t o = WRITE ? poff[k] : 0;
    int64_t cursor = s;
    const uint64_t p0 = offsets[k];
    uint64_t b = offsets[k + 1];             // runs are taken from the back: ascending position
    while (b > p0) {
        uint64_t a = b - 1;
        const int32_t run_start = keys[a].x;
        while (a > p0 && keys[a - 1].x == run_start) --a;
        for (uint64_t t = a; t < b; ++t) {   // inside the run: descending position = end ascending
            const int2 kp = keys[t];
            const int32_t cs = max(kp.x, s), ce = min(kp.y, e);
            if ((int64_t)cs > cursor) {
                if (WRITE) { out_s[o] = (int32_t)cursor; out_e[o] = cs - 1; out_d[o] = d; ++o; }
                ++c;
            }
            if ((int64_t)ce + 1 > cursor) cursor = (int64_t)ce + 1;
        }
        b = a;
    }
    if (cursor <= (int64_t)e) {
        if (WRITE) { out_s[o] = (int32_t)cursor; out_e[o] = e; out_d[o] = d; ++o; }
        ++c;
    }
    if (!WRITE) cnt[k] = c;
}

// ---- elementwise ----------------------------------------------------------------------------
// expand (c.h:1000-1016): 64-bit arithmetic, clamp to [lo, hi], drop what shrank past itself.
// flank  (c.h:1018-1040): left strip first, then right strip; originals are not emitted.
template <bool FLANK, bool WRITE>
__global__ void __launch_bounds__(SO_THREADS)
so_resize_kernel(const int32_t* __restrict__ ss, const int32_t* __restrict__ ee, const int32_t* __restrict__ dd,
                 uint32_t n, int32_t left, int32_t right, int32_t lo, int32_t hi, uint32_t* __restrict__ cnt,
                 const uint64_t* __restrict__ poff, int32_t* __restrict__ out_s, int32_t* __restrict__ out_e,
                 int32_t* __restrict__ out_d) {
    const uint64_t k = (uint64_t)blockIdx.x * SO_THREADS + threadIdx.x;
    if (k >= n) return;
    const int32_t s = ss[k], e = ee[k];
    const int32_t d = WRITE ? dd[k] : 0;
    uint32_t c = 0;
    uint64_t o = WRITE ? poff[k] : 0;
    if (!FLANK) {
        int64_t a = (int64_t)s - left, b = (int64_t)e + right;
        if (a < lo) a = lo;
        if (b > hi) b = hi;
        if (a <= b) {
            if (WRITE) { out_s[o] = (int32_t)a; out_e[o] = (int32_t)b; out_d[o] = d; }
            ++c;
        }
    } else {
        if (left > 0 && s > lo) {
            const int32_t le = s - 1;
            int64_t ls = (int64_t)s - left;
            if (ls < lo) ls = lo;
            if (ls <= (int64_t)le && le <= hi) {
                if (WRITE) { out_s[o] = (int32_t)ls; out_e[o] = le; out_d[o] = d; ++o; }
                ++c;
            }
        }
        if (right > 0 && e < hi) {
            const int32_t rs = e + 1;
            int64_t re = (int64_t)e + right;
            if (re > hi) re = hi;
            if ((int64_t)rs <= re && rs >= lo) {
                if (WRITE) { out_s[o] = rs; out_e[o] = (int32_t)re; out_d[o] = d; ++o; }
                ++c;
            }
        }
    }
    if (!WRITE) cnt[k] = c;
}

// span (c.h:985-998): res[0] = min start, res[1] = max end; caller initialises INT_MAX / INT_MIN
__global__ void __launch_bounds__(SO_THREADS)
so_span_kernel(const int32_t* __restrict__ ss, const int32_t* __restrict__ ee, uint32_t n, int32_t* __restrict__ res) {
    int32_t lo = INT_MAX, hi = INT_MIN;
    const uint64_t stride = (uint64_t)gridDim.x * SO_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * SO_THREADS + threadIdx.x; i < n; i += stride) {
        lo = min(lo, ss[i]);
        hi = max(hi, ee[i]);
    }
    lo = __reduce_min_sync(FULL_MASK, lo);
    hi = __reduce_max_sync(FULL_MASK, hi);
    if (lane_id() == 0) {
        atomicMin(res, lo);
        atomicMax(res + 1, hi);
    }
}

}  // namespace sib
