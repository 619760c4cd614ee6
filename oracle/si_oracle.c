/*
 * si_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See si_oracle.h for scope, citations and pinning status.
 *
 * Everything is integer arithmetic: int32 coordinates/payloads, size_t indices.
 */
#include "si_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---- ordering: (start asc, end DESC), ties keep insertion order ----------
 * hpp:1415-1420 sorts with std::sort on that comparator (unstable for exact
 * (start,end) duplicates). The oracle fixes the tie order to
 * insertion order, which is what the reference itself yields whenever its
 * input arrives already sorted (hpp:1416/1421 skip the sort).             */
static int before(const int32_t* s, const int32_t* e, uint32_t a, uint32_t b) {
    if (s[a] != s[b]) return s[a] < s[b];
    return e[a] > e[b];
}

static void merge_sort_perm(uint32_t* perm, uint32_t* tmp, size_t lo, size_t hi,
                            const int32_t* s, const int32_t* e) {
    if (hi - lo < 2) return;
    if (hi - lo <= 12) { /* insertion sort: stable */
        for (size_t i = lo + 1; i < hi; ++i) {
            uint32_t v = perm[i];
            size_t k = i;
            while (k > lo && before(s, e, v, perm[k - 1])) { perm[k] = perm[k - 1]; --k; }
            perm[k] = v;
        }
        return;
    }
    size_t mid = lo + (hi - lo) / 2;
    merge_sort_perm(perm, tmp, lo, mid, s, e);
    merge_sort_perm(perm, tmp, mid, hi, s, e);
    size_t a = lo, b = mid, o = lo;
    while (a < mid && b < hi) {
        /* take right only if strictly before left -> stability */
        if (before(s, e, perm[b], perm[a])) tmp[o++] = perm[b++];
        else tmp[o++] = perm[a++];
    }
    while (a < mid) tmp[o++] = perm[a++];
    while (b < hi) tmp[o++] = perm[b++];
    memcpy(perm + lo, tmp + lo, (hi - lo) * sizeof(uint32_t));
}

si_oracle_index* si_oracle_build(const int32_t* starts, const int32_t* ends,
                                 const int32_t* data, size_t n) {
    si_oracle_index* ix = (si_oracle_index*)calloc(1, sizeof(*ix));
    ix->n = n;
    ix->start_sorted_in = 1;
    ix->end_sorted_in = 1;
    if (n == 0) return ix; /* hpp:113-115: empty build is a no-op */

    /* add() bookkeeping, hpp:96-101: start_sorted stays true while starts are
     * non-decreasing; end_sorted is cleared when an equal start brings a
     * larger end. Once start_sorted drops, end_sorted is no longer tracked. */
    for (size_t i = 1; i < n; ++i) {
        if (!ix->start_sorted_in) break;
        if (starts[i] < starts[i - 1]) ix->start_sorted_in = 0;
        else if (starts[i] == starts[i - 1] && ends[i] > ends[i - 1]) ix->end_sorted_in = 0;
    }

    uint32_t* perm = (uint32_t*)malloc(n * sizeof(uint32_t));
    for (size_t i = 0; i < n; ++i) perm[i] = (uint32_t)i;

    /* sort_intervals(), hpp:1415-1438: three cases. */
    if (!ix->start_sorted_in) {
        uint32_t* tmp = (uint32_t*)malloc(n * sizeof(uint32_t));
        merge_sort_perm(perm, tmp, 0, n, starts, ends);
        free(tmp);
    } else if (!ix->end_sorted_in) {
        /* only equal-start runs whose ends are not already descending */
        uint32_t* tmp = (uint32_t*)malloc(n * sizeof(uint32_t));
        size_t run = 0;
        while (run < n) {
            size_t stop = run + 1;
            int unordered = 0;
            while (stop < n && starts[stop] == starts[run]) {
                if (ends[stop] > ends[stop - 1]) unordered = 1;
                ++stop;
            }
            if (unordered) merge_sort_perm(perm, tmp, run, stop, starts, ends);
            run = stop;
        }
        free(tmp);
    } /* else: both flags true -> order is insertion order */

    ix->starts = (int32_t*)malloc(n * sizeof(int32_t));
    ix->ends = (int32_t*)malloc(n * sizeof(int32_t));
    ix->data = (int32_t*)malloc(n * sizeof(int32_t));
    ix->branch = (size_t*)malloc(n * sizeof(size_t));
    for (size_t i = 0; i < n; ++i) {
        ix->starts[i] = starts[perm[i]];
        ix->ends[i] = ends[perm[i]];
        ix->data[i] = data ? data[perm[i]] : (int32_t)perm[i];
    }
    free(perm);

    /* branch loop, hpp:117-129: monotonic stack of (end, idx); pop while the
     * top end is STRICTLY smaller (so equal ends chain, Q5). */
    size_t* stk = (size_t*)malloc(n * sizeof(size_t));
    size_t top = 0;
    for (size_t i = 0; i < n; ++i) {
        int32_t e = ix->ends[i];
        while (top > 0 && ix->ends[stk[top - 1]] < e) --top;
        ix->branch[i] = top ? stk[top - 1] : SI_ORACLE_NONE;
        stk[top++] = i;
    }
    free(stk);
    return ix;
}

void si_oracle_free(si_oracle_index* ix) {
    if (!ix) return;
    free(ix->starts); free(ix->ends); free(ix->data); free(ix->branch);
    free(ix);
}

/* hpp:501-513: halving search; returns last index with starts[idx] <= value,
 * SI_ORACLE_NONE when none (the --idx underflow of hpp:509-511). */
size_t si_oracle_upper_bound(const si_oracle_index* ix, int32_t value) {
    size_t len = ix->n, pos = 0;
    if (len == 0) return SI_ORACLE_NONE;
    while (len > 1) {
        size_t half = len / 2;
        if (ix->starts[pos + half] <= value) pos += len - half;
        len = half;
    }
    if (ix->starts[pos] > value) return pos == 0 ? SI_ORACLE_NONE : pos - 1;
    return pos;
}

/* The backward walk shared by every query (c.h:583-607, hpp:559-578):
 * test ends[i]; hit -> emit, step to i-1; miss -> jump to branch[i]. */
size_t si_oracle_search(const si_oracle_index* ix, int32_t qs, int32_t qe,
                        int32_t* values_out, uint32_t* idx_out, int32_t* keys_out) {
    if (ix->n == 0) return 0;
    size_t i = si_oracle_upper_bound(ix, qe);
    size_t k = 0;
    while (i != SI_ORACLE_NONE) {
        if (qs <= ix->ends[i]) {
            if (values_out) values_out[k] = ix->data[i];
            if (idx_out) idx_out[k] = (uint32_t)i;
            if (keys_out) { keys_out[2 * k] = ix->starts[i]; keys_out[2 * k + 1] = ix->ends[i]; }
            ++k;
            i = (i == 0) ? SI_ORACLE_NONE : i - 1;
        } else {
            i = ix->branch[i];
        }
    }
    return k;
}

size_t si_oracle_count(const si_oracle_index* ix, int32_t qs, int32_t qe) {
    return si_oracle_search(ix, qs, qe, NULL, NULL, NULL);
}

int si_oracle_has_overlaps(const si_oracle_index* ix, int32_t qs, int32_t qe) {
    if (ix->n == 0) return 0; /* hpp:866-868 */
    size_t i = si_oracle_upper_bound(ix, qe);
    return i != SI_ORACLE_NONE && qs <= ix->ends[i]; /* hpp:869-870, Q1 */
}

void si_oracle_count_batch(const si_oracle_index* ix, const int32_t* qs,
                           const int32_t* qe, size_t nq, uint64_t* counts) {
    for (size_t q = 0; q < nq; ++q) counts[q] = si_oracle_count(ix, qs[q], qe[q]);
}

void si_oracle_has_overlaps_batch(const si_oracle_index* ix, const int32_t* qs,
                                  const int32_t* qe, size_t nq, uint8_t* out) {
    for (size_t q = 0; q < nq; ++q) out[q] = (uint8_t)si_oracle_has_overlaps(ix, qs[q], qe[q]);
}

void si_oracle_search_batch(const si_oracle_index* ix, const int32_t* qs,
                            const int32_t* qe, size_t nq, const uint64_t* offsets,
                            int32_t* values_out, uint32_t* idx_out, int32_t* keys_out) {
    for (size_t q = 0; q < nq; ++q) {
        uint64_t o = offsets[q];
        si_oracle_search(ix, qs[q], qe[q],
                         values_out ? values_out + o : NULL,
                         idx_out ? idx_out + o : NULL,
                         keys_out ? keys_out + 2 * o : NULL);
    }
}

void si_oracle_walk_stats(const si_oracle_index* ix, const int32_t* qs,
                          const int32_t* qe, size_t nq,
                          uint64_t* hits_out, uint64_t* jumps_out) {
    uint64_t h = 0, j = 0;
    for (size_t q = 0; q < nq; ++q) {
        if (ix->n == 0) break;
        size_t i = si_oracle_upper_bound(ix, qe[q]);
        while (i != SI_ORACLE_NONE) {
            if (qs[q] <= ix->ends[i]) { ++h; i = (i == 0) ? SI_ORACLE_NONE : i - 1; }
            else { ++j; i = ix->branch[i]; }
        }
    }
    *hits_out = h;
    *jumps_out = j;
}
