/*
 * si_oracle_setops.c -- CPU ORACLE for the set-algebra callers of the query path
 * (test infrastructure, NOT product code; see si_oracle.h).
 *
 * Restates, over plain arrays, what the reference's C ABI computes in
 * c_superintervals.h:823-1064 (C++ twins: superintervals.hpp:1037-1390):
 * merge / gaps / union / intersection / difference / symmetric difference /
 * span / expand / flank / unique. Every routine cites the lines it follows.
 * Pinned by tests/test_setops_oracle.py against the UNMODIFIED reference C
 * header compiled in place (oracle/_ref/libsi_cref.so) and against the
 * fixtures in tests/golden/setops_*.npz made from it (tools/make_golden_setops.py).
 *
 * An input "set" is three parallel arrays in STORED order (as add() received
 * them, or position order once indexed). Outputs are growable lists in the
 * reference's emission order.
 */
#include "si_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---- output list ---------------------------------------------------------- */
static si_oracle_set* set_new(void) { return (si_oracle_set*)calloc(1, sizeof(si_oracle_set)); }

static void set_push(si_oracle_set* o, int32_t s, int32_t e, int32_t d) {
    if (o->n == o->cap) {
        o->cap = o->cap ? 2 * o->cap : 64;
        o->starts = (int32_t*)realloc(o->starts, o->cap * sizeof(int32_t));
        o->ends = (int32_t*)realloc(o->ends, o->cap * sizeof(int32_t));
        o->data = (int32_t*)realloc(o->data, o->cap * sizeof(int32_t));
    }
    o->starts[o->n] = s;
    o->ends[o->n] = e;
    o->data[o->n] = d;
    o->n++;
}

void si_oracle_set_free(si_oracle_set* o) {
    if (!o) return;
    free(o->starts);
    free(o->ends);
    free(o->data);
    free(o);
}

static int32_t first_of(int32_t a, int32_t b) { (void)b; return a; }

/* ---- stored order -> (start asc, end desc) visiting order --------------------
 * c.h:831-852: identity when the add()-time flags say the list is already in
 * that order (c.h:402-419), otherwise a qsort by (start asc, end desc)
 * (glibc's qsort is a stable merge sort, so ties keep stored order; the oracle's
 * own build uses the same stable order, si_oracle.c). The order is obtained by
 * building an oracle index over payload = stored position.                    */
static size_t* visiting_order(const int32_t* s, const int32_t* e, size_t n) {
    size_t* order = (size_t*)malloc((n ? n : 1) * sizeof(size_t));
    int32_t* pos = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
    for (size_t k = 0; k < n; ++k) pos[k] = (int32_t)k;
    si_oracle_index* ix = si_oracle_build(s, e, pos, n);
    for (size_t k = 0; k < n; ++k) order[k] = (size_t)ix->data[k];
    si_oracle_free(ix);
    free(pos);
    return order;
}

/* c.h:854-881 (hpp:1057-1087): sweep in visiting order; an interval that starts
 * at or before the running end joins the open cluster (end = max, data folded),
 * otherwise the cluster is flushed and a new one opens.                       */
si_oracle_set* si_oracle_merge(const int32_t* s, const int32_t* e, const int32_t* d, size_t n,
                               si_oracle_combine combine) {
    si_oracle_set* out = set_new();
    if (n == 0) return out;
    if (!combine) combine = first_of;
    size_t* order = visiting_order(s, e, n);
    int32_t open_s = s[order[0]], open_e = e[order[0]], open_d = d[order[0]];
    for (size_t k = 1; k < n; ++k) {
        const size_t i = order[k];
        if (s[i] > open_e) {
            set_push(out, open_s, open_e, open_d);
            open_s = s[i]; open_e = e[i]; open_d = d[i];
        } else {
            if (e[i] > open_e) open_e = e[i];
            open_d = combine(open_d, d[i]);
        }
    }
    set_push(out, open_s, open_e, open_d);
    free(order);
    return out;
}

/* c.h:883-905 (hpp:1104-1127): complement of the merged set inside [lo, hi]. */
si_oracle_set* si_oracle_gaps(const int32_t* s, const int32_t* e, const int32_t* d, size_t n,
                              int32_t lo, int32_t hi, int32_t fill) {
    si_oracle_set* out = set_new();
    si_oracle_set* m = si_oracle_merge(s, e, d, n, NULL);
    int64_t cursor = lo;   /* 64-bit: the reference's e + 1 overflows int32 only at e == INT32_MAX (UB there) */
    for (size_t k = 0; k < m->n; ++k) {
        if (m->ends[k] < lo || m->starts[k] > hi) continue;
        if ((int64_t)m->starts[k] > cursor) set_push(out, (int32_t)cursor, m->starts[k] - 1, fill);
        if ((int64_t)m->ends[k] + 1 > cursor) cursor = (int64_t)m->ends[k] + 1;
    }
    if (cursor <= (int64_t)hi) set_push(out, (int32_t)cursor, hi, fill);
    si_oracle_set_free(m);
    return out;
}

/* c.h:907-919 (hpp:1137-1147): concatenate (this first), then merge. */
si_oracle_set* si_oracle_union(const int32_t* s1, const int32_t* e1, const int32_t* d1, size_t n1,
                               const int32_t* s2, const int32_t* e2, const int32_t* d2, size_t n2,
                               si_oracle_combine combine) {
    const size_t n = n1 + n2;
    int32_t* s = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
    int32_t* e = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
    int32_t* d = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
    memcpy(s, s1, n1 * 4); memcpy(s + n1, s2, n2 * 4);
    memcpy(e, e1, n1 * 4); memcpy(e + n1, e2, n2 * 4);
    memcpy(d, d1, n1 * 4); memcpy(d + n1, d2, n2 * 4);
    si_oracle_set* out = si_oracle_merge(s, e, d, n, combine);
    free(s); free(e); free(d);
    return out;
}

/* c.h:921-939 (hpp:1164-1179): every stored interval of A queries B's index
 * (searchIdxs: hits in descending position); each hit contributes the clipped
 * piece when it is non-empty; data = combine(a, b).                           */
si_oracle_set* si_oracle_intersection(const int32_t* s, const int32_t* e, const int32_t* d, size_t n,
                                      const si_oracle_index* other, si_oracle_combine combine) {
    si_oracle_set* out = set_new();
    if (!combine) combine = first_of;
    uint32_t* hits = (uint32_t*)malloc((other->n ? other->n : 1) * sizeof(uint32_t));
    for (size_t k = 0; k < n; ++k) {
        const size_t h = si_oracle_search(other, s[k], e[k], NULL, hits, NULL);
        for (size_t t = 0; t < h; ++t) {
            const uint32_t j = hits[t];
            const int32_t ps = s[k] > other->starts[j] ? s[k] : other->starts[j];
            const int32_t pe = e[k] < other->ends[j] ? e[k] : other->ends[j];
            if (ps <= pe) set_push(out, ps, pe, combine(d[k], other->data[j]));
        }
    }
    free(hits);
    return out;
}

/* (start asc, end asc), c.h:941-948 */
static int pair_cmp(const void* a, const void* b) {
    const int32_t* x = (const int32_t*)a;
    const int32_t* y = (const int32_t*)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    if (x[1] != y[1]) return x[1] < y[1] ? -1 : 1;
    return 0;
}

/* c.h:950-974 (hpp:1189-1211): for every stored interval of A, the (start,end)
 * pairs of B that overlap it, visited by (start asc, end asc); a cursor sweeps
 * the interval and the stretches no pair covers are emitted with A's data.    */
si_oracle_set* si_oracle_difference(const int32_t* s, const int32_t* e, const int32_t* d, size_t n,
                                    const si_oracle_index* other) {
    si_oracle_set* out = set_new();
    int32_t* cover = (int32_t*)malloc((other->n ? other->n : 1) * 2 * sizeof(int32_t));
    for (size_t k = 0; k < n; ++k) {
        const size_t h = si_oracle_search(other, s[k], e[k], NULL, NULL, cover);
        qsort(cover, h, 2 * sizeof(int32_t), pair_cmp);   /* equal pairs are indistinguishable: stability is moot */
        int64_t cursor = s[k];
        for (size_t t = 0; t < h; ++t) {
            const int32_t cs = cover[2 * t] > s[k] ? cover[2 * t] : s[k];
            const int32_t ce = cover[2 * t + 1] < e[k] ? cover[2 * t + 1] : e[k];
            if ((int64_t)cs > cursor) set_push(out, (int32_t)cursor, cs - 1, d[k]);
            if ((int64_t)ce + 1 > cursor) cursor = (int64_t)ce + 1;
        }
        if (cursor <= (int64_t)e[k]) set_push(out, (int32_t)cursor, e[k], d[k]);
    }
    free(cover);
    return out;
}

/* c.h:976-983 (hpp:1219-1223): (A \ B) united with (B \ A), first data kept.
 * Both inputs are the indexed sets (the reference requires both to be indexed),
 * i.e. position order. */
si_oracle_set* si_oracle_symmetric_difference(const si_oracle_index* a, const si_oracle_index* b) {
    si_oracle_set* ab = si_oracle_difference(a->starts, a->ends, a->data, a->n, b);
    si_oracle_set* ba = si_oracle_difference(b->starts, b->ends, b->data, b->n, a);
    si_oracle_set* out = si_oracle_union(ab->starts, ab->ends, ab->data, ab->n, ba->starts, ba->ends, ba->data, ba->n, NULL);
    si_oracle_set_free(ab);
    si_oracle_set_free(ba);
    return out;
}

/* c.h:985-998 (hpp:1230-1243) */
int si_oracle_span(const int32_t* s, const int32_t* e, size_t n, int32_t* lo, int32_t* hi) {
    if (n == 0) return 0;
    int32_t a = s[0], b = e[0];
    for (size_t k = 1; k < n; ++k) {
        if (s[k] < a) a = s[k];
        if (e[k] > b) b = e[k];
    }
    *lo = a;
    *hi = b;
    return 1;
}

/* c.h:1000-1016: start - left, end + right in 64-bit, clamped to [lo, hi]; an
 * interval shrunk past itself is dropped; stored order kept. */
si_oracle_set* si_oracle_expand(const int32_t* s, const int32_t* e, const int32_t* d, size_t n,
                                int32_t left, int32_t right, int32_t lo, int32_t hi) {
    si_oracle_set* out = set_new();
    for (size_t k = 0; k < n; ++k) {
        int64_t a = (int64_t)s[k] - left, b = (int64_t)e[k] + right;
        if (a < lo) a = lo;
        if (b > hi) b = hi;
        if (a <= b) set_push(out, (int32_t)a, (int32_t)b, d[k]);
    }
    return out;
}

/* c.h:1018-1040: the strip of width `left` before each start and of width `right`
 * after each end, clamped to [lo, hi]; left flank first; originals not emitted. */
si_oracle_set* si_oracle_flank(const int32_t* s, const int32_t* e, const int32_t* d, size_t n,
                               int32_t left, int32_t right, int32_t lo, int32_t hi) {
    si_oracle_set* out = set_new();
    for (size_t k = 0; k < n; ++k) {
        if (left > 0 && s[k] > lo) {
            const int32_t le = s[k] - 1;
            int64_t ls = (int64_t)s[k] - left;
            if (ls < lo) ls = lo;
            if (ls <= le && le <= hi) set_push(out, (int32_t)ls, le, d[k]);
        }
        if (right > 0 && e[k] < hi) {
            const int32_t rs = e[k] + 1;
            int64_t re = (int64_t)e[k] + right;
            if (re > hi) re = hi;
            if (rs <= re && rs >= lo) set_push(out, rs, (int32_t)re, d[k]);
        }
    }
    return out;
}

/* c.h:1042-1064 (hpp:1341-1360): visiting order groups exact (start, end)
 * duplicates; one interval per group, data folded left to right. */
si_oracle_set* si_oracle_unique(const int32_t* s, const int32_t* e, const int32_t* d, size_t n,
                                si_oracle_combine combine) {
    si_oracle_set* out = set_new();
    if (n == 0) return out;
    if (!combine) combine = first_of;
    size_t* order = visiting_order(s, e, n);
    size_t head = order[0];
    int32_t acc = d[head];
    for (size_t k = 1; k < n; ++k) {
        const size_t i = order[k];
        if (s[i] == s[head] && e[i] == e[head]) {
            acc = combine(acc, d[i]);
        } else {
            set_push(out, s[head], e[head], acc);
            head = i;
            acc = d[i];
        }
    }
    set_push(out, s[head], e[head], acc);
    free(order);
    return out;
}
