/*
 * c_superintervals.h -- the SuperIntervals C ABI, backed by libsuperintervals_b200.so
 * (CUDA, sm_100a). DECLARATIONS ONLY.
 *
 * This is the drop-in boundary: struct layouts, names, argument meaning and
 * (absent) error behaviour are those of the reference header
 * reference/src/c_superintervals.h (declarations at :61-357, cited per entry
 * below as ref:LINE). The reference header also *defines* these functions
 * (ref:363-1089); a caller switches implementation by including this header
 * instead and linking -lsuperintervals_b200. Nothing here mentions CUDA.
 *
 * Scope: build + overlap queries, and the callers either
 * side of them: the reference's set algebra (ref:264-332: mergeOverlaps,
 * intervalGaps, unionWith, intersection, difference, symmetricDifference,
 * intervalSpan, expandIntervals, flankIntervals, uniqueIntervals) is declared
 * below and runs on the device as well.
 *
 * Semantics kept from the reference:
 *   - intervals and queries are END-INCLUSIVE;
 *   - indexSuperIntervals() must run after adding and before querying;
 *   - results come in DESCENDING position order (the backward walk's order);
 *   - search* APPEND to the result buffer (ref:209); clear it for fresh results;
 *   - anyOverlaps tests only the last candidate (ref:570-573), a known
 *     false-negative on nested data that callers may depend on;
 *   - no function reports errors; CUDA failures are latched and readable through
 *     si_b200_last_error() in superintervals_b200.h.
 * Every query -- including the single-query forms -- executes on the GPU; there
 * is no CPU fallback. Use the *Batch entry points (superintervals_b200.h) for
 * throughput.
 */
#ifndef SUPERINTERVALS_B200_C_ABI_INCLUDED
#define SUPERINTERVALS_B200_C_ABI_INCLUDED 1

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SUPERINTERVALS_C_VERSION "1.0.0"   /* ref:42 */
#define SI_NONE ((size_t)-1)               /* ref:56 */

typedef struct { int32_t start; int32_t end; int32_t data; } Interval;   /* ref:61-65 */
typedef struct { int32_t start; int32_t end; } KeyPair;                  /* ref:70-73 */

/* ref:81-91 -- same field order and sizes (64 bytes on LP64). After
 * indexSuperIntervals() the arrays are host mirrors of the device index. */
typedef struct {
    int32_t* starts;
    int32_t* ends;
    int32_t* data;
    size_t*  branch;
    size_t   size;
    size_t   capacity;
    size_t   idx;          /* cursor written by upperBound (ref:562-566) */
    bool     startSorted;
    bool     endSorted;
} cSuperIntervals;

typedef struct { int32_t*  data; size_t size; size_t capacity; } cIndexResult;   /* ref:98-102 */
typedef struct { KeyPair*  data; size_t size; size_t capacity; } cKeyResult;     /* ref:107-111 */
typedef struct { Interval* data; size_t size; size_t capacity; } cItemResult;    /* ref:116-120 */

/* lifecycle -- ref:135-156 */
cSuperIntervals* createSuperIntervals(void);
void   destroySuperIntervals(cSuperIntervals* si);
void   clearSuperIntervals(cSuperIntervals* si);
void   reserveSuperIntervals(cSuperIntervals* si, size_t n);
void   addInterval(cSuperIntervals* si, int32_t start, int32_t end, int32_t value);
size_t sizeSuperIntervals(const cSuperIntervals* si);

/* indexing -- ref:163,169. Device radix sort + parallel branch pass; host mirrors refreshed. */
void sortIntervals(cSuperIntervals* si);
void indexSuperIntervals(cSuperIntervals* si);

/* element access -- ref:176-185 */
bool    intervalAt(const cSuperIntervals* si, size_t index, Interval* out);
int32_t startAt(const cSuperIntervals* si, size_t index);
int32_t endAt(const cSuperIntervals* si, size_t index);
int32_t dataAt(const cSuperIntervals* si, size_t index);

/* queries -- ref:195-240 */
size_t upperBound(cSuperIntervals* si, int32_t value);
bool   anyOverlaps(cSuperIntervals* si, int32_t start, int32_t end);
size_t countOverlaps(cSuperIntervals* si, int32_t start, int32_t end);
void   searchValues(cSuperIntervals* si, int32_t start, int32_t end, cIndexResult* found);
void   searchIdxs(cSuperIntervals* si, int32_t start, int32_t end, cIndexResult* found);
void   searchKeys(cSuperIntervals* si, int32_t start, int32_t end, cKeyResult* found);
void   searchItems(cSuperIntervals* si, int32_t start, int32_t end, cItemResult* found);
void   searchPoint(cSuperIntervals* si, int32_t point, cIndexResult* found);
void   coverage(cSuperIntervals* si, int32_t start, int32_t end, size_t* count_out, int32_t* coverage_out);
void   findOverlaps(cSuperIntervals* si, int32_t start, int32_t end, int32_t* found, size_t* found_size);

/* set operations -- ref:243-324 (definitions ref:823-1064). Each returns a NEW, NOT-yet-indexed
 * set (owning pointer; free with destroySuperIntervals) in the reference's emission order.
 * Intervals are end-inclusive; merging coalesces overlapping intervals only ([1,5],[6,10] stay
 * apart). `combine` resolves the data when intervals fold together (NULL keeps the first/left).
 * intersection / difference / symmetricDifference query `other` through its index: it must be
 * indexed (symmetricDifference: both). Here: device sort + head flags + scan + scatter
 * (merge / unique / gaps), A's stored intervals as one query batch against B's index
 * (intersection / difference), elementwise count + scan + scatter (expand / flank).
 * Like the queries, an operation that queries `other` uses that handle's device staging: do
 * not run two of them on the same `other` concurrently. */
typedef int32_t (*cCombineFn)(int32_t, int32_t);                                           /* ref:128 */
cSuperIntervals* mergeOverlaps(const cSuperIntervals* si, cCombineFn combine);            /* ref:264 */
cSuperIntervals* intervalGaps(const cSuperIntervals* si, int32_t lo, int32_t hi, int32_t fill);   /* ref:271 */
cSuperIntervals* unionWith(const cSuperIntervals* si, const cSuperIntervals* other, cCombineFn combine);   /* ref:278 */
cSuperIntervals* intersection(const cSuperIntervals* si, cSuperIntervals* other, cCombineFn combine);      /* ref:286 */
cSuperIntervals* difference(const cSuperIntervals* si, cSuperIntervals* other);           /* ref:292 */
cSuperIntervals* symmetricDifference(cSuperIntervals* si, cSuperIntervals* other);        /* ref:298 */
bool intervalSpan(const cSuperIntervals* si, int32_t* lo_out, int32_t* hi_out);           /* ref:305 */
cSuperIntervals* expandIntervals(const cSuperIntervals* si, int32_t left, int32_t right, int32_t lo, int32_t hi);   /* ref:314 */
cSuperIntervals* flankIntervals(const cSuperIntervals* si, int32_t left, int32_t right, int32_t lo, int32_t hi);    /* ref:324 */
cSuperIntervals* uniqueIntervals(const cSuperIntervals* si, cCombineFn combine);          /* ref:332 */

/* result buffers -- ref:339-357 */
cIndexResult createIndexResult(void);
void clearIndexResult(cIndexResult* r);
void destroyIndexResult(cIndexResult* r);
cKeyResult createKeyResult(void);
void clearKeyResult(cKeyResult* r);
void destroyKeyResult(cKeyResult* r);
cItemResult createItemResult(void);
void clearItemResult(cItemResult* r);
void destroyItemResult(cItemResult* r);

#ifdef __cplusplus
}
#endif
#endif /* SUPERINTERVALS_B200_C_ABI_INCLUDED */
