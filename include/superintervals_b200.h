/*
 * superintervals_b200.h -- ADDITIVE entry points of libsuperintervals_b200.so.
 *
 * The reference C ABI (c_superintervals.h) is one-query-per-call and has no
 * error channel. This header adds, without touching any reference signature:
 *   1. a sticky error side channel for CUDA failures;
 *   2. batch queries over HOST buffers, shaped after the only batch API the
 *      reference defines -- the Python methods count_batch / search_idxs_batch /
 *      search_values_batch (reference src/superintervals/intervalmap.pyx:363-494),
 *      with their ragged list-of-lists flattened to CSR (offsets[n+1] + flat data);
 *   3. a device-resident core (opaque siIndex) taking raw DEVICE pointers and a
 *      stream, for callers that already hold queries in HBM and for multi-GPU
 *      sharding. No torch / CUDA types appear in any signature: streams travel as
 *      void* (a cudaStream_t), device arrays as plain pointers.
 *
 * All functions returning int return 0 on success or a cudaError_t value (also
 * latched into the side channel).
 */
#ifndef SUPERINTERVALS_B200_H_INCLUDED
#define SUPERINTERVALS_B200_H_INCLUDED 1

#include "c_superintervals.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SUPERINTERVALS_B200_VERSION "0.1.0"

/* ---- 1. error side channel (SURVEY 8b "Error convention") ------------------------ */
int         si_b200_last_error(void);          /* 0 = no error since the last clear */
const char* si_b200_last_error_string(void);   /* static buffer, "" when no error */
void        si_b200_clear_error(void);
const char* si_b200_version(void);
int         si_b200_device_count(void);        /* visible CUDA devices, <0 on error */

/* ---- 2. batch queries over host buffers ------------------------------------------- */
/* Bulk add: n x addInterval (ref c_superintervals.h:402-419) in one call, same
 * startSorted/endSorted bookkeeping. values == NULL stores the insertion index
 * (what reference test/bench.cpp:210 passes). Shaped after from_arrays (pyx:63-131). */
void addIntervals(cSuperIntervals* si, const int32_t* starts, const int32_t* ends,
                  const int32_t* values, size_t n);

/* When false, indexSuperIntervals() leaves si->starts/ends/data/branch untouched
 * (insertion order, branch NULL) and skips the device->host mirror copy; only
 * the *Batch calls remain meaningful for element access. Default: true. */
void siSetHostMirror(cSuperIntervals* si, bool enabled);

/* count_batch (pyx:363-400): counts_out[i] = countOverlaps(si, starts[i], ends[i]). */
void countOverlapsBatch(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n,
                        size_t* counts_out);
/* The same with 32-bit counts: a count never exceeds the number of stored intervals (< 2^32), and the
 * copy back to the host is half as long -- the D2H half is what bounds the end-to-end rate of count. */
void countOverlapsBatch32(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n,
                          uint32_t* counts_out);
/* out[i] = anyOverlaps(si, starts[i], ends[i]) -- same last-candidate-only test. */
void anyOverlapsBatch(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n,
                      bool* out);
/* search_values_batch (pyx:448-494) as CSR. offsets_out has n+1 entries and is
 * relative to the values APPENDED by this call: query i's results are
 * found->data[size_before + offsets_out[i] .. size_before + offsets_out[i+1]),
 * each list in the reference's descending order. found grows by realloc. */
void searchValuesBatch(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n,
                       size_t* offsets_out, cIndexResult* found);
/* search_idxs_batch (pyx:402-446) as CSR; all-descending order as the reference C
 * ABI's searchIdxs (c_superintervals.h:610-643). */
void searchIdxsBatch(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n,
                     size_t* offsets_out, cIndexResult* found);
void searchKeysBatch(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n,
                     size_t* offsets_out, cKeyResult* found);
void searchItemsBatch(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n,
                      size_t* offsets_out, cItemResult* found);
/* coverage (c_superintervals.h:758-792) per query. */
void coverageBatch(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n,
                   size_t* count_out, int32_t* coverage_out);
/* intersection() (c_superintervals.h:286) without a callback: the pieces carry `si`'s data, and
 * the `other` side's data of every piece is appended to other_data in the same order, so a
 * host-language binding can pair or combine payloads itself (intervalmap.pyx:577-609 pairs
 * them into tuples). */
cSuperIntervals* intersectionPairs(const cSuperIntervals* si, cSuperIntervals* other, cIndexResult* other_data);

/* ---- 2b. BED ingest: the step before the path ------------------------------------------
 * Tokenises BED text on the device (csrc/bed.cu): one record per line, tab-separated
 * chrom, start, end (further columns ignored), numbers read with std::stoi's rules -- what
 * the reference's callers do one line at a time on the host (test/bench.cpp:67-102,
 * examples/bed-intersect-si.rs:63-123). Lines without a chrom, a numeric start and a numeric
 * end (headers, comments, blanks; the reference would throw) are skipped and counted.
 * normalize != 0 swaps start/end when start > end (bench.cpp:89); end_shift is added to every
 * end (-1 turns BED's half-open ends into the inclusive ends the index stores, bench.cpp:210).
 * contig[i] indexes names[], which lists the distinct chroms in order of first appearance.
 * group_by_contig != 0 reorders the records by contig (stable: line order inside a contig) and fills
 * contig_offsets[0..n_contigs]: contig k owns records [contig_offsets[k], contig_offsets[k+1]) --
 * the per-chrom containers of bed-intersect-si.rs:100-123, ready for one index each.
 * All arrays are malloc'd by the library; release with siBedTableFree. Returns 0 or a CUDA error. */
typedef struct {
    int32_t* contig;
    int32_t* starts;
    int32_t* ends;
    size_t n;          /* records parsed (line order) */
    size_t lines;      /* lines seen */
    size_t skipped;    /* lines - n */
    char** names;
    size_t n_contigs;
    size_t* contig_offsets;   /* n_contigs + 1 entries when grouped, else NULL */
} siBedTable;
int siParseBed(const char* text, size_t bytes, int normalize, int end_shift, int group_by_contig, siBedTable* out);
void siBedTableFree(siBedTable* t);

/* ---- 3. device-resident core ------------------------------------------------------- */
typedef struct siIndex siIndex;

typedef struct {
    const int32_t*  starts;   /* device, position order (start asc, end desc) */
    const int32_t*  ends;     /* device, padded to a multiple of 128 entries */
    const int32_t*  values;   /* device */
    const uint32_t* branch;   /* device, 0xFFFFFFFF = none */
    size_t          n;
    int             device;
} siDeviceView;

/* how the query arrays are ordered; results are always returned in the caller's order */
enum {
    SI_ORDER_AUTO = 0,       /* check on device (one small host sync), sort if needed */
    SI_ORDER_SORTED = 1,     /* caller guarantees query STARTS are non-decreasing (position-sorted, as `bedtools sort`) */
    SI_ORDER_UNSORTED = 2,   /* partition the batch by position on device first; results still go to the caller's slots */
    SI_ORDER_ASIS = 3        /* process in the given order whatever it is */
};
enum { SI_FILL_VALUES = 0, SI_FILL_IDXS = 1, SI_FILL_KEYS = 2, SI_FILL_ITEMS = 3 };

siIndex* siIndexCreate(void);            /* on the calling thread's current CUDA device */
void     siIndexDestroy(siIndex* ix);
siIndex* siIndexOf(cSuperIntervals* si); /* the device index behind a handle (NULL before indexing) */
size_t   siIndexSize(const siIndex* ix);
int      siIndexDeviceView(const siIndex* ix, siDeviceView* out);

/* build(): radix sort by (start asc, end desc, insertion order) + parallel branch pass.
 * values may be NULL (payload = insertion index). Inputs are not modified. */
int siIndexBuildHost(siIndex* ix, const int32_t* starts, const int32_t* ends, const int32_t* values, size_t n);
int siIndexBuildDevice(siIndex* ix, const int32_t* d_starts, const int32_t* d_ends,
                       const int32_t* d_values, size_t n, void* stream);
/* copy the built index to host; any pointer may be NULL. branch is widened to
 * size_t with SI_NONE; perm[i] = insertion index of the interval at position i. */
int siIndexExport(const siIndex* ix, int32_t* starts, int32_t* ends, int32_t* values,
                  size_t* branch, uint32_t* perm);

/* d_counts[i] = number of stored intervals overlapping [d_qs[i], d_qe[i]]. */
int siCountDevice(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n,
                  uint32_t* d_counts, int order, void* stream);
int siCountDevice64(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n,
                    uint64_t* d_counts, int order, void* stream);
/* Optional explicit first half of an SI_ORDER_UNSORTED count: partition the batch by
 * position now (records of start, end, index grouped by result window and position
 * bucket); the NEXT siCountDevice(.., SI_ORDER_UNSORTED, ..) on the same (d_qs, d_qe, n)
 * consumes that partition instead of redoing it (one-shot). Lets callers overlap or time
 * the two phases separately. No-op for batches above 2^27 queries (those are sliced). */
int siSortQueriesDevice(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n, void* stream);

/* Tunables of one index. SI_OPT_COUNT_ALGO: which count kernel answers batches --
 * SI_COUNT_WALK is the tile sweep + branch-array walk (the reference's algorithm,
 * hpp:651-825, on any index); SI_COUNT_RANK is the closed form
 * #{starts <= qe} - #{ends < qs} (valid on an index whose intervals all have
 * start <= end; queries with qs > qe inside such a batch still take the walk);
 * SI_COUNT_CELLS is the same closed form answered from build()'s rank cells (one
 * 32-byte record per 2^k coordinates: one sector read per rank, no locality needed, so a
 * shuffled batch is answered without partitioning it, from L2 while the cells fit there, else from HBM);
 * SI_COUNT_AUTO picks CELLS, else RANK, whenever the index allows it. Results are identical. */
enum { SI_OPT_COUNT_ALGO = 0, SI_OPT_BUCKET_INTERVALS = 1, SI_OPT_WINDOW_SHIFT = 2, SI_OPT_TIMING = 3,
       SI_OPT_GRID_INTERVALS = 4, /* intervals per cell of the rank grid; applies to the next build */
       SI_OPT_CELLS_DIRECT_BYTES = 5, /* rank cells up to this many bytes answer batches unpartitioned (0 = always, the default: the gather from
                                         HBM beats partitioning first at every size measured; 1 = never) */
       SI_OPT_CELLS_FILL = 6, /* mean values per rank cell (1..28); applies to the next build */
       SI_OPT_STAB_LISTS = 7, /* 1 (default): the CSR fill of a well-formed index reads its stab lists; 0: it walks the branch array */
       SI_OPT_STAB_BUDGET = 8, /* stab-list entries per interval at most (default 6); deeper indexes double the checkpoint spacing, then walk */
       SI_OPT_STREAM = 9, /* 1 (default): position-sorted batches are counted by the streaming kernel (TMA-staged rank bits) when the
                             index carries them; 0: never; 2: every batch, whatever its order (a tile whose window does not fit reads the cells) */
       SI_OPT_L2_PERSIST = 11, /* 1: the rank-cells count is launched with an L2 access-policy window that keeps the cells resident
                                   (persisting) and lets the query / count streams pass (streaming); raises the device-wide persisting
                                   set-aside to its maximum. 0 (default): plain launch -- measured on C2: the window gains the count 1 %
                                   and the set-aside costs every other kernel of the process 15-45 % (stream count, scan, fill) */
       SI_OPT_NARROW_SORT = 12, /* 1 (default): build() sorts an unsorted input by start only (32-bit keys: half the passes, two thirds of
                                    the bytes per pass) and puts equal starts in end-descending order in place; a run of more than 16
                                    equal starts falls back to the composite (start, end desc) 64-bit key. 0: always the composite key.
                                    The built arrays are identical either way. */
       SI_OPT_RESIDENT_QUERIES = 13, /* 1 (default): the single-query C calls (countOverlaps, searchValues, upperBound ...) are answered by ONE
                                    resident warp that polls the handle's mapped pinned mailbox (~5 us per call) instead of a kernel launch
                                    per call (~11.5 us): for callers that loop over single queries, as every reference driver does. The
                                    kernel leaves by itself after 0.2 ms without a request and at most 2 ms after its launch (a device-wide
                                    synchronise elsewhere in the process never waits longer), is stopped before a rebuild, and is relaunched
                                    on demand. 0 (or SIB_RESIDENT_QUERIES=0 in the environment): one launch per call */
       SI_OPT_STAB_VALUE_LISTS = 14, /* 1 (default): short stab lists (8-byte records) are kept twice, as (position, end) and as (value, end),
                                    so that search_values reads a hit's payload with its record instead of gathering it; 0: positions only */
       SI_OPT_PAIR_CELLS = 15, /* pair cells: both ranks of a coordinate cell in ONE 32-byte record, so that a stabbing or short query
                                    gathers one sector instead of two. 1 (default): built when the rank cells exceed 3/4 of L2 (the gather
                                    is then served by HBM, which charges per sector); 0: never; 2: always. Applies to the next build */
       SI_OPT_STREAM_BUDGET = 10 /* rank bits are built when they cost at most this many bytes per interval (default 64; 0 = never); next build */ };
enum { SI_COUNT_AUTO = 0, SI_COUNT_WALK = 1, SI_COUNT_RANK = 2, SI_COUNT_CELLS = 3 };
int siIndexSetOption(siIndex* ix, int option, long long value);
/* What the last build did with its input: 0 = already in (start asc, end desc) order, no sort (hpp:1416,1421);
 * 1 = narrow sort + tie fix; 2 = composite 64-bit key; -1 = not built. */
int siIndexLastSort(const siIndex* ix);
/* The rank cells build() made (which = 0: over starts, 1: over ends). format 0 = none
 * (malformed index or >= 2^31 intervals), 1 = 28 one-byte offsets, 2 = 14 two-byte offsets per
 * 32-byte cell of 2^shift coordinates; overfull = cells answered from the sorted array instead;
 * direct = 1 when count answers unpartitioned batches straight from the cells.
 * which = 2: the pair cells (SI_OPT_PAIR_CELLS); format 0 = not built, 4 = 24 four-bit offsets per side in cells of
 * 16 coordinates, 8 = 12 one-byte offsets per side; overfull = sides answered from the sorted arrays. */
typedef struct {
    unsigned format, shift;
    unsigned long long cells, bytes, overfull;
    int direct;
} siCellsInfo;
int siIndexCellsInfo(const siIndex* ix, int which, siCellsInfo* out);
/* The rank bits build() made for the streaming count (csrc/stream_kernels.cuh): built = 1 when position-sorted
 * batches stream; words / bytes over both tables; slow_words = words holding a coordinate with three or more
 * values (answered from the rank cells). */
typedef struct {
    int built;
    unsigned long long words, bytes, slow_words;
} siBitsInfo;
int siIndexBitsInfo(const siIndex* ix, siBitsInfo* out);
/* The last streaming count (its last slice of 2^27 queries): tiles of 2048 queries launched, and how many of them the
 * streaming kernel handed back to the rank-cells code (window too wide for the staging buffers, or a query with
 * start > end inside). Synchronises the device. */
int siIndexStreamStats(siIndex* ix, unsigned long long* tiles, unsigned long long* handed_back);
/* The stab lists of the CSR fill (made by the first siFillDevice after a build): state 0 = not
 * made yet, 1 = in use, 2 = over budget or too small (the fill walks); a checkpoint every
 * 2^shift positions, `lists` lists holding `entries` records of `record_bytes` bytes. */
typedef struct {
    int state;
    unsigned shift;
    unsigned long long lists, entries;
    unsigned record_bytes;   /* 8: (position, end); 16: (position, end, value, -) on dense data */
} siStabInfo;
int siIndexStabInfo(const siIndex* ix, siStabInfo* out);
/* With SI_OPT_TIMING = 1 every hot kernel launch is bracketed by a CUDA event pair on its
 * own stream. Returns the number of (tag, milliseconds) records written (oldest first) and
 * clears them; waits for the recorded work. Tags: 1 partition histogram + scan, 2 partition
 * pass, 3 count (walk), 4 count (rank), 5 CSR scan, 6 CSR fill (walk), 7 count (cells), 8 CSR fill (runs + stab lists). bench.py's roofline uses it. */
enum { SI_TAG_PT_HIST = 1, SI_TAG_PT_PASS = 2, SI_TAG_COUNT_WALK = 3, SI_TAG_COUNT_RANK = 4, SI_TAG_SCAN = 5, SI_TAG_FILL = 6,
       SI_TAG_COUNT_CELLS = 7, SI_TAG_FILL_RUNS = 8, SI_TAG_COUNT_STREAM = 9 };
int siIndexReadTimings(siIndex* ix, int* tags, float* ms, int max_out);

int siAnyDevice(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n,
                uint8_t* d_out, void* stream);
/* d_offsets[0..n]: exclusive scan of d_counts, d_offsets[n] = total hits. 16-byte aligned pointers. */
int siScanDevice(siIndex* ix, const uint32_t* d_counts, size_t n, uint64_t* d_offsets, void* stream);
/* CSR fill: for query i writes its hits to d_out[d_offsets[i] .. d_offsets[i+1]) in
 * descending position order. what = SI_FILL_*: int32 values, uint32 positions,
 * KeyPair, or Interval records. With SI_ORDER_UNSORTED the partition made by the
 * preceding siCountDevice call on the same (d_qs, d_qe, n) is reused (do not modify the batch in between). */
int siFillDevice(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n,
                 const uint64_t* d_offsets, int what, void* d_out, int order, void* stream);
/* count + clipped-length sum per query (c_superintervals.h:758-792). */
int siCoverageDevice(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n,
                     uint32_t* d_counts, int32_t* d_cov, void* stream);

/* ---- 3b. mode B routing: one index per contig (the reference's callers' convention, examples/bed-intersect-si.rs:100-123)
 * A mixed batch (contig id, start, end per query) is grouped by contig ON THE DEVICE: d_perm[i] = original index of the i-th
 * query in contig-major stable order, d_qs_out / d_qe_out = the queries in that order, offsets_out (HOST, n_contigs + 1
 * entries) = where each contig's queries begin. Contig k's queries are then one contiguous device range for
 * siCountDevice on contig k's index; siScatterCountsDevice puts the routed counts back in the caller's order
 * (d_out[d_perm[i]] = d_counts[i]). `ix` lends its sort scratch (any index on the device; it is not modified).
 * Ids must lie in [0, n_contigs). Synchronises `stream` once (the offsets). */
int siRouteByContigDevice(siIndex* ix, const int32_t* d_contig, const int32_t* d_qs, const int32_t* d_qe, size_t n,
                          int n_contigs, int32_t* d_qs_out, int32_t* d_qe_out, uint32_t* d_perm, size_t* offsets_out,
                          void* stream);
int siScatterCountsDevice(siIndex* ix, const uint32_t* d_counts, const uint32_t* d_perm, size_t n, uint32_t* d_out, void* stream);
/* The same batch WITHOUT routing: ixs[k] = the index of contig k (NULL or empty: its queries count 0; all on one device),
 * d_counts[i] = count of query i on the index of contig d_contig[i] (ids outside [0, n_contigs) count 0), in the caller's
 * order, one launch; d_totals (device, n_contigs entries, may be NULL) receives the hit total of every contig -- what the
 * CSR bases of a contig-major result are made of. Needs every index to answer from rank cells (any well-formed index, and one with a few start > end
 * intervals); otherwise returns SI_MIXED_UNSUPPORTED without latching an error and the caller routes instead. */
#define SI_MIXED_UNSUPPORTED (-2)
int siCountMixedDevice(siIndex* const* ixs, int n_contigs, const int32_t* d_contig, const int32_t* d_qs, const int32_t* d_qe,
                       size_t n, uint32_t* d_counts, unsigned long long* d_totals, void* stream);
/* Mode B across GPUs WITHOUT a dispatch (the reference keeps one map per contig, examples/bed-intersect-si.rs:100-123).
 * The contigs are partitioned over the GPUs and every GPU's slice of the mixed batch stays where it is, in buffers its peers have
 * mapped (cudaDeviceEnablePeerAccess inside one process, or siIpcAlloc / siIpcOpen between processes). Each GPU calls this with
 * the same n_src slices (slice k lives on GPU k; home = this GPU's own position) and ITS table: ixs[c] = its index of contig c or
 * NULL, foreign[c] != 0 where another GPU holds contig c's index. The GPU walks all slices -- the remote ones are read in place
 * over NVLink -- answers exactly the queries of the contigs it indexes, and stores every count into the slice's own counts array
 * (a peer store for a remote slice). Queries of a contig nobody indexes, or with an id outside [0, n_contigs), get their 0 from
 * the slice's home GPU. No all-to-all, no routing sort, no scatter back: per query that crosses GPUs 4 B of contig id are read by
 * every GPU, 8 B of coordinates are read and 4 B of count written by the owner. d_totals (n_contigs entries, may be NULL) receives
 * the hit totals of this GPU's contigs over the whole batch. The caller puts a barrier between the GPUs on both sides
 * (siPeerBarrierDevice): slices complete before peers read them, all stores landed before counts are used. Returns
 * SI_MIXED_UNSUPPORTED as siCountMixedDevice does, and for a GPU that holds no index at all while other GPUs do. */
int siCountMixedPeerDevice(siIndex* const* ixs, int n_contigs, const unsigned char* foreign, int n_src, int home,
                           const int32_t* const* d_contig, const int32_t* const* d_qs, const int32_t* const* d_qe, const size_t* n,
                           uint32_t* const* d_counts, unsigned long long* d_totals, void* stream);

/* ---- 3c. count fused with the all-gather of the counts ---------------------------------------------
 * siCountDevice that additionally stores every count at the same index of up to 15 further arrays. With the arrays
 * being the OTHER GPUs' copies of a gathered count vector (peer memory: cudaDeviceEnablePeerAccess inside one process,
 * or a CUDA IPC mapping of another process's buffer, below), "count, then ncclAllGather of the counts" becomes one
 * kernel whose results travel over NVLink as peer stores while it is still ranking: d_counts = this GPU's slot of its
 * own copy, peers[k] = the same slot in GPU k's copy. The caller orders the GPUs afterwards (event waits inside one
 * process; any barrier across processes) before reading slots written by others. */
int siCountFanoutDevice(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n, uint32_t* d_counts,
                        uint32_t* const* peers, int n_peers, int order, void* stream);
/* Device buffers shareable between the processes of one node (one process per GPU, torchrun): siIpcAlloc = cudaMalloc
 * on the current device + its 64-byte CUDA IPC handle; siIpcOpen maps another process's buffer here (peer access over
 * NVLink); siIpcClose / siIpcFree undo them. Exchange the handles by any means (an all_gather of 64 bytes per rank). */
/* Barrier between GPUs through peer memory, stream-ordered (one 32-thread kernel, no collective library): stores `seq` to
 * signal_ptrs[k] (a flag word in peer k's memory reserved for this rank), then waits until wait_ptrs[k] (this GPU's own
 * flag word for peer k) shows seq or later. Use increasing seq values. *d_timed_out (device word, zeroed by the caller
 * once) becomes 1 if a peer did not arrive within about 10 s. */
int siPeerBarrierDevice(uint32_t* const* signal_ptrs, const uint32_t* const* wait_ptrs, int n_peers, uint32_t seq, uint32_t* d_timed_out,
                        void* stream);
int siIpcAlloc(size_t bytes, void** d_ptr, unsigned char handle_out[64]);
int siIpcOpen(const unsigned char handle[64], void** d_ptr);
int siIpcClose(void* d_ptr);
int siIpcFree(void* d_ptr);

/* ---- 4. several GPUs of one node (single host process; csrc/multi.cu) --------------------------------
 * The reference has no notion of devices: its callers hold one map per chromosome and loop over the
 * queries (examples/bed-intersect-si.rs:100-123). siMulti keeps one replica of ONE index per device and
 * cuts every host batch into one contiguous range per device:
 *   siMultiBuildReplicated    one upload to the first device, ncclBroadcast of the interval columns to
 *                             the others over NVLink, every device builds its replica from device memory;
 *   siMultiCountBatch         H2D of each range on its own PCIe link, one count launch per device whose kernel stores
 *                             every count into EVERY device's gathered vector as it is produced (peer stores over
 *                             NVLink: the all-gather fused into the count, siCountFanoutDevice; the streams then wait
 *                             for each other's kernels through events). Without peer access between the devices
 *                             (or SIB_MULTI_P2P=0): count, then ONE ncclAllGather of the per-query counts. Either way
 *                             every device then holds the whole count vector (siMultiDeviceCounts) and returns its
 *                             own range to counts_out;
 *   siMultiSearchValuesBatch  the gathered counts are scanned on every device into GLOBAL 64-bit CSR offsets,
 *                             each device fills and returns its own segment of `found` (same layout and order
 *                             as searchValuesBatch).
 * devices == NULL: ordinals 0..n_devices-1; n_devices <= 0: every visible device. NCCL (libnccl.so.2) is
 * loaded at run time and only when n_devices > 1. All functions return 0 or a cudaError_t value. */
typedef struct siMulti siMulti;
typedef struct {
    double ms_total;            /* host wall clock of the last batch call */
    double ms_h2d, ms_count, ms_gather, ms_d2h;   /* device time per phase, max over devices (CUDA events) */
    unsigned long long nccl_bytes;                /* bytes received through NCCL collectives since creation, all ranks */
    int nccl_version;
    int peer_access;                              /* 1: every device can store into every other's memory (NVLink): the all-gather of
                                                     the counts is fused into the count kernels (siCountFanoutDevice) */
    unsigned long long peer_bytes;                /* bytes the count kernels stored into other devices' memory since creation */
} siMultiStats;
siMulti* siMultiCreate(const int* devices, int n_devices);
void     siMultiDestroy(siMulti* m);
int      siMultiDeviceCount(const siMulti* m);
siIndex* siMultiIndexOf(siMulti* m, int rank);   /* the replica on one device (siIndexSetOption, siIndexCellsInfo ...) */
int siMultiBuildReplicated(siMulti* m, const int32_t* starts, const int32_t* ends, const int32_t* values, size_t n);
int siMultiCountBatch(siMulti* m, const int32_t* starts, const int32_t* ends, size_t n, uint32_t* counts_out);
int siMultiSearchValuesBatch(siMulti* m, const int32_t* starts, const int32_t* ends, size_t n, size_t* offsets_out,
                             cIndexResult* found);
/* the gathered counts of the last batch on device `rank`: n_devices ranges of *per entries; range r holds the
 * queries [r * per, (r + 1) * per) of the batch (zeros past its end) */
int siMultiDeviceCounts(siMulti* m, int rank, const uint32_t** d_counts, size_t* per);
int siMultiLastStats(const siMulti* m, siMultiStats* out);

/* bytes of device memory currently held by the index + its workspaces */
size_t siIndexDeviceBytes(const siIndex* ix);
/* kernels launched by this library since process start (bench.py's gpu_launches) */
unsigned long long si_b200_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* SUPERINTERVALS_B200_H_INCLUDED */
