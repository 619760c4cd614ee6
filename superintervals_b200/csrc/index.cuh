// index.cuh -- the device-resident index object behind siIndex* (host-side C++).
#pragma once

#include "common.cuh"

#include <mutex>

constexpr int SI_PIPE_SLOTS = 4;   // device slots of the chunked host-batch pipeline (c_abi.cu)
struct siIndex;   // C-visible opaque name

namespace sib {

// cudaMalloc'd buffer that only ever grows (reallocation drops old contents)
// request block of the resident single-query kernel, inside the mapped pinned mailbox (c_abi.cu). The first 16 bytes are
// ONE record the kernel fetches with a single 128-bit load per poll (a read of host memory costs the GPU a PCIe round
// trip, ~2 us: one per poll, not one per field); the host writes a, b, opcap and then seq (x86 store order), and a 16-byte
// aligned read never straddles a cache line, so a record showing the new seq shows the new query.
struct alignas(16) SingleReq {
    uint32_t seq;      // host -> kernel: the call's sequence number, written LAST
    int32_t a, b;      // the query
    uint32_t opcap;    // op << 28 | cap. op: 0 upper_bound, 1 has_overlaps, 2 count, 3 + SI_FILL_* search, 15 leave now; cap: hits that fit the mailbox
    uint32_t alive;    // kernel -> host: lowered when the kernel has left (or is leaving)
    uint32_t pad[3];
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);      // a larger request first looks in the process-wide scratch pool (recycle()), then cudaMalloc
    void release();                // cudaFree (synchronises the device)
    // hands the allocation to the process-wide scratch pool instead of freeing it: the next build's ensure() takes it
    // back without a cudaMalloc / cudaFree pair (each synchronises the device). ONLY when no work that touches the
    // buffer is in flight (the build calls it after its final stream synchronise).
    void recycle();
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

void note_launch(unsigned n = 1);
unsigned long long launches();

// Optional per-launch device timing (siIndexSetOption SI_OPT_TIMING): a CUDA event pair on the
// launching stream around each hot kernel, read back by siIndexReadTimings. Off by default.
enum LaunchTag : int { TAG_PT_HIST = 1, TAG_PT_PASS = 2, TAG_COUNT_WALK = 3, TAG_COUNT_RANK = 4, TAG_SCAN = 5, TAG_FILL = 6, TAG_COUNT_CELLS = 7, TAG_FILL_RUNS = 8,
                       TAG_COUNT_STREAM = 9 };
struct LaunchTimer {
    bool on = false;
    cudaEvent_t* ev = nullptr;   // 2 * cap events
    int* tags = nullptr;
    size_t cap = 0, used = 0;
    bool open = false;
    void begin(int tag, cudaStream_t s);
    void end(cudaStream_t s);
    int read(int* out_tags, float* out_ms, int max_out);   // synchronises the recorded events, then clears
    void release();
};

}  // namespace sib

struct siIndex {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;

    // ---- the index, position order -------------------------------------------------
    uint32_t n = 0, n_padded = 0;
    bool built = false;
    bool wellformed = false;   // every interval has start <= end
    // The closed-form count needs start <= end. A FEW malformed intervals (the reference accepts them, quirk Q6)
    // do not forfeit it: the rank tables are built over the well-formed intervals and the malformed ones are
    // tested one by one per query (mal_*). rank_ok = closed form available; n_rank = intervals the tables cover.
    bool rank_ok = false;
    uint32_t n_rank = 0, n_mal = 0;
    uint32_t mal_pos[8] = {0};
    int32_t mal_s[8] = {0}, mal_e[8] = {0};
    sib::DevBuf starts_wf;     // starts of the well-formed intervals (only when n_mal > 0)
    sib::DevBuf starts, ends, values, branch, perm;
    sib::DevBuf esort;         // ends, each aligned 32-block sorted ascending (count sweep)
    sib::DevBuf eall;          // all ends sorted ascending (count by rank); well-formed indexes only
    int32_t lo = 0, hi = 0;    // smallest start / largest end: the span the partition key buckets
    sib::DevBuf grid_tab;      // rank grid tables (tab_s | tab_e), cells + 1 entries each
    uint32_t grid_shift = 0, grid_cells = 0;
    uint32_t grid_intervals = 8;    // target intervals per grid cell
    // rank cells (RankCells in query_kernels.cuh): one 32-byte record per 2^shift coordinates
    struct CellsMeta { int32_t lo = 0; uint32_t span = 0, shift = 0, fmt = 0, cells = 0; unsigned long long overfull = 0; };
    sib::DevBuf cells_s, cells_e;                  // cells_s owns both tables (one L2 policy window); cells_e stays empty
    uint4* cells_e_ptr = nullptr;                  // inside cells_s
    size_t cells_total_bytes = 0;
    bool narrow_sort = true;                       // SI_OPT_NARROW_SORT: build() sorts by start (32-bit keys) and fixes ties, before falling back to the 64-bit key
    int last_sort = 0;                             // what the last build did: 0 input already sorted, 1 narrow sort, 2 composite 64-bit key
    bool l2_persist = false;                       // SI_OPT_L2_PERSIST / SIB_L2_PERSIST (off: the set-aside costs every other kernel more than it gives, r02g)
    size_t l2_persist_max = 0;                     // cudaLimitPersistingL2CacheSize in force
    CellsMeta cm_s, cm_e;
    // pair cells (PairCells in query_kernels.cuh): both ranks of a coordinate cell in one record, for indexes whose
    // rank cells do not fit L2 -- a stabbing or short query then gathers one sector from HBM instead of two
    sib::DevBuf pair_cells;
    CellsMeta cm_pair;                             // overfull = over-full sides
    bool pair_ok = false;
    int pair_mode = 1;                             // SI_OPT_PAIR_CELLS: 0 never, 1 when the rank cells exceed 3/4 of L2, 2 always; next build
    unsigned long long pair_counters[2] = {0, 0};
    bool pair_counters_pending = false;
    uint32_t cells_fill8 = 16, cells_fill16 = 7;   // target mean values per cell (28 / 14 slots)
    size_t cells_direct_bytes = 0;                 // cells up to this size answer unpartitioned batches (0: always)
    size_t l2_bytes = 0;
    // rank bits (RankBits in query_kernels.cuh): 12 B per 32 coordinates of the span, per table; the
    // streaming count kernel's tables. Built only when they cost at most bits_budget bytes per interval.
    sib::DevBuf bits_s_t, bits_s_d, bits_e_t, bits_e_d;
    unsigned long long stream_tiles = 0;           // tiles of the last streaming count (siIndexStreamStats)
    sib::DevBuf stream_ws;                         // tiles the streaming kernel hands to the rank-cells code (count + list)
    uint32_t bits_words_s = 0, bits_words_e = 0;
    bool bits_ok = false;
    unsigned long long build_counters[4] = {0, 0, 0, 0};   // over-full cells (2 tables), slow rank-bit words (2 tables): read at the build's last synchronise
    bool build_counters_pending = false;
    uint64_t bits_pending_words = 0;
    unsigned long long bits_slow[2] = {0, 0};      // words holding a coordinate with >= 3 values (answered from the cells)
    uint32_t bits_budget = 64;                     // SI_OPT_STREAM_BUDGET: bytes per interval at most (0 = never build)
    int stream_mode = 1;                           // SI_OPT_STREAM: 0 off, 1 position-sorted batches, 2 every batch
    // stab lists (StabLists in query_kernels.cuh), made by the first CSR fill that can use them
    sib::DevBuf stab_off, stab_hdr, stab_ent, stab_cnt;   // scan offsets (build only), list headers, records, counts
    sib::DevBuf stab_entv;                                // 8-byte lists again as (value, end): search_values' payload rides in the record
    bool stab_value_lists = true;                         // SI_OPT_STAB_VALUE_LISTS
    uint32_t stab_kshift = 0, stab_nlists = 0;
    bool stab_rec16 = false;                       // 16-byte (position, end, value) records instead of 8-byte (position, end)
    int stab_state = 0;                            // 0 = not tried yet, 1 = built, 2 = over budget (the fill walks)
    unsigned long long stab_entries = 0;
    bool stab_enabled = true;                      // SI_OPT_STAB_LISTS
    uint32_t stab_budget = 6;                      // SI_OPT_STAB_BUDGET: list entries per interval at most
    sib::DevBuf tree;          // 32-ary max tree levels + prefix-max levels
    const int32_t* pmax32 = nullptr;   // inside `tree`: exclusive prefix max of ends per 32-block

    // ---- build scratch (released after build for large n) --------------------------
    sib::DevBuf b_in_s, b_in_e, b_in_v;        // staged host inputs
    sib::DevBuf b_kA, b_kB, b_vA, b_vB, b_ws;  // 64-bit key sort

    // ---- query scratch ---------------------------------------------------------------
    sib::DevBuf small;                          // device scalars: [0] flags/sorted, ...
    sib::DevBuf q_A, q_B, q_ws;                 // partitioned query records (qs | qe | idx) x 2, workspace
    sib::DevBuf scan_status;
    const int32_t* plan_qs = nullptr;           // query batch the cached partition belongs to
    const int32_t* plan_qe = nullptr;
    size_t plan_n = 0;
    const int32_t* plan_rec_qs = nullptr;       // where its records are (inside q_A / q_B)
    const int32_t* plan_rec_qe = nullptr;
    const uint32_t* plan_rec_idx = nullptr;
    bool plan_valid = false;
    bool plan_armed = false;                    // set by siSortQueriesDevice: next count may reuse

    sib::LaunchTimer timer;
    // The host-buffer entry points of one handle (countOverlapsBatch ... upperBound) share its stream and
    // staging buffers: they take this lock, so that concurrent const queries from several host threads
    // are safe, as with the reference (SURVEY 8b "Threading"). The device API (siCountDevice ...) on caller
    // streams is not locked: callers serialise their calls on one siIndex.
    std::mutex api_mu;

    // ---- options (siIndexSetOption) ----------------------------------------------------
    int count_algo = 0;                         // SI_COUNT_AUTO / SI_COUNT_WALK / SI_COUNT_RANK / SI_COUNT_CELLS
    uint32_t bucket_intervals = 1024;           // partition: index intervals per position bucket
    uint32_t window_shift = 22;                 // partition: log2(queries per result window)

    // ---- staging for the host-buffer API ---------------------------------------------
    sib::DevBuf h_qs, h_qe, h_counts, h_offsets, h_out, h_cov;
    bool pipe_ready = false;                    // chunked host-batch pipeline (c_abi.cu)
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t e_in[SI_PIPE_SLOTS] = {}, e_k[SI_PIPE_SLOTS] = {}, e_out[SI_PIPE_SLOTS] = {};
    void* pinned = nullptr;                     // two pinned staging slots for pageable caller buffers (c_abi.cu)
    size_t pinned_bytes = 0;
    cudaEvent_t e_stage[2] = {nullptr, nullptr};
    sib::DevBuf mixed_tab;                         // siCountMixedDevice: per-contig descriptors of the last mixed launch
    // single-query calls (countOverlaps, searchValues ...): a mapped pinned mailbox the kernels read the query from and
    // write the answer to, so that a call is one launch + one stream synchronise (c_abi.cu)
    void* mailbox = nullptr;
    uint32_t single_seq = 0;
    bool resident = true;                       // SI_OPT_RESIDENT_QUERIES / SIB_RESIDENT_QUERIES: single-query calls are answered by a resident polling warp
    cudaStream_t srv_stream = nullptr;          // its stream
    sib::SingleReq* srv_req = nullptr;          // its request block (to stop it before a rebuild / destroy)
    uint32_t* srv_done = nullptr;               // the mailbox word its answers' sequence numbers go to                    // sequence number of the last single-query call (published by its kernel when done)
    // siCountMixedPeerDevice: the walks over the other GPUs' slices run beside the walk over the home slice (NVLink-bound
    // scans overlap HBM-bound gathers): up to SI_PEER_SIDE side streams, forked from and joined to the caller's stream
    static constexpr int SI_PEER_SIDE = 15;
    cudaStream_t peer_side[SI_PEER_SIDE] = {};
    cudaEvent_t peer_join[SI_PEER_SIDE] = {};
    cudaEvent_t peer_fork = nullptr;
    int peer_side_ready = 0;
    cudaStream_t s_out2 = nullptr;              // second copy-out stream (offsets travel while the fill runs)
    bool pipe_ready_out = false;

    size_t device_bytes() const;
};
