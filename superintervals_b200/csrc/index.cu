// index.cu -- device-resident core of libsuperintervals_b200: the siIndex object,
// build() on device, batch query launchers. Exported with C linkage; declared in
// include/superintervals_b200.h.
#include "../../include/superintervals_b200.h"

#include "build_kernels.cuh"
#include "index.cuh"
#include "partition.cuh"
#include "query_kernels.cuh"
#include "radix_sort.cuh"
#include "stream_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

// ------------------------------------------------------------------------------------
// error side channel + launch counter
// ------------------------------------------------------------------------------------
namespace sib {
namespace {
std::mutex g_err_mu;
int g_err_code = 0;
char g_err_msg[512] = "";
std::atomic<unsigned long long> g_launches{0};
}  // namespace

void set_error(cudaError_t e, const char* what, const char* file, int line) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    if (g_err_code == 0) {   // keep the first failure
        g_err_code = (int)e;
        snprintf(g_err_msg, sizeof(g_err_msg), "CUDA error %d (%s) at %s:%d: %s", (int)e,
                 cudaGetErrorString(e), file, line, what);
    }
    (void)cudaGetLastError();   // un-stick the runtime's last-error slot
}
void set_error_msg(int code, const char* msg) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    if (g_err_code == 0) {
        g_err_code = code;
        snprintf(g_err_msg, sizeof(g_err_msg), "%s", msg);
    }
}
int last_error_code() { return g_err_code; }
void note_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }

// ---- scratch pool: build scratch (sort ping-pong buffers, 24 B/interval) outlives one build ---------------
// cudaMalloc and cudaFree synchronise the device and cost tens to hundreds of microseconds each; a caller that
// builds one index per contig, or rebuilds after add(), would pay ten of them per build. Recycled buffers wait here
// (per device, bounded) and the next ensure() of a fitting size takes one back.
namespace {
struct PoolEntry { void* p; size_t cap; int dev; };
std::mutex g_pool_mu;
std::vector<PoolEntry> g_pool;
size_t g_pool_bytes = 0;
constexpr size_t POOL_MAX_BYTES = (size_t)6 << 30;
constexpr size_t POOL_MAX_ENTRIES = 48;

void* pool_take(size_t want, size_t* cap_out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    size_t best = g_pool.size();
    for (size_t i = 0; i < g_pool.size(); ++i)
        if (g_pool[i].dev == dev && g_pool[i].cap >= want && g_pool[i].cap <= 2 * want + ((size_t)1 << 20) &&
            (best == g_pool.size() || g_pool[i].cap < g_pool[best].cap))
            best = i;
    if (best == g_pool.size()) return nullptr;
    void* p = g_pool[best].p;
    *cap_out = g_pool[best].cap;
    g_pool_bytes -= g_pool[best].cap;
    g_pool[best] = g_pool.back();
    g_pool.pop_back();
    return p;
}
bool pool_give(void* p, size_t cap) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool.size() >= POOL_MAX_ENTRIES || g_pool_bytes + cap > POOL_MAX_BYTES) return false;
    g_pool.push_back({p, cap, dev});
    g_pool_bytes += cap;
    return true;
}
}  // namespace

int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) { SIB_CHECK(cudaFree(p)); p = nullptr; cap = 0; }
    size_t want = (bytes + 255) & ~(size_t)255;
    size_t got = 0;
    if (void* q = pool_take(want, &got)) { p = q; cap = got; return 0; }
    if (cudaMalloc(&p, want) != cudaSuccess) {
        // out of memory with buffers parked in the pool: give them back to the driver and try once more
        (void)cudaGetLastError();
        p = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_pool_mu);
            for (auto& e : g_pool) cudaFree(e.p);
            g_pool.clear();
            g_pool_bytes = 0;
        }
        SIB_CHECK(cudaMalloc(&p, want));
    }
    cap = want;
    return 0;
}
void DevBuf::recycle() {
    if (!p) return;
    if (!pool_give(p, cap)) cudaFree(p);
    p = nullptr;
    cap = 0;
}
void LaunchTimer::begin(int tag, cudaStream_t s) {
    if (!on) return;
    if (used == cap) {
        const size_t ncap = cap ? cap * 2 : 256;
        cudaEvent_t* nev = (cudaEvent_t*)realloc(ev, sizeof(cudaEvent_t) * 2 * ncap);
        int* ntags = (int*)realloc(tags, sizeof(int) * ncap);
        if (!nev || !ntags) { if (nev) ev = nev; if (ntags) tags = ntags; return; }
        ev = nev; tags = ntags;
        for (size_t i = 2 * cap; i < 2 * ncap; ++i)
            if (cudaEventCreate(&ev[i]) != cudaSuccess) { ev[i] = nullptr; }
        cap = ncap;
    }
    tags[used] = tag;
    open = cudaEventRecord(ev[2 * used], s) == cudaSuccess;
}
void LaunchTimer::end(cudaStream_t s) {
    if (!on || !open) return;
    open = false;
    if (cudaEventRecord(ev[2 * used + 1], s) == cudaSuccess) ++used;
}
int LaunchTimer::read(int* out_tags, float* out_ms, int max_out) {
    int k = 0;
    for (size_t i = 0; i < used && k < max_out; ++i) {
        float ms = 0.f;
        if (cudaEventSynchronize(ev[2 * i + 1]) != cudaSuccess) break;
        if (cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]) != cudaSuccess) break;
        out_tags[k] = tags[i];
        out_ms[k] = ms;
        ++k;
    }
    used = 0;
    return k;
}
void LaunchTimer::release() {
    for (size_t i = 0; i < 2 * cap; ++i)
        if (ev[i]) cudaEventDestroy(ev[i]);
    free(ev); free(tags);
    ev = nullptr; tags = nullptr; cap = used = 0;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
}  // namespace sib

using namespace sib;

#define SIB_LAUNCH_T(ix, tag, kernel, grid, block, smem, stream, ...)            \
    do {                                                                        \
        (ix)->timer.begin((tag), (stream));                                     \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);             \
        SIB_CHECK_LAUNCH();                                                     \
        (ix)->timer.end((stream));                                              \
        note_launch();                                                          \
    } while (0)

#define SIB_LAUNCH(kernel, grid, block, smem, stream, ...)                      \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);             \
        SIB_CHECK_LAUNCH();                                                     \
        note_launch();                                                          \
    } while (0)

size_t siIndex::device_bytes() const {
    const DevBuf* all[] = {&starts, &ends, &values, &branch, &perm, &tree, &esort, &eall, &grid_tab, &cells_s, &cells_e, &pair_cells, &bits_s_t, &bits_s_d, &bits_e_t, &bits_e_d, &stream_ws, &starts_wf, &stab_off, &stab_hdr, &stab_ent, &stab_entv, &stab_cnt, &b_in_s, &b_in_e, &b_in_v,
                           &b_kA, &b_kB, &b_vA, &b_vB, &b_ws, &small, &q_A, &q_B,
                           &q_ws, &scan_status, &h_qs, &h_qe, &h_counts, &h_offsets, &h_out, &h_cov};
    size_t s = 0;
    for (auto* b : all) s += b->cap;
    return s;
}

extern "C" int siScanDevice(siIndex* ix, const uint32_t* d_counts, size_t n, uint64_t* d_offsets, void* stream);

namespace {

constexpr size_t MAX_N = 0xFFFFFFFFull - 16384;   // uint32 positions, NONE32 reserved, tile slack

// Device-API streams are taken as given: NULL is the legacy default stream (CUDA
// convention), which is also what torch hands over when no stream context is active.
inline cudaStream_t pick_stream(siIndex*, void* stream) { return (cudaStream_t)stream; }

inline int grid_for(uint64_t n, int threads, int cap) {
    uint64_t g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > (uint64_t)cap) g = cap;
    return (int)g;
}

IndexView view_of(const siIndex* ix) {
    IndexView v;
    v.starts = ix->starts.as<int32_t>();
    v.ends = ix->ends.as<int32_t>();
    v.values = ix->values.as<int32_t>();
    v.branch = ix->branch.as<uint32_t>();
    v.pmax32 = ix->pmax32;
    v.esort = ix->esort.as<int32_t>();
    v.eall = ix->eall.as<int32_t>();
    v.grid.tab_s = ix->grid_tab.as<uint32_t>();
    v.grid.tab_e = ix->grid_tab.as<uint32_t>() + (((size_t)ix->grid_cells + 1 + 31) & ~(size_t)31);
    v.grid.lo = ix->lo;
    v.grid.shift = ix->grid_shift;
    v.grid.cells = ix->grid_cells;
    v.cells_s = RankCells{ix->cells_s.as<uint4>(), ix->cm_s.lo, ix->cm_s.span, ix->cm_s.shift, ix->cm_s.fmt};
    v.cells_e = RankCells{ix->cells_e_ptr, ix->cm_e.lo, ix->cm_e.span, ix->cm_e.shift, ix->cm_e.fmt};
    v.pair = ix->pair_ok ? PairCells{ix->pair_cells.as<uint4>(), ix->cm_pair.lo, ix->cm_pair.span, ix->cm_pair.shift, ix->cm_pair.fmt}
                         : PairCells{nullptr, 0, 0u, 0u, 0u};
    const bool bits = ix->bits_ok;
    v.bits_s = RankBits{bits ? ix->bits_s_t.as<uint2>() : nullptr, ix->bits_s_d.as<uint32_t>(), ix->cm_s.lo, ix->cm_s.span, ix->bits_words_s};
    v.bits_e = RankBits{bits ? ix->bits_e_t.as<uint2>() : nullptr, ix->bits_e_d.as<uint32_t>(), ix->cm_e.lo, ix->cm_e.span, ix->bits_words_e};
    const bool stab = ix->stab_state == 1 && ix->stab_enabled;
    v.stab = StabLists{ix->stab_hdr.as<uint4>(), stab ? ix->stab_ent.p : nullptr, (stab && !ix->stab_rec16 && ix->stab_value_lists) ? ix->stab_entv.p : nullptr,
                       ix->stab_rec16 ? 1u : 0u, ix->stab_kshift, ix->stab_nlists};
    v.n = ix->n;
    v.wellformed = ix->wellformed ? 1u : 0u;
    v.rstarts = ix->n_mal ? ix->starts_wf.as<int32_t>() : ix->starts.as<int32_t>();
    v.n_mal = ix->rank_ok ? ix->n_mal : 0u;
    for (int k = 0; k < 8; ++k) { v.mal_s[k] = ix->mal_s[k]; v.mal_e[k] = ix->mal_e[k]; }
    return v;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

// small device scalars live in ix->small: word 0 = build sortedness flags,
// word 1 = query sortedness flag, word 2 = scan ticket, word 3 = upper_bound result
int ensure_small(siIndex* ix) { return ix->small.ensure(256); }

int build_branch(siIndex* ix, cudaStream_t s) {
    // level sizes
    uint32_t nl[BK_MAX_LEVELS + 2];
    nl[0] = ix->n;
    int produced = 0;
    for (int L = 1; L <= BK_MAX_LEVELS + 1; ++L) nl[L] = (nl[L - 1] + 31) / 32;
    // always build at least levels 1 and 2; stop once a level fits one group of 32
    int top = 1;
    while (nl[top] > 32) ++top;
    produced = top + (top & 1);   // max2 emits levels in pairs: 1-2, 3-4, ...
    if (produced > BK_MAX_LEVELS) produced = BK_MAX_LEVELS + 1;

    size_t total = 0;
    for (int L = 1; L <= produced; ++L) total += ((size_t)nl[L] + 31) & ~(size_t)31;
    size_t ptotal = 0;
    for (int L = 1; L <= top; ++L) ptotal += ((size_t)nl[L] + 31) & ~(size_t)31;
    if (ix->tree.ensure((total + ptotal) * sizeof(int32_t))) return last_error_code();

    MaxTree t;
    memset(&t, 0, sizeof(t));
    int32_t* base = ix->tree.as<int32_t>();
    int32_t* M[BK_MAX_LEVELS + 2];
    int32_t* P[BK_MAX_LEVELS + 2];
    M[0] = ix->ends.as<int32_t>();
    size_t off = 0;
    for (int L = 1; L <= produced; ++L) { M[L] = base + off; off += ((size_t)nl[L] + 31) & ~(size_t)31; }
    for (int L = 1; L <= top; ++L) { P[L] = base + off; off += ((size_t)nl[L] + 31) & ~(size_t)31; }

    for (int L = 0; L < produced; L += 2) {
        int grid = (int)(((uint64_t)nl[L] + 1023) / 1024);
        SIB_LAUNCH(bk_max2_kernel, grid, 1024, 0, s, M[L], nl[L], M[L + 1], M[L + 2]);
    }
    for (int L = top; L >= 1; --L) {
        int grid = (int)(((uint64_t)nl[L] + BK_THREADS - 1) / BK_THREADS);
        SIB_LAUNCH(bk_prefix_level_kernel, grid, BK_THREADS, 0, s, M[L], nl[L],
                   L == top ? (const int32_t*)nullptr : (const int32_t*)P[L + 1], P[L]);
    }
    for (int L = 0; L <= top; ++L) { t.M[L] = M[L]; t.n[L] = nl[L]; }
    t.P1 = P[1];
    t.top = top;
    ix->pmax32 = P[1];   // kept for the query kernels' sweep termination
    int grid = grid_for(nl[1], BK_THREADS / 32, ix->sm_count * 16);
    SIB_LAUNCH(bk_branch_kernel, grid, BK_THREADS, 0, s, t, ix->n, ix->branch.as<uint32_t>());
    return 0;
}

// Rank cells over a sorted device array A[0..n) whose first and last values are known on the
// host: pick the record format and the cell width from the density (cells_plan), then one thread per cell
// (cells_fill). Both tables share one allocation, so that one L2 access-policy window covers them.
void cells_plan(const siIndex* ix, uint32_t n, int32_t first, int32_t last, siIndex::CellsMeta* m) {
    const uint64_t range = (uint64_t)((int64_t)last - (int64_t)first) + 1;
    // largest shift <= smax with a mean of at most `fill` values per cell
    auto pick = [&](uint32_t fill, uint32_t smax) {
        uint32_t sh = 0;
        while (sh < smax && ((uint64_t)n << (sh + 1)) <= (uint64_t)fill * range) ++sh;
        return sh;
    };
    uint32_t shift = pick(ix->cells_fill8, 8), fmt = 1;
    if (((uint64_t)n << shift) < 3 * range) {   // sparse: one-byte offsets would leave the cells empty
        fmt = 2;
        shift = pick(ix->cells_fill16, 16);
    }
    m->lo = first;
    m->span = (uint32_t)(range - 1);
    m->shift = shift;
    m->cells = (uint32_t)(range >> shift) + 1u;   // cell of hi + 1 included
    m->fmt = fmt;
}
size_t cells_bytes(const siIndex::CellsMeta& m) { return (((size_t)m.cells + 1) * 32 + 255) & ~(size_t)255; }

int cells_fill(siIndex* ix, const int32_t* A, uint32_t n, const siIndex::CellsMeta& m, uint4* rec, unsigned long long* d_overfull, cudaStream_t s) {
    const int grid = grid_for((uint64_t)m.cells + 1, BK_THREADS, ix->sm_count * 16);
    if (m.fmt == 1)
        SIB_LAUNCH((bk_rank_cells_kernel<1>), grid, BK_THREADS, 0, s, A, n, m.lo, m.shift, m.cells, rec, d_overfull);
    else
        SIB_LAUNCH((bk_rank_cells_kernel<2>), grid, BK_THREADS, 0, s, A, n, m.lo, m.shift, m.cells, rec, d_overfull);
    return 0;
}

// Pair cells over both arrays (PairCells in query_kernels.cuh). Format by density: four-bit offsets in cells of 16
// coordinates when a cell would hold 3-11 values per side on average, else one-byte offsets in the widest cell
// (<= 256 coordinates) with a mean of at most 4 per side. Returns false when neither fits (denser than ~0.7 values per
// coordinate) or the table would cost more than 24 B per interval (a sparse index: most cells would be empty): the caller
// then keeps the separate tables only.
bool pair_plan(uint32_t n, int64_t first, int64_t last, siIndex::CellsMeta* m) {
    if (first < (int64_t)INT32_MIN || last - first >= (int64_t)0xFFFFFFF0ll) return false;
    const uint64_t range = (uint64_t)(last - first) + 1;
    uint32_t fmt, shift;
    if ((uint64_t)n * 16 >= 3 * range) {
        if ((uint64_t)n * 16 > 11 * range) return false;
        fmt = 4; shift = 4;
    } else {
        fmt = 8; shift = 0;
        while (shift < 8 && ((uint64_t)n << (shift + 1)) <= 4 * range) ++shift;
    }
    const uint64_t cells = (range >> shift) + 1;          // the cell of last + 1 included
    if (cells * 32 > (uint64_t)n * 24 + (1u << 20)) return false;   // too sparse to be worth its memory
    m->lo = (int32_t)first;
    m->span = (uint32_t)(range - 1);
    m->shift = shift;
    m->cells = (uint32_t)cells;
    m->fmt = fmt;
    m->overfull = 0;
    return true;
}
int pair_fill(siIndex* ix, const int32_t* S, const int32_t* E, uint32_t n, const siIndex::CellsMeta& m, unsigned long long* d_overfull, cudaStream_t s) {
    const RankCells cs{ix->cells_s.as<uint4>(), ix->cm_s.lo, ix->cm_s.span, ix->cm_s.shift, ix->cm_s.fmt};
    const RankCells ce{ix->cells_e_ptr, ix->cm_e.lo, ix->cm_e.span, ix->cm_e.shift, ix->cm_e.fmt};
    const int grid = grid_for((uint64_t)m.cells + 1, PC_BUILD_THREADS, ix->sm_count * 16);
    if (m.fmt == 4)
        SIB_LAUNCH((bk_pair_cells_kernel<4>), grid, PC_BUILD_THREADS, 0, s, cs, ce, S, E, n, m.lo, m.shift, m.cells, ix->pair_cells.as<uint4>(), d_overfull);
    else
        SIB_LAUNCH((bk_pair_cells_kernel<8>), grid, PC_BUILD_THREADS, 0, s, cs, ce, S, E, n, m.lo, m.shift, m.cells, ix->pair_cells.as<uint4>(), d_overfull);
    return 0;
}

// Rank bits over a sorted device array whose rank cells exist (stream_kernels.cuh).
int build_bits(siIndex* ix, const int32_t* A, const uint4* cells, const siIndex::CellsMeta& m, DevBuf* t, DevBuf* d2,
               uint32_t* nwords_out, unsigned long long* d_slow, cudaStream_t s) {
    const uint32_t nwords = (uint32_t)(((uint64_t)m.span + 1) >> 5) + 1u;
    const uint32_t padded = ((nwords + 3u) & ~3u) + 4u;
    if (t->ensure((size_t)padded * 8) || d2->ensure((size_t)padded * 4)) return last_error_code();
    const RankCells rc{cells, m.lo, m.span, m.shift, m.fmt};
    SIB_LAUNCH(sk_rank_bits_kernel, grid_for(padded, 256, ix->sm_count * 16), 256, 0, s, rc, A, ix->n, nwords, padded,
               t->as<uint2>(), d2->as<uint32_t>(), d_slow);
    *nwords_out = nwords;
    return 0;
}

extern "C" int si_b200_server_stop_(siIndex* ix);

int build_device_impl(siIndex* ix, const int32_t* d_s, const int32_t* d_e, const int32_t* d_v, size_t n,
                      cudaStream_t s) {
    si_b200_server_stop_(ix);      // a resident single-query kernel reads the arrays that are about to change
    ix->built = false;
    ix->plan_valid = false;
    ix->cm_s.fmt = ix->cm_e.fmt = 0;
    ix->pair_ok = false;
    ix->pair_counters_pending = false;
    ix->bits_ok = false;
    ix->build_counters_pending = false;
    ix->rank_ok = false;
    ix->n_mal = 0;
    ix->stab_state = 0;
    ix->stab_entries = 0;
    if (n > MAX_N) {
        set_error_msg(cudaErrorInvalidValue, "siIndexBuild: more than 2^32-16385 intervals");
        return cudaErrorInvalidValue;
    }
    ix->n = (uint32_t)n;
    ix->n_padded = (uint32_t)(((uint64_t)n + 127) & ~(uint64_t)127);
    if (n == 0) { ix->built = true; return 0; }   // empty build is a no-op (hpp:113-115)
    if (ensure_small(ix)) return last_error_code();
    const size_t pad_b = (size_t)ix->n_padded * 4;
    if (ix->starts.ensure(pad_b) || ix->ends.ensure(pad_b) || ix->branch.ensure(pad_b) ||
        ix->values.ensure(pad_b) || ix->perm.ensure(n * 4))   // values padded: the fill reads aligned 128-bit groups
        return last_error_code();
    uint32_t* d_flags = ix->small.as<uint32_t>();
    const int cap = ix->sm_count * 16;

    // sortedness, as add() would have tracked it (hpp:96-101)
    uint32_t three = 7;
    SIB_CHECK(cudaMemcpyAsync(d_flags, &three, 4, cudaMemcpyHostToDevice, s));
    SIB_LAUNCH(bk_check_sorted_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, d_s, d_e, ix->n, d_flags);
    uint32_t flags = 0;
    SIB_CHECK(cudaMemcpyAsync(&flags, d_flags, 4, cudaMemcpyDeviceToHost, s));
    SIB_CHECK(cudaStreamSynchronize(s));

    ix->wellformed = (flags & 4u) != 0;
    if ((flags & 3u) == 3u) {
        // already (start asc, end desc): the reference does not sort (hpp:1416,1421)
        ix->last_sort = 0;
        SIB_LAUNCH(bk_identity_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, d_s, d_e, d_v, ix->n,
                   ix->starts.as<int32_t>(), ix->ends.as<int32_t>(), ix->values.as<int32_t>(),
                   ix->perm.as<uint32_t>());
    } else {
        // Narrow sort first (SI_OPT_NARROW_SORT, default on): 32-bit keys = the starts, four passes of 16 B per record instead of
        // eight of 24 B, then equal starts put in end-descending order in place (bk_fix_ties_kernel). A run of more
        // than BK_TIE_MAX equal starts sends the build to the composite 64-bit key below.
        bool sorted_narrow = false;
        if (ix->narrow_sort) {
            if (ix->b_kA.ensure(n * 4) || ix->b_kB.ensure(n * 4) || ix->b_vA.ensure(n * 4) || ix->b_vB.ensure(n * 4) ||
                ix->b_ws.ensure(rs_workspace_bytes<uint32_t>(ix->n)))
                return last_error_code();
            uint32_t* d_tie = ix->small.as<uint32_t>() + 3;
            SIB_CHECK(cudaMemsetAsync(d_tie, 0, 4, s));
            SIB_LAUNCH(bk_make_start_keys_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, d_s, ix->n,
                       ix->b_kA.as<uint32_t>(), ix->b_vA.as<uint32_t>());
            int rc = radix_sort_pairs<uint32_t>(ix->b_kA.as<uint32_t>(), ix->b_kB.as<uint32_t>(), ix->b_vA.as<uint32_t>(),
                                                ix->b_vB.as<uint32_t>(), ix->n, 32, ix->b_ws.p, ix->sm_count, s);
            if (rc) return rc;
            RsWorkspace ws = rs_carve(ix->b_ws.p);
            SIB_LAUNCH(bk_gather_narrow_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, ix->b_kA.as<uint32_t>(),
                       ix->b_kB.as<uint32_t>(), ix->b_vA.as<uint32_t>(), ix->b_vB.as<uint32_t>(), ws.final_sel, d_e, d_v, ix->n,
                       ix->starts.as<int32_t>(), ix->ends.as<int32_t>(), ix->values.as<int32_t>(), ix->perm.as<uint32_t>());
            SIB_LAUNCH(bk_fix_ties_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, ix->starts.as<int32_t>(),
                       ix->ends.as<int32_t>(), ix->values.as<int32_t>(), ix->perm.as<uint32_t>(), ix->n, d_tie);
            uint32_t tie = 0;
            SIB_CHECK(cudaMemcpyAsync(&tie, d_tie, 4, cudaMemcpyDeviceToHost, s));
            SIB_CHECK(cudaStreamSynchronize(s));
            sorted_narrow = tie == 0;
            ix->last_sort = sorted_narrow ? 1 : 2;
        } else {
            ix->last_sort = 2;
        }
        if (!sorted_narrow) {
        if (ix->b_kA.ensure(n * 8) || ix->b_kB.ensure(n * 8) || ix->b_vA.ensure(n * 4) ||
            ix->b_vB.ensure(n * 4) || ix->b_ws.ensure(rs_workspace_bytes<uint64_t>(ix->n)))
            return last_error_code();
        SIB_LAUNCH(bk_make_keys_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, d_s, d_e, ix->n,
                   ix->b_kA.as<uint64_t>(), ix->b_vA.as<uint32_t>());
        int rc = radix_sort_pairs<uint64_t>(ix->b_kA.as<uint64_t>(), ix->b_kB.as<uint64_t>(),
                                            ix->b_vA.as<uint32_t>(), ix->b_vB.as<uint32_t>(), ix->n, 64,
                                            ix->b_ws.p, ix->sm_count, s);
        if (rc) return rc;
        RsWorkspace ws = rs_carve(ix->b_ws.p);
        SIB_LAUNCH(bk_gather_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, ix->b_kA.as<uint64_t>(),
                   ix->b_kB.as<uint64_t>(), ix->b_vA.as<uint32_t>(), ix->b_vB.as<uint32_t>(), ws.final_sel, d_v,
                   ix->n, ix->starts.as<int32_t>(), ix->ends.as<int32_t>(), ix->values.as<int32_t>(),
                   ix->perm.as<uint32_t>());
        }
    }
    if (ix->n_padded > ix->n) {
        SIB_LAUNCH(bk_pad_kernel, 1, 128, 0, s, ix->starts.as<int32_t>(), ix->ends.as<int32_t>(), ix->values.as<int32_t>(),
                   ix->branch.as<uint32_t>(), ix->n, ix->n_padded);
    }
    if (ix->esort.ensure(pad_b)) return last_error_code();
    SIB_LAUNCH(bk_sort_blocks_kernel, grid_for(ix->n_padded / 32, BK_THREADS / 32, cap), BK_THREADS, 0, s,
               ix->ends.as<int32_t>(), ix->n_padded, ix->esort.as<int32_t>());
    int rc = build_branch(ix, s);
    if (rc) return rc;

    // span of the index on the host (the partition key buckets it); callers synchronise `s`
    SIB_CHECK(cudaMemcpyAsync(&ix->lo, ix->starts.p, 4, cudaMemcpyDeviceToHost, s));
    ix->rank_ok = ix->wellformed;
    ix->n_mal = 0;
    ix->n_rank = ix->n;
    const int32_t* rstarts = ix->starts.as<int32_t>();
    if (!ix->wellformed) {
        // A few intervals with start > end (the reference stores them, quirk Q6) do not forfeit the closed-form
        // count: list them; the rank tables then cover the well-formed intervals and every query tests the listed ones.
        uint32_t* d_mal = ix->small.as<uint32_t>() + 32;   // count + 8 x (position, start, end)
        SIB_CHECK(cudaMemsetAsync(d_mal, 0, 4 * (1 + 3 * BK_MAX_MALFORMED), s));
        SIB_LAUNCH(bk_find_malformed_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, ix->starts.as<int32_t>(),
                   ix->ends.as<int32_t>(), ix->n, (uint32_t)BK_MAX_MALFORMED, d_mal);
        uint32_t h_mal[1 + 3 * BK_MAX_MALFORMED] = {0};
        SIB_CHECK(cudaMemcpyAsync(h_mal, d_mal, sizeof(h_mal), cudaMemcpyDeviceToHost, s));
        SIB_CHECK(cudaStreamSynchronize(s));
        if (h_mal[0] <= (uint32_t)BK_MAX_MALFORMED && h_mal[0] < ix->n) {
            ix->n_mal = h_mal[0];
            // by position, ascending (insertion sort: at most 8 entries)
            for (uint32_t a = 0; a < ix->n_mal; ++a) { ix->mal_pos[a] = h_mal[1 + 3 * a]; ix->mal_s[a] = (int32_t)h_mal[2 + 3 * a]; ix->mal_e[a] = (int32_t)h_mal[3 + 3 * a]; }
            for (uint32_t a = 1; a < ix->n_mal; ++a)
                for (uint32_t b = a; b > 0 && ix->mal_pos[b - 1] > ix->mal_pos[b]; --b) {
                    std::swap(ix->mal_pos[b - 1], ix->mal_pos[b]); std::swap(ix->mal_s[b - 1], ix->mal_s[b]); std::swap(ix->mal_e[b - 1], ix->mal_e[b]);
                }
            ix->n_rank = ix->n - ix->n_mal;
            const uint32_t wf_padded = (uint32_t)(((uint64_t)ix->n_rank + 127) & ~(uint64_t)127);
            if (ix->starts_wf.ensure((size_t)wf_padded * 4)) return last_error_code();
            MalformedList ml;
            ml.n = ix->n_mal;
            for (uint32_t a = 0; a < (uint32_t)BK_MAX_MALFORMED; ++a) ml.pos[a] = a < ix->n_mal ? ix->mal_pos[a] : 0u;
            SIB_LAUNCH(bk_compact_wellformed_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, ix->starts.as<int32_t>(), ix->n, ml,
                       wf_padded, ix->starts_wf.as<int32_t>());
            rstarts = ix->starts_wf.as<int32_t>();
            ix->rank_ok = true;
        }
    }
    if (ix->rank_ok) {
        const uint32_t nr = ix->n_rank;
        // every (well-formed) end, ascending: the second array of the count-by-rank kernel
        if (ix->eall.ensure(pad_b) || ix->b_kA.ensure(n * 4) || ix->b_kB.ensure(n * 4) || ix->b_ws.ensure(rs_workspace_bytes<uint32_t>(ix->n)))
            return last_error_code();
        SIB_LAUNCH(bk_end_keys_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, ix->starts.as<int32_t>(), ix->ends.as<int32_t>(),
                   ix->n, ix->b_kA.as<uint32_t>());
        // keys only: no payload rides along (radix_sort.cuh, vA == nullptr)
        rc = radix_sort_pairs<uint32_t>(ix->b_kA.as<uint32_t>(), ix->b_kB.as<uint32_t>(), nullptr, nullptr, ix->n, 32, ix->b_ws.p,
                                        ix->sm_count, s);
        if (rc) return rc;
        RsWorkspace ws = rs_carve(ix->b_ws.p);
        SIB_LAUNCH(bk_sorted_ends_kernel, grid_for(ix->n_padded, BK_THREADS, cap), BK_THREADS, 0, s,
                   ix->b_kA.as<uint32_t>(), ix->b_kB.as<uint32_t>(), ws.final_sel, nr, ix->n_padded,
                   ix->eall.as<int32_t>());
        SIB_CHECK(cudaMemcpyAsync(&ix->hi, ix->eall.as<int32_t>() + (nr - 1), 4, cudaMemcpyDeviceToHost, s));
        int32_t last_start = 0, first_end = 0, first_start = 0;
        SIB_CHECK(cudaMemcpyAsync(&first_start, rstarts, 4, cudaMemcpyDeviceToHost, s));
        SIB_CHECK(cudaMemcpyAsync(&last_start, rstarts + (nr - 1), 4, cudaMemcpyDeviceToHost, s));
        SIB_CHECK(cudaMemcpyAsync(&first_end, ix->eall.as<int32_t>(), 4, cudaMemcpyDeviceToHost, s));
        // rank grid over [lo, hi]: the span has to be known on the host to size the tables
        SIB_CHECK(cudaStreamSynchronize(s));
        if (nr < 0x80000000u) {   // bit 31 of a cell's rank word flags an over-full cell
            unsigned long long* d_over = reinterpret_cast<unsigned long long*>(ix->small.as<uint32_t>() + 8);
            SIB_CHECK(cudaMemsetAsync(d_over, 0, 32, s));   // [0,1] over-full cells per table, [2,3] slow rank-bit words per table
            siIndex::CellsMeta ms, me;
            cells_plan(ix, nr, first_start, last_start, &ms);
            cells_plan(ix, nr, first_end, ix->hi, &me);
            if (ix->cells_s.ensure(cells_bytes(ms) + cells_bytes(me))) return last_error_code();
            ix->cells_e_ptr = reinterpret_cast<uint4*>(reinterpret_cast<char*>(ix->cells_s.p) + cells_bytes(ms));
            ix->cells_total_bytes = cells_bytes(ms) + cells_bytes(me);
            rc = cells_fill(ix, rstarts, nr, ms, ix->cells_s.as<uint4>(), d_over, s);
            if (rc) return rc;
            rc = cells_fill(ix, ix->eall.as<int32_t>(), nr, me, ix->cells_e_ptr, d_over + 1, s);
            if (rc) return rc;
            ix->cm_s = ms;
            ix->cm_e = me;
            // pair cells: only where the gather is served by HBM (rank cells beyond 3/4 of L2), or on request
            ix->pair_ok = false;
            ix->pair_counters_pending = false;
            if (ix->pair_mode == 2 || (ix->pair_mode == 1 && ix->cells_total_bytes > (ix->l2_bytes / 4) * 3)) {
                siIndex::CellsMeta mp;
                const int64_t pfirst = std::min<int64_t>((int64_t)first_start - 1, (int64_t)first_end);
                const int64_t plast = std::max<int64_t>((int64_t)last_start - 1, (int64_t)ix->hi);
                if (pair_plan(nr, pfirst, plast, &mp)) {
                    if (ix->pair_cells.ensure(((size_t)mp.cells + 1) * 32)) return last_error_code();
                    unsigned long long* d_pover = reinterpret_cast<unsigned long long*>(ix->small.as<uint32_t>() + 58);
                    SIB_CHECK(cudaMemsetAsync(d_pover, 0, 16, s));
                    rc = pair_fill(ix, rstarts, ix->eall.as<int32_t>(), nr, mp, d_pover, s);
                    if (rc) return rc;
                    ix->cm_pair = mp;
                    ix->pair_ok = true;
                    SIB_CHECK(cudaMemcpyAsync(ix->pair_counters, d_pover, 16, cudaMemcpyDeviceToHost, s));
                    ix->pair_counters_pending = true;
                }
            }
            if (!ix->pair_ok) ix->pair_cells.release();
            // rank bits for the streaming count: only where they stay affordable next to the index
            const uint64_t words = (((uint64_t)ix->cm_s.span + 1) >> 5) + (((uint64_t)ix->cm_e.span + 1) >> 5) + 16;
            ix->bits_pending_words = 0;
            if (ix->stream_mode != 0 && ix->n_mal == 0 && n < 0x40000000ull && words * 12 <= (uint64_t)ix->bits_budget * n &&
                ix->cm_s.span < 0xFFFFFFF0u && ix->cm_e.span < 0xFFFFFFF0u) {   // the kernel's clamps are 32-bit
                rc = build_bits(ix, ix->starts.as<int32_t>(), ix->cells_s.as<uint4>(), ix->cm_s, &ix->bits_s_t, &ix->bits_s_d, &ix->bits_words_s, d_over + 2, s);
                if (rc) return rc;
                rc = build_bits(ix, ix->eall.as<int32_t>(), ix->cells_e_ptr, ix->cm_e, &ix->bits_e_t, &ix->bits_e_d, &ix->bits_words_e, d_over + 3, s);
                if (rc) return rc;
                ix->bits_pending_words = words;
            }
            // over-full cells and slow words are read once, with the build's last synchronise (finish_build)
            SIB_CHECK(cudaMemcpyAsync(ix->build_counters, d_over, 32, cudaMemcpyDeviceToHost, s));
            ix->build_counters_pending = true;
        }
        if (ix->wellformed) {
            const uint64_t range = (uint64_t)((int64_t)ix->hi - (int64_t)ix->lo) + 1;   // >= 1 on a well-formed index
            int gbits = 0;
            while (gbits < 24 && ((uint64_t)ix->grid_intervals << gbits) < n) ++gbits;
            uint32_t shift = 0;
            while (((range - 1) >> shift) >= ((uint64_t)1 << gbits)) ++shift;
            ix->grid_shift = shift;
            ix->grid_cells = (uint32_t)((range - 1) >> shift) + 1u;
            const size_t per = ((size_t)ix->grid_cells + 1 + 31) & ~(size_t)31;
            if (ix->grid_tab.ensure(2 * per * sizeof(uint32_t))) return last_error_code();
            SIB_LAUNCH(bk_rank_grid_kernel, grid_for((uint64_t)ix->grid_cells + 1, BK_THREADS, cap), BK_THREADS, 0, s,
                       ix->starts.as<int32_t>(), ix->eall.as<int32_t>(), ix->n, ix->lo, shift, ix->grid_cells,
                       ix->grid_tab.as<uint32_t>(), ix->grid_tab.as<uint32_t>() + per);
        }
    } else {
        SIB_CHECK(cudaMemcpyAsync(&ix->hi, ix->starts.as<int32_t>() + (ix->n - 1), 4, cudaMemcpyDeviceToHost, s));
    }
    ix->built = true;
    return 0;
}

// the counters the build's kernels left behind, once its stream has been synchronised
void finish_build(siIndex* ix) {
    if (ix->pair_counters_pending) {
        ix->pair_counters_pending = false;
        ix->cm_pair.overfull = ix->pair_counters[0] + ix->pair_counters[1];
        // an over-full side costs a halving search in HBM: beyond one side in a thousand the separate tables answer
        if (ix->cm_pair.overfull * 1000 > (unsigned long long)ix->cm_pair.cells * 2) ix->pair_ok = false;
    }
    if (!ix->build_counters_pending) return;
    ix->build_counters_pending = false;
    ix->cm_s.overfull = ix->build_counters[0];
    ix->cm_e.overfull = ix->build_counters[1];
    if (ix->bits_pending_words) {
        ix->bits_slow[0] = ix->build_counters[2];
        ix->bits_slow[1] = ix->build_counters[3];
        // words with a triple coordinate are answered from the cells, one dependent global load each:
        // beyond a few percent of the words the streaming kernel would not stream
        ix->bits_ok = (ix->bits_slow[0] + ix->bits_slow[1]) * 32 <= ix->bits_pending_words;
    }
}

void release_build_scratch(siIndex* ix) {
    // the sort's ping-pong buffers are 24 B/interval: give them back once the index stands
    if ((size_t)ix->n * 24 > ((size_t)64 << 20)) {
        // (to the scratch pool: the caller has synchronised the build's stream, nothing that touches them is in flight)
        ix->b_kA.recycle(); ix->b_kB.recycle(); ix->b_vA.recycle(); ix->b_vB.recycle(); ix->b_ws.recycle();
        ix->b_in_s.recycle(); ix->b_in_e.recycle(); ix->b_in_v.recycle();
    }
}

// which count kernel answers this index (SI_OPT_COUNT_ALGO; results are identical)
int count_algo_of(const siIndex* ix) {
    const bool cells_ok = ix->rank_ok && ix->cm_s.fmt && ix->cm_e.fmt;
    switch (ix->count_algo) {
        case SI_COUNT_WALK: return SI_COUNT_WALK;
        case SI_COUNT_RANK: return ix->wellformed ? SI_COUNT_RANK : SI_COUNT_WALK;
        default: return cells_ok ? SI_COUNT_CELLS : ix->wellformed ? SI_COUNT_RANK : SI_COUNT_WALK;
    }
}

// Rank cells small enough to stay in L2 answer a batch in whatever order it arrives: no partition.
bool cells_direct(const siIndex* ix) {
    if (count_algo_of(ix) != SI_COUNT_CELLS) return false;
    // Measured (tools/exp_r02g.py, profiles/README.md r02g): gathering the two 32-byte records straight from HBM beats
    // partitioning the batch first at every table size tried (256 MB of cells: 1.6 ms against 4.1 ms for 64 M queries;
    // 768 MB: 10.0 against 21.2 ms for 256 M) -- two partition passes move 48 B/query, the gather 64 B/query of sectors
    // plus nothing else. The partition stays for the walk / rank-grid kernels and behind SI_OPT_CELLS_DIRECT_BYTES.
    if (!ix->cells_direct_bytes) return true;
    return ((size_t)ix->cm_s.cells + ix->cm_e.cells + 2) * 32 <= ix->cells_direct_bytes;
}

// The CSR fill follows the count kernel: with rank cells it copies each query's certain run and
// walks only below it (qk_fill_runs_kernel); otherwise the plain walk (qk_fill_kernel).
bool fill_by_runs(const siIndex* ix) { return ix->wellformed && count_algo_of(ix) == SI_COUNT_CELLS; }

// Partition a query batch for locality (partition.cuh): records grouped by (result window,
// position bucket). *out describes the partitioned records; the partition of the same
// (d_qs, d_qe, nq) may be reused by the fill that follows a count (documented contract).
int partition_queries(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, uint32_t nq, cudaStream_t s,
                      QueryRecords* out, bool may_reuse) {
    if (may_reuse && ix->plan_valid && ix->plan_qs == d_qs && ix->plan_qe == d_qe && ix->plan_n == nq) {
        out->qs = ix->plan_rec_qs; out->qe = ix->plan_rec_qe; out->idx = ix->plan_rec_idx;
        return 0;
    }
    ix->plan_valid = false;
    // the rank kernel only needs its queries in the same few hundred intervals (table and index
    // lines shared in L1); the walk kernel's tile sweep wants a tile of 32 queries within a few
    // dozen intervals of each other, i.e. (almost) start-sorted: finer buckets, one more pass
    const bool walk = count_algo_of(ix) == SI_COUNT_WALK;
    const uint32_t bucket = walk && ix->bucket_intervals > 8 ? 8 : ix->bucket_intervals;
    const PtPlan plan = pt_make_plan(nq, ix->n, ix->lo, ix->hi, bucket, ix->window_shift);
    const size_t cap = ((size_t)nq + 63) & ~(size_t)63;
    if (plan.passes > 0 && (ix->q_A.ensure(cap * 12) || (plan.passes > 1 && ix->q_B.ensure(cap * 12)) ||
                            ix->q_ws.ensure(pt_workspace_bytes(nq))))
        return last_error_code();
    int rc = pt_partition(plan, d_qs, d_qe, nq, ix->q_A.p, ix->q_B.p, cap, ix->q_ws.p, ix->sm_count, s, out,
                          &ix->timer);
    if (rc) return rc;
    ix->plan_qs = d_qs; ix->plan_qe = d_qe; ix->plan_n = nq;
    ix->plan_rec_qs = out->qs; ix->plan_rec_qe = out->qe; ix->plan_rec_idx = out->idx;
    ix->plan_valid = true;
    return 0;
}

// Resolve SI_ORDER_AUTO with one device check. Returns <0 on error, else the order to use.
int resolve_order(siIndex* ix, const int32_t* d_qs, uint32_t nq, int order, cudaStream_t s) {
    if (order != SI_ORDER_AUTO) return order;
    if (nq < 4096) return SI_ORDER_ASIS;   // too small for a sort to pay
    if (ensure_small(ix)) return -1;
    uint32_t* d_flag = ix->small.as<uint32_t>() + 1;
    uint32_t one = 1, flag = 0;
    if (cudaMemcpyAsync(d_flag, &one, 4, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
    qk_check_sorted_kernel<<<grid_for(nq, QK_THREADS, ix->sm_count * 16), QK_THREADS, 0, s>>>(d_qs, nq, d_flag);
    note_launch();
    if (cudaMemcpyAsync(&flag, d_flag, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
    return flag ? SI_ORDER_SORTED : SI_ORDER_UNSORTED;
}

template <typename CountT>
int launch_count(siIndex* ix, const QueryRecords& rec, uint32_t nq, CountT* d_counts, cudaStream_t s, const Fanout& fan) {
    const int algo = count_algo_of(ix);
    if (algo == SI_COUNT_CELLS) {
        const int tiles = (int)(((uint64_t)nq + QC_TILE - 1) / QC_TILE);
        const int grid = tiles;
        if (ix->l2_persist && ix->cells_total_bytes <= ix->l2_persist_max) {
            // the rank cells are the only data read more than once: ask L2 to keep them (persisting) while the
            // query and count streams pass through (streaming) -- a per-launch access-policy window
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)grid);
            cfg.blockDim = dim3(QC_THREADS);
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeAccessPolicyWindow;
            at[0].val.accessPolicyWindow.base_ptr = ix->cells_s.p;
            at[0].val.accessPolicyWindow.num_bytes = ix->cells_total_bytes;
            at[0].val.accessPolicyWindow.hitRatio = 1.0f;
            at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            ix->timer.begin(TAG_COUNT_CELLS, s);
            SIB_CHECK(cudaLaunchKernelEx(&cfg, qk_count_cells_kernel<CountT>, view_of(ix), rec, nq, d_counts, fan));
            ix->timer.end(s);
            note_launch();
        } else {
            if (ix->pair_ok)
                SIB_LAUNCH_T(ix, TAG_COUNT_CELLS, (qk_count_cells_kernel<CountT, true>), grid, QC_THREADS, 0, s, view_of(ix), rec, nq, d_counts, fan);
            else
                SIB_LAUNCH_T(ix, TAG_COUNT_CELLS, (qk_count_cells_kernel<CountT>), grid, QC_THREADS, 0, s, view_of(ix), rec, nq, d_counts, fan);
        }
    } else if (algo == SI_COUNT_RANK) {
        const int grid = (int)(((uint64_t)nq + QR_TILE - 1) / QR_TILE);
        SIB_LAUNCH_T(ix, TAG_COUNT_RANK, (qk_count_rank_kernel<CountT>), grid, QR_THREADS, 0, s, view_of(ix), rec, nq, d_counts);
    } else {
        const int grid = (int)(((uint64_t)nq + QK_THREADS - 1) / QK_THREADS);
        SIB_LAUNCH_T(ix, TAG_COUNT_WALK, (qk_count_kernel<CountT>), grid, QK_THREADS, 0, s, view_of(ix), rec, nq, d_counts);
    }
    return 0;
}

// The streaming kernel answers position-sorted batches of an index that carries rank bits.
bool stream_ready(const siIndex* ix) {
    return ix->bits_ok && ix->stream_mode != 0 && ix->count_algo == SI_COUNT_AUTO && count_algo_of(ix) == SI_COUNT_CELLS;
}

template <typename CountT>
int launch_count_stream(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, uint32_t nq, CountT* d_counts, cudaStream_t s, const Fanout& fan) {
    static_assert(SK_TILE % QC_TILE == 0, "a streaming tile is a whole number of rank-cells tiles");
    const int tiles = (int)(((uint64_t)nq + SK_TILE - 1) / SK_TILE);
    uintptr_t align = ((uintptr_t)d_qs) | ((uintptr_t)d_qe) | ((uintptr_t)d_counts);
    for (int f = 0; f < fan.n; ++f) align |= (uintptr_t)fan.p[f];
    const uint32_t vec_ok = (align & 31u) == 0 ? 1u : 0u;
    if (ix->stream_ws.ensure(((size_t)tiles + 16) * 4)) return last_error_code();
    uint32_t* fail_count = ix->stream_ws.as<uint32_t>();
    uint32_t* fail_list = fail_count + 8;
    SIB_CHECK(cudaMemsetAsync(fail_count, 0, 4, s));
    ix->stream_tiles = (unsigned long long)tiles;
    // persistent CTAs: as many as the device holds at once (two staging stages of dynamic shared memory each)
    auto kern = sk_count_stream_kernel<CountT>;
    SIB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM));
    int per_sm = 0;
    SIB_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SK_THREADS, SK_SMEM));
    if (per_sm < 1) per_sm = 1;
    const int grid = tiles < ix->sm_count * per_sm ? tiles : ix->sm_count * per_sm;
    ix->timer.begin(TAG_COUNT_STREAM, s);
    kern<<<grid, SK_THREADS, SK_SMEM, s>>>(view_of(ix), d_qs, d_qe, nq, d_counts, vec_ok, (uint32_t)tiles, fail_list, fail_count, fan);
    SIB_CHECK_LAUNCH();
    const int grid2 = tiles < ix->sm_count * 4 ? tiles : ix->sm_count * 4;
    sk_count_failed_tiles_kernel<CountT><<<grid2, QC_THREADS, 0, s>>>(view_of(ix), d_qs, d_qe, nq, d_counts, fail_list, fail_count, fan);
    SIB_CHECK_LAUNCH();
    ix->timer.end(s);
    note_launch(2);
    return 0;
}

template <typename CountT>
int count_impl(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n, CountT* d_counts, int order,
               void* stream, const Fanout* fanout = nullptr) {
    if (!ix || !ix->built) {
        set_error_msg(cudaErrorNotReady, "siCountDevice: index not built (call build/indexSuperIntervals first)");
        return cudaErrorNotReady;
    }
    if (n == 0) return 0;
    if (n > 0xFFFFFFFFull) {
        set_error_msg(cudaErrorInvalidValue, "siCountDevice: more than 2^32-1 queries in one batch");
        return cudaErrorInvalidValue;
    }
    DeviceGuard g(ix->device);
    cudaStream_t s = pick_stream(ix, stream);
    if (ix->n == 0) {
        SIB_CHECK(cudaMemsetAsync(d_counts, 0, n * sizeof(CountT), s));
        if (fanout)
            for (int f = 0; f < fanout->n; ++f) SIB_CHECK(cudaMemsetAsync(fanout->p[f], 0, n * 4, s));
        return 0;
    }
    // fan-out (fused all-gather): the cells and streaming kernels store into the peers' arrays themselves; the walk and
    // rank-grid kernels (forced by option, or an index without rank cells) are followed by plain copies instead
    const bool fan_in_kernel = fanout && fanout->n > 0 && count_algo_of(ix) == SI_COUNT_CELLS;
    const bool armed = ix->plan_armed;
    ix->plan_armed = false;
    bool streaming = false;
    if (stream_ready(ix)) {
        // a position-sorted batch streams (stream_kernels.cuh); SI_ORDER_AUTO pays one pass over the starts to find out
        if (order == SI_ORDER_AUTO && ix->stream_mode == 1) {
            order = resolve_order(ix, d_qs, (uint32_t)n, order, s);
            if (order < 0) { set_error(cudaGetLastError(), "resolve_order", __FILE__, __LINE__); return last_error_code(); }
        }
        streaming = ix->stream_mode == 2 || order == SI_ORDER_SORTED;
    }
    if (!streaming && cells_direct(ix) && !(armed && order == SI_ORDER_UNSORTED)) order = SI_ORDER_ASIS;   // order is irrelevant to the cells kernel
    order = resolve_order(ix, d_qs, (uint32_t)n, order, s);
    if (order < 0) { set_error(cudaGetLastError(), "resolve_order", __FILE__, __LINE__); return last_error_code(); }
    // batches beyond the partition's size are processed in slices (their scratch is bounded too)
    for (size_t at = 0; at < n; at += PT_MAX_BATCH) {
        const uint32_t m = (uint32_t)(n - at < PT_MAX_BATCH ? n - at : PT_MAX_BATCH);
        Fanout fan;
        fan.n = 0;
        if (fan_in_kernel) {
            fan.n = fanout->n;
            for (int f = 0; f < fanout->n; ++f) fan.p[f] = fanout->p[f] + at;
        }
        if (streaming) {
            ix->plan_valid = false;
            int rc = launch_count_stream<CountT>(ix, d_qs + at, d_qe + at, m, d_counts + at, s, fan);
            if (rc) return rc;
            continue;
        }
        QueryRecords rec{d_qs + at, d_qe + at, nullptr};
        if (order == SI_ORDER_UNSORTED) {
            // an explicit siSortQueriesDevice() on this batch arms a one-shot reuse; otherwise partition now
            int rc = partition_queries(ix, d_qs + at, d_qe + at, m, s, &rec, armed && n <= PT_MAX_BATCH);
            if (rc) return rc;
        } else {
            ix->plan_valid = false;
        }
        int rc = launch_count<CountT>(ix, rec, m, d_counts + at, s, fan);
        if (rc) return rc;
    }
    if (fanout && fanout->n > 0 && !fan_in_kernel && sizeof(CountT) == 4)
        for (int f = 0; f < fanout->n; ++f) SIB_CHECK(cudaMemcpyAsync(fanout->p[f], d_counts, n * 4, cudaMemcpyDefault, s));
    return 0;
}

// Stab lists for qk_fill_runs_kernel, made once per build by the first fill that can use them.
// Checkpoint spacing: 8 positions, doubled (up to 1024) until the lists fit the budget of
// stab_budget entries per interval; an index nested deeper than that keeps the walk.
int ensure_stab_lists(siIndex* ix, cudaStream_t s) {
    if (ix->stab_state != 0 || !ix->stab_enabled) return 0;
    ix->stab_state = 2;
    if (ix->n < 64) return 0;   // nothing to skip
    const uint32_t n = ix->n;
    const uint32_t nl0 = (n >> QK_STAB_SHIFT0) + 1;
    if (ix->stab_cnt.ensure(((size_t)nl0 + 4) * 4) || ensure_small(ix)) return last_error_code();
    unsigned long long* d_tot = reinterpret_cast<unsigned long long*>(ix->small.as<uint32_t>() + 16);
    SIB_CHECK(cudaMemsetAsync(d_tot, 0, sizeof(unsigned long long) * QK_STAB_SPACINGS, s));
    const IndexView v = view_of(ix);
    SIB_LAUNCH((qk_stab_lists_kernel<false, false>), (nl0 + QK_THREADS - 1) / QK_THREADS, QK_THREADS, 0, s, v, QK_STAB_SHIFT0, nl0,
               ix->stab_cnt.as<uint32_t>(), (const uint64_t*)nullptr, (void*)nullptr, (uint4*)nullptr, (int2*)nullptr);
    SIB_LAUNCH(qk_stab_totals_kernel, grid_for(nl0, QK_THREADS, ix->sm_count * 8), QK_THREADS, 0, s,
               ix->stab_cnt.as<uint32_t>(), nl0, d_tot);
    unsigned long long tot[QK_STAB_SPACINGS];
    SIB_CHECK(cudaMemcpyAsync(tot, d_tot, sizeof(tot), cudaMemcpyDeviceToHost, s));
    SIB_CHECK(cudaStreamSynchronize(s));
    const unsigned long long budget = (unsigned long long)ix->stab_budget * n;
    int k = 0;
    while (k < QK_STAB_SPACINGS && tot[k] > budget) ++k;
    if (k == QK_STAB_SPACINGS) return 0;   // too deep: stay with the walk
    const uint32_t kshift = QK_STAB_SHIFT0 + (uint32_t)k;
    const uint32_t nl = (n >> kshift) + 1;
    if (k > 0)   // the kept checkpoints' counts, contiguous
        SIB_LAUNCH((qk_stab_lists_kernel<false, false>), (nl + QK_THREADS - 1) / QK_THREADS, QK_THREADS, 0, s, v, kshift, nl,
                   ix->stab_cnt.as<uint32_t>(), (const uint64_t*)nullptr, (void*)nullptr, (uint4*)nullptr, (int2*)nullptr);
    // long lists (dense data): nearly every record read is a hit, so the value rides in the record
    const bool rec16 = tot[k] >= 16ull * nl;
    if (ix->stab_hdr.ensure((size_t)nl * 16) || ix->stab_off.ensure(((size_t)nl + 1) * 8) || ix->stab_ent.ensure((size_t)(tot[k] ? tot[k] : 1) * (rec16 ? 16 : 8))) return last_error_code();
    int rc = siScanDevice(ix, ix->stab_cnt.as<uint32_t>(), nl, ix->stab_off.as<uint64_t>(), (void*)s);
    if (rc) return rc;
    // short lists (8-byte records): a second copy as (value, end), so that search_values reads a hit's payload with the record
    // instead of gathering it -- one random sector less per stab hit for 8 B per list entry (SI_OPT_STAB_VALUE_LISTS)
    const bool vlists = !rec16 && ix->stab_value_lists;
    if (vlists && ix->stab_entv.ensure((size_t)(tot[k] ? tot[k] : 1) * 8)) return last_error_code();
    if (rec16)
        SIB_LAUNCH((qk_stab_lists_kernel<true, true>), (nl + QK_THREADS - 1) / QK_THREADS, QK_THREADS, 0, s, v, kshift, nl,
                   (uint32_t*)nullptr, ix->stab_off.as<uint64_t>(), ix->stab_ent.p, ix->stab_hdr.as<uint4>(), (int2*)nullptr);
    else
        SIB_LAUNCH((qk_stab_lists_kernel<true, false>), (nl + QK_THREADS - 1) / QK_THREADS, QK_THREADS, 0, s, v, kshift, nl,
                   (uint32_t*)nullptr, ix->stab_off.as<uint64_t>(), ix->stab_ent.p, ix->stab_hdr.as<uint4>(),
                   vlists ? ix->stab_entv.as<int2>() : (int2*)nullptr);
    ix->stab_rec16 = rec16;
    ix->stab_kshift = kshift;
    ix->stab_nlists = nl;
    ix->stab_entries = tot[k];
    ix->stab_state = 1;
    return 0;
}

template <int MODE>
int launch_fill(siIndex* ix, const QueryRecords& rec, uint32_t nq, const uint64_t* d_offsets, void* d_out,
                cudaStream_t s) {
    if (fill_by_runs(ix)) {
        const int grid = (int)(((uint64_t)nq + QF_THREADS - 1) / QF_THREADS);
        SIB_LAUNCH_T(ix, TAG_FILL_RUNS, (qk_fill_runs_kernel<MODE>), grid, QF_THREADS, 0, s, view_of(ix), rec, nq, d_offsets,
                     reinterpret_cast<typename FillOut<MODE>::T*>(d_out));
        return 0;
    }
    const int grid = (int)(((uint64_t)nq + QK_THREADS - 1) / QK_THREADS);
    SIB_LAUNCH_T(ix, TAG_FILL, (qk_fill_kernel<MODE>), grid, QK_THREADS, 0, s, view_of(ix), rec, nq, d_offsets,
               reinterpret_cast<typename FillOut<MODE>::T*>(d_out));
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------
// exported C ABI (superintervals_b200.h section 1 and 3)
// ------------------------------------------------------------------------------------
extern "C" {

int si_b200_last_error(void) { return sib::last_error_code(); }
const char* si_b200_last_error_string(void) { return sib::g_err_msg; }
void si_b200_clear_error(void) {
    std::lock_guard<std::mutex> lk(sib::g_err_mu);
    sib::g_err_code = 0;
    sib::g_err_msg[0] = 0;
}
const char* si_b200_version(void) { return SUPERINTERVALS_B200_VERSION; }
int si_b200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { sib::set_error(e, "cudaGetDeviceCount", __FILE__, __LINE__); return -1; }
    return n;
}
unsigned long long si_b200_kernel_launches(void) { return sib::launches(); }

// the persisting set-aside is a device-wide limit: raised to the maximum on request (never lowered)
static void raise_persisting_set_aside(siIndex* ix) {
    int pmax = 0;
    if (cudaDeviceGetAttribute(&pmax, cudaDevAttrMaxPersistingL2CacheSize, ix->device) == cudaSuccess && pmax > 0) {
        size_t cur = 0;
        if (cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) == cudaSuccess && cur < (size_t)pmax)
            (void)cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)pmax);
        if (cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) == cudaSuccess) ix->l2_persist_max = cur;
        (void)cudaGetLastError();
    }
}

siIndex* siIndexCreate(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { sib::set_error(e, "cudaGetDevice (no CUDA device?)", __FILE__, __LINE__); return nullptr; }
    siIndex* ix = new siIndex();
    ix->device = dev;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) ix->sm_count = sms;
    int l2 = 0;
    if (cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev) == cudaSuccess && l2 > 0) ix->l2_bytes = (size_t)l2;
    if (const char* e = getenv("SIB_L2_PERSIST")) ix->l2_persist = atoi(e) != 0;
    if (const char* e = getenv("SIB_CELLS_DIRECT_BYTES")) ix->cells_direct_bytes = (size_t)atoll(e);
    if (const char* e = getenv("SIB_RESIDENT_QUERIES")) ix->resident = atoi(e) != 0;
    if (ix->l2_persist) raise_persisting_set_aside(ix);
    e = cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { sib::set_error(e, "cudaStreamCreate", __FILE__, __LINE__); delete ix; return nullptr; }
    return ix;
}

void siIndexDestroy(siIndex* ix) {
    if (!ix) return;
    DeviceGuard g(ix->device);
    si_b200_server_stop_(ix);
    if (ix->srv_stream) cudaStreamDestroy(ix->srv_stream);
    DevBuf* all[] = {&ix->starts, &ix->ends, &ix->values, &ix->branch, &ix->perm, &ix->tree, &ix->esort, &ix->eall, &ix->grid_tab,
                     &ix->cells_s, &ix->cells_e, &ix->pair_cells, &ix->bits_s_t, &ix->bits_s_d, &ix->bits_e_t, &ix->bits_e_d, &ix->stream_ws, &ix->mixed_tab, &ix->starts_wf, &ix->stab_off, &ix->stab_hdr, &ix->stab_ent, &ix->stab_entv, &ix->stab_cnt,
                     &ix->b_in_s, &ix->b_in_e, &ix->b_in_v, &ix->b_kA, &ix->b_kB, &ix->b_vA, &ix->b_vB, &ix->b_ws,
                     &ix->small, &ix->q_A, &ix->q_B, &ix->q_ws, &ix->scan_status, &ix->h_qs,
                     &ix->h_qe, &ix->h_counts, &ix->h_offsets, &ix->h_out, &ix->h_cov};
    for (auto* b : all) b->release();
    if (ix->pipe_ready) {
        cudaStreamDestroy(ix->s_in);
        cudaStreamDestroy(ix->s_out);
        for (int k = 0; k < SI_PIPE_SLOTS; ++k) { cudaEventDestroy(ix->e_in[k]); cudaEventDestroy(ix->e_k[k]); cudaEventDestroy(ix->e_out[k]); }
    }
    ix->timer.release();
    if (ix->e_stage[0]) { cudaEventDestroy(ix->e_stage[0]); cudaEventDestroy(ix->e_stage[1]); }
    if (ix->pipe_ready_out) cudaStreamDestroy(ix->s_out2);
    for (int k = 0; k < ix->peer_side_ready; ++k) { cudaStreamDestroy(ix->peer_side[k]); cudaEventDestroy(ix->peer_join[k]); }
    if (ix->peer_fork) cudaEventDestroy(ix->peer_fork);
    if (ix->pinned) cudaFreeHost(ix->pinned);
    if (ix->mailbox) cudaFreeHost(ix->mailbox);
    if (ix->own_stream) cudaStreamDestroy(ix->own_stream);
    delete ix;
}

size_t siIndexSize(const siIndex* ix) { return ix ? ix->n : 0; }
size_t siIndexDeviceBytes(const siIndex* ix) { return ix ? ix->device_bytes() : 0; }

int siIndexCellsInfo(const siIndex* ix, int which, siCellsInfo* out) {
    if (!ix || !out || !ix->built || which < 0 || which > 2) return cudaErrorInvalidValue;
    if (which == 2) {   // the pair cells
        const bool ok = ix->pair_ok;
        out->format = ok ? ix->cm_pair.fmt : 0;
        out->shift = ok ? ix->cm_pair.shift : 0;
        out->cells = ok ? (unsigned long long)ix->cm_pair.cells + 1 : 0;
        out->bytes = out->cells * 32;
        out->overfull = ok ? ix->cm_pair.overfull : 0;
        out->direct = ok ? 1 : 0;
        return 0;
    }
    const siIndex::CellsMeta& m = which ? ix->cm_e : ix->cm_s;
    out->format = m.fmt;
    out->shift = m.shift;
    out->cells = m.fmt ? (unsigned long long)m.cells + 1 : 0;
    out->bytes = out->cells * 32;
    out->overfull = m.overfull;
    out->direct = cells_direct(ix) ? 1 : 0;
    return 0;
}

int siIndexBitsInfo(const siIndex* ix, siBitsInfo* out) {
    if (!ix || !out || !ix->built) return cudaErrorInvalidValue;
    out->built = ix->bits_ok ? 1 : 0;
    out->words = ix->bits_ok ? (unsigned long long)ix->bits_words_s + ix->bits_words_e : 0;
    out->bytes = out->words * 12;
    out->slow_words = ix->bits_slow[0] + ix->bits_slow[1];
    return 0;
}

int siIndexStreamStats(siIndex* ix, unsigned long long* tiles, unsigned long long* handed_back) {
    if (!ix || !tiles || !handed_back) return cudaErrorInvalidValue;
    *tiles = ix->stream_tiles;
    *handed_back = 0;
    if (!ix->stream_tiles || !ix->stream_ws.p) return 0;
    DeviceGuard g(ix->device);
    uint32_t f = 0;
    SIB_CHECK(cudaDeviceSynchronize());
    SIB_CHECK(cudaMemcpy(&f, ix->stream_ws.p, 4, cudaMemcpyDeviceToHost));
    *handed_back = f;
    return 0;
}

int siIndexLastSort(const siIndex* ix) { return ix && ix->built ? ix->last_sort : -1; }

int siIndexStabInfo(const siIndex* ix, siStabInfo* out) {
    if (!ix || !out || !ix->built) return cudaErrorInvalidValue;
    out->state = ix->stab_state;
    out->shift = ix->stab_state == 1 ? ix->stab_kshift : 0;
    out->lists = ix->stab_state == 1 ? ix->stab_nlists : 0;
    out->entries = ix->stab_state == 1 ? ix->stab_entries : 0;
    out->record_bytes = ix->stab_state == 1 ? (ix->stab_rec16 ? 16u : 8u) : 0u;
    return 0;
}

int siIndexDeviceView(const siIndex* ix, siDeviceView* out) {
    if (!ix || !out || !ix->built) return cudaErrorNotReady;
    out->starts = ix->starts.as<int32_t>();
    out->ends = ix->ends.as<int32_t>();
    out->values = ix->values.as<int32_t>();
    out->branch = ix->branch.as<uint32_t>();
    out->n = ix->n;
    out->device = ix->device;
    return 0;
}

int siIndexBuildDevice(siIndex* ix, const int32_t* d_starts, const int32_t* d_ends, const int32_t* d_values,
                       size_t n, void* stream) {
    if (!ix) return cudaErrorInvalidValue;
    DeviceGuard g(ix->device);
    cudaStream_t s = pick_stream(ix, stream);
    int rc = build_device_impl(ix, d_starts, d_ends, d_values, n, s);
    if (rc) return rc;
    SIB_CHECK(cudaStreamSynchronize(s));
    finish_build(ix);
    release_build_scratch(ix);
    return 0;
}

int siIndexBuildHost(siIndex* ix, const int32_t* starts, const int32_t* ends, const int32_t* values, size_t n) {
    if (!ix) return cudaErrorInvalidValue;
    DeviceGuard g(ix->device);
    cudaStream_t s = ix->own_stream;
    if (n == 0) return build_device_impl(ix, nullptr, nullptr, nullptr, 0, s);
    if (ix->b_in_s.ensure(n * 4) || ix->b_in_e.ensure(n * 4) || (values && ix->b_in_v.ensure(n * 4)))
        return last_error_code();
    SIB_CHECK(cudaMemcpyAsync(ix->b_in_s.p, starts, n * 4, cudaMemcpyHostToDevice, s));
    SIB_CHECK(cudaMemcpyAsync(ix->b_in_e.p, ends, n * 4, cudaMemcpyHostToDevice, s));
    if (values) SIB_CHECK(cudaMemcpyAsync(ix->b_in_v.p, values, n * 4, cudaMemcpyHostToDevice, s));
    int rc = build_device_impl(ix, ix->b_in_s.as<int32_t>(), ix->b_in_e.as<int32_t>(),
                               values ? ix->b_in_v.as<int32_t>() : nullptr, n, s);
    if (rc) return rc;
    SIB_CHECK(cudaStreamSynchronize(s));
    finish_build(ix);
    release_build_scratch(ix);
    return 0;
}

int siIndexExport(const siIndex* ix, int32_t* starts, int32_t* ends, int32_t* values, size_t* branch,
                  uint32_t* perm) {
    if (!ix || !ix->built) return cudaErrorNotReady;
    if (ix->n == 0) return 0;
    DeviceGuard g(ix->device);
    const size_t n = ix->n;
    if (starts) SIB_CHECK(cudaMemcpy(starts, ix->starts.p, n * 4, cudaMemcpyDeviceToHost));
    if (ends) SIB_CHECK(cudaMemcpy(ends, ix->ends.p, n * 4, cudaMemcpyDeviceToHost));
    if (values) SIB_CHECK(cudaMemcpy(values, ix->values.p, n * 4, cudaMemcpyDeviceToHost));
    if (perm) SIB_CHECK(cudaMemcpy(perm, ix->perm.p, n * 4, cudaMemcpyDeviceToHost));
    if (branch) {
        // widen in place, back to front: the 32-bit copy sits in the first half of the buffer
        uint32_t* tmp = reinterpret_cast<uint32_t*>(branch);
        SIB_CHECK(cudaMemcpy(tmp, ix->branch.p, n * 4, cudaMemcpyDeviceToHost));
        for (size_t i = n; i-- > 0;) {
            uint32_t b = tmp[i];
            branch[i] = b == NONE32 ? SI_NONE : (size_t)b;
        }
    }
    return 0;
}

int siCountDevice(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n, uint32_t* d_counts, int order,
                  void* stream) {
    return count_impl<uint32_t>(ix, d_qs, d_qe, n, d_counts, order, stream);
}
int siCountDevice64(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n, uint64_t* d_counts, int order,
                    void* stream) {
    return count_impl<uint64_t>(ix, d_qs, d_qe, n, d_counts, order, stream);
}

// count + fan-out of every count to further arrays (superintervals_b200.h section 3c): the fused form of
// "count, then all-gather the counts" when the arrays are the other GPUs' copies of the gathered vector
int siCountFanoutDevice(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n, uint32_t* d_counts,
                        uint32_t* const* peers, int n_peers, int order, void* stream) {
    if (n_peers < 0 || n_peers > SI_FANOUT_MAX || (n_peers && !peers)) {
        set_error_msg(cudaErrorInvalidValue, "siCountFanoutDevice: 0..15 peer arrays");
        return cudaErrorInvalidValue;
    }
    Fanout fan;
    fan.n = n_peers;
    for (int f = 0; f < n_peers; ++f) fan.p[f] = peers[f];
    return count_impl<uint32_t>(ix, d_qs, d_qe, n, d_counts, order, stream, &fan);
}


int siSortQueriesDevice(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n, void* stream) {
    if (!ix || !ix->built) {
        set_error_msg(cudaErrorNotReady, "siSortQueriesDevice: index not built");
        return cudaErrorNotReady;
    }
    if (n == 0 || ix->n == 0) return 0;
    if (n > PT_MAX_BATCH) return 0;   // larger batches are partitioned slice by slice inside the query calls
    DeviceGuard g(ix->device);
    QueryRecords rec;
    int rc = partition_queries(ix, d_qs, d_qe, (uint32_t)n, pick_stream(ix, stream), &rec, false);
    ix->plan_armed = rc == 0;
    return rc;
}

int siIndexReadTimings(siIndex* ix, int* tags, float* ms, int max_out) {
    if (!ix || !tags || !ms || max_out < 0) return 0;
    DeviceGuard g(ix->device);
    return ix->timer.read(tags, ms, max_out);
}

int siIndexSetOption(siIndex* ix, int option, long long value) {
    if (!ix) return cudaErrorInvalidValue;
    switch (option) {
        case SI_OPT_COUNT_ALGO:
            if (value < SI_COUNT_AUTO || value > SI_COUNT_CELLS) break;
            ix->count_algo = (int)value;
            return 0;
        case SI_OPT_BUCKET_INTERVALS:
            if (value < 1 || value > (1ll << 30)) break;
            ix->bucket_intervals = (uint32_t)value;
            ix->plan_valid = false;
            return 0;
        case SI_OPT_GRID_INTERVALS:   // takes effect at the next build()
            if (value < 1 || value > (1ll << 20)) break;
            ix->grid_intervals = (uint32_t)value;
            return 0;
        case SI_OPT_TIMING:
            ix->timer.on = value != 0;
            ix->timer.used = 0;
            return 0;
        case SI_OPT_CELLS_DIRECT_BYTES:   // 0 restores the default (60 % of L2); 1 = never direct
            if (value < 0) break;
            ix->cells_direct_bytes = (size_t)value;
            return 0;
        case SI_OPT_CELLS_FILL:           // mean values per cell, one-byte format; applies to the next build
            if (value < 1 || value > 28) break;
            ix->cells_fill8 = (uint32_t)value;
            ix->cells_fill16 = (uint32_t)(value + 1) / 2 > 14 ? 14 : (uint32_t)(value + 1) / 2;
            return 0;
        case SI_OPT_STAB_LISTS:            // 0: the fill walks the branch array below each run; 1 (default): stab lists
            if (value < 0 || value > 1) break;
            ix->stab_enabled = value != 0;
            return 0;
        case SI_OPT_STAB_BUDGET:           // list entries per interval at most; applies to lists not built yet
            if (value < 0 || value > 4096) break;
            ix->stab_budget = (uint32_t)value;
            if (ix->stab_state == 2) ix->stab_state = 0;
            return 0;
        case SI_OPT_STREAM:                // 0: never; 1 (default): position-sorted batches stream; 2: every batch (tests)
            if (value < 0 || value > 2) break;
            ix->stream_mode = (int)value;
            return 0;
        case SI_OPT_STREAM_BUDGET:         // bytes of rank bits per interval at most; applies to the next build
            if (value < 0 || value > 4096) break;
            ix->bits_budget = (uint32_t)value;
            return 0;
        case SI_OPT_STAB_VALUE_LISTS:      // 1 (default): short stab lists also exist as (value, end) records; applies to lists not built yet
            if (value < 0 || value > 1) break;
            ix->stab_value_lists = value != 0;
            return 0;
        case SI_OPT_RESIDENT_QUERIES:      // 1 (default): single-query calls are answered by a resident polling warp; 0: one launch per call
            if (value < 0 || value > 1) break;
            if (!value) si_b200_server_stop_(ix);
            ix->resident = value != 0;
            return 0;
        case SI_OPT_NARROW_SORT:           // 1 (default): build() sorts by start and fixes ties; 0: always the composite 64-bit key
            if (value < 0 || value > 1) break;
            ix->narrow_sort = value != 0;
            return 0;
        case SI_OPT_L2_PERSIST:            // 1: the cells kernel is launched with an L2 access-policy window over the rank cells
            if (value < 0 || value > 1) break;
            ix->l2_persist = value != 0;
            if (ix->l2_persist) raise_persisting_set_aside(ix);
            return 0;
        case SI_OPT_PAIR_CELLS:            // 0: never; 1 (default): when the rank cells exceed 3/4 of L2; 2: always; applies to the next build
            if (value < 0 || value > 2) break;
            ix->pair_mode = (int)value;
            return 0;
        case SI_OPT_WINDOW_SHIFT:
            if (value < 10 || value > 31) break;
            ix->window_shift = (uint32_t)value;
            ix->plan_valid = false;
            return 0;
        default: break;
    }
    set_error_msg(cudaErrorInvalidValue, "siIndexSetOption: unknown option or value out of range");
    return cudaErrorInvalidValue;
}

int siAnyDevice(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n, uint8_t* d_out, void* stream) {
    if (!ix || !ix->built) {
        set_error_msg(cudaErrorNotReady, "siAnyDevice: index not built");
        return cudaErrorNotReady;
    }
    if (n == 0) return 0;
    if (n > 0xFFFFFFFFull) return cudaErrorInvalidValue;
    DeviceGuard g(ix->device);
    cudaStream_t s = pick_stream(ix, stream);
    const int grid = (int)((n + QK_THREADS - 1) / QK_THREADS);
    SIB_LAUNCH(qk_any_kernel, grid, QK_THREADS, 0, s, view_of(ix), d_qs, d_qe, (uint32_t)n, d_out);
    return 0;
}

int siScanDevice(siIndex* ix, const uint32_t* d_counts, size_t n, uint64_t* d_offsets, void* stream) {
    if (!ix) return cudaErrorInvalidValue;
    if (n > 0xFFFFFFFFull) return cudaErrorInvalidValue;
    DeviceGuard g(ix->device);
    cudaStream_t s = pick_stream(ix, stream);
    if (n == 0) {
        SIB_CHECK(cudaMemsetAsync(d_offsets, 0, sizeof(uint64_t), s));
        return 0;
    }
    if ((((uintptr_t)d_counts) | ((uintptr_t)d_offsets)) & 15u) {
        set_error_msg(cudaErrorMisalignedAddress, "siScanDevice: d_counts and d_offsets must be 16-byte aligned");
        return cudaErrorMisalignedAddress;
    }
    const uint32_t tiles = (uint32_t)((n + SC_TILE - 1) / SC_TILE);
    if (ix->scan_status.ensure((size_t)tiles * 8 + 64) || ensure_small(ix)) return last_error_code();
    uint32_t* ticket = ix->small.as<uint32_t>() + 2;
    SIB_CHECK(cudaMemsetAsync(ix->scan_status.p, 0, (size_t)tiles * 8, s));
    SIB_CHECK(cudaMemsetAsync(ticket, 0, 4, s));
    SIB_LAUNCH_T(ix, TAG_SCAN, qk_scan_kernel, tiles, SC_THREADS, 0, s, d_counts, (uint32_t)n, d_offsets,
               ix->scan_status.as<unsigned long long>(), ticket);
    return 0;
}

int siFillDevice(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n, const uint64_t* d_offsets,
                 int what, void* d_out, int order, void* stream) {
    if (!ix || !ix->built) {
        set_error_msg(cudaErrorNotReady, "siFillDevice: index not built");
        return cudaErrorNotReady;
    }
    if (n == 0 || ix->n == 0) return 0;
    if (n > 0xFFFFFFFFull) return cudaErrorInvalidValue;
    DeviceGuard g(ix->device);
    cudaStream_t s = pick_stream(ix, stream);
    const uint32_t nq = (uint32_t)n;
    order = resolve_order(ix, d_qs, nq, order, s);
    if (order < 0) { set_error(cudaGetLastError(), "resolve_order", __FILE__, __LINE__); return last_error_code(); }
    if (what < SI_FILL_VALUES || what > SI_FILL_ITEMS) {
        set_error_msg(cudaErrorInvalidValue, "siFillDevice: unknown fill mode");
        return cudaErrorInvalidValue;
    }
    if (fill_by_runs(ix)) {
        int rc = ensure_stab_lists(ix, s);
        if (rc) return rc;
    }
    for (size_t at = 0; at < n; at += PT_MAX_BATCH) {
        const uint32_t m = (uint32_t)(n - at < PT_MAX_BATCH ? n - at : PT_MAX_BATCH);
        QueryRecords rec{d_qs + at, d_qe + at, nullptr};
        // rank cells that stay in L2 need no locality: the batch is filled in the caller's order
        if (order == SI_ORDER_UNSORTED && !(fill_by_runs(ix) && cells_direct(ix))) {
            int rc = partition_queries(ix, d_qs + at, d_qe + at, m, s, &rec, n <= PT_MAX_BATCH);
            if (rc) return rc;
        }
        int rc = 0;
        switch (what) {
            case SI_FILL_VALUES: rc = launch_fill<FILL_VALUES>(ix, rec, m, d_offsets + at, d_out, s); break;
            case SI_FILL_IDXS: rc = launch_fill<FILL_IDXS>(ix, rec, m, d_offsets + at, d_out, s); break;
            case SI_FILL_KEYS: rc = launch_fill<FILL_KEYS>(ix, rec, m, d_offsets + at, d_out, s); break;
            default: rc = launch_fill<FILL_ITEMS>(ix, rec, m, d_offsets + at, d_out, s); break;
        }
        if (rc) return rc;
    }
    return 0;
}

int siCoverageDevice(siIndex* ix, const int32_t* d_qs, const int32_t* d_qe, size_t n, uint32_t* d_counts,
                     int32_t* d_cov, void* stream) {
    if (!ix || !ix->built) {
        set_error_msg(cudaErrorNotReady, "siCoverageDevice: index not built");
        return cudaErrorNotReady;
    }
    if (n == 0) return 0;
    if (n > 0xFFFFFFFFull) return cudaErrorInvalidValue;
    DeviceGuard g(ix->device);
    cudaStream_t s = pick_stream(ix, stream);
    const int grid = (int)((n + QK_THREADS - 1) / QK_THREADS);
    SIB_LAUNCH(qk_coverage_kernel, grid, QK_THREADS, 0, s, view_of(ix), d_qs, d_qe, (uint32_t)n, d_counts, d_cov);
    return 0;
}

// used by bed.cu: d_perm[i] = index of the i-th record in stable order of d_key (0 <= key < 2^key_bits),
// through the build's radix sort and its scratch buffers
int si_b200_stable_order_(siIndex* ix, const int32_t* d_key, size_t n, int key_bits, uint32_t* d_perm, void* stream) {
    if (!ix || n > 0xFFFFFFFFull) return cudaErrorInvalidValue;
    if (n == 0) return 0;
    DeviceGuard g(ix->device);
    cudaStream_t s = pick_stream(ix, stream);
    const uint32_t n32 = (uint32_t)n;
    if (ix->b_kA.ensure(n * 4) || ix->b_kB.ensure(n * 4) || ix->b_vA.ensure(n * 4) || ix->b_vB.ensure(n * 4) ||
        ix->b_ws.ensure(rs_workspace_bytes<uint32_t>(n32)))
        return last_error_code();
    const int cap = ix->sm_count * 16;
    SIB_LAUNCH(bk_iota_key_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, d_key, n32, ix->b_kA.as<uint32_t>(), ix->b_vA.as<uint32_t>());
    int rc = radix_sort_pairs<uint32_t>(ix->b_kA.as<uint32_t>(), ix->b_kB.as<uint32_t>(), ix->b_vA.as<uint32_t>(), ix->b_vB.as<uint32_t>(),
                                        n32, key_bits, ix->b_ws.p, ix->sm_count, s);
    if (rc) return rc;
    RsWorkspace ws = rs_carve(ix->b_ws.p);
    SIB_LAUNCH(bk_take_perm_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, ix->b_vA.as<uint32_t>(), ix->b_vB.as<uint32_t>(),
               ws.final_sel, n32, d_perm);
    return 0;
}

// Mode B routing (superintervals_b200.h section 3b): the stable order of the contig ids through the build's radix sort,
// the query columns gathered through it, and where each contig's queries begin (one small D2H, synchronises `stream`).
int siRouteByContigDevice(siIndex* ix, const int32_t* d_contig, const int32_t* d_qs, const int32_t* d_qe, size_t n,
                          int n_contigs, int32_t* d_qs_out, int32_t* d_qe_out, uint32_t* d_perm, size_t* offsets_out,
                          void* stream) {
    if (!ix || n > 0xFFFFFFFFull || n_contigs < 1 || n_contigs > (1 << 24)) {
        set_error_msg(cudaErrorInvalidValue, "siRouteByContigDevice: bad arguments");
        return cudaErrorInvalidValue;
    }
    for (int k = 0; k <= n_contigs; ++k) offsets_out[k] = 0;
    if (n == 0) return 0;
    DeviceGuard g(ix->device);
    cudaStream_t s = pick_stream(ix, stream);
    int bits = 1;
    while ((1 << bits) < n_contigs) ++bits;
    int rc = si_b200_stable_order_(ix, d_contig, n, bits, d_perm, stream);
    if (rc) return rc;
    if (ix->b_in_s.ensure(n * 4) || ix->b_in_e.ensure(((size_t)n_contigs + 2) * 8)) return last_error_code();
    const int cap = ix->sm_count * 16;
    unsigned long long* d_bad = ix->b_in_e.as<unsigned long long>() + n_contigs + 1;
    SIB_CHECK(cudaMemsetAsync(d_bad, 0, 8, s));
    SIB_LAUNCH(bk_route_gather_kernel, grid_for(n, BK_THREADS, cap), BK_THREADS, 0, s, d_perm, d_contig, d_qs, d_qe, (uint32_t)n,
               (uint32_t)n_contigs, ix->b_in_s.as<int32_t>(), d_qs_out, d_qe_out, d_bad);
    SIB_LAUNCH(bk_key_offsets_kernel, (n_contigs + 1 + 127) / 128, 128, 0, s, ix->b_in_s.as<int32_t>(), (uint32_t)n, (uint32_t)n_contigs,
               ix->b_in_e.as<unsigned long long>());
    static_assert(sizeof(size_t) == 8, "LP64 only");
    SIB_CHECK(cudaMemcpyAsync(offsets_out, ix->b_in_e.p, ((size_t)n_contigs + 1) * 8, cudaMemcpyDeviceToHost, s));
    unsigned long long bad = 0;
    SIB_CHECK(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, s));
    SIB_CHECK(cudaStreamSynchronize(s));
    if (bad) {
        set_error_msg(cudaErrorInvalidValue, "siRouteByContigDevice: contig id out of range");
        return cudaErrorInvalidValue;
    }
    return 0;
}

int siScatterCountsDevice(siIndex* ix, const uint32_t* d_counts, const uint32_t* d_perm, size_t n, uint32_t* d_out, void* stream) {
    if (!ix || n > 0xFFFFFFFFull) return cudaErrorInvalidValue;
    if (n == 0) return 0;
    DeviceGuard g(ix->device);
    cudaStream_t s = pick_stream(ix, stream);
    SIB_LAUNCH(bk_scatter_u32_kernel, grid_for(n, BK_THREADS, ix->sm_count * 16), BK_THREADS, 0, s, d_counts, d_perm, (uint32_t)n, d_out);
    return 0;
}

// Mode B in one launch (superintervals_b200.h section 3b): the mixed batch answered in the caller's order by
// qk_count_mixed_kernel. Returns SI_MIXED_UNSUPPORTED (no error latched) when one of the indexes cannot answer from
// rank cells (malformed beyond the side list, >= 2^31 intervals): the caller then routes by contig instead.
// one launch of the mixed-batch kernel over one slice; the descriptor table is already on the device
static int mixed_launch(siIndex* host, siIndex* const* ixs, int n_contigs, size_t table_bytes_on_device, const int32_t* d_contig,
                        const int32_t* d_qs, const int32_t* d_qe, size_t n, uint32_t* d_counts, unsigned long long* d_totals,
                        uint32_t peer_mode, cudaStream_t s) {
    const size_t bytes = table_bytes_on_device;
    const size_t smem = n_contigs <= QM_SMEM_ENTRIES ? ((bytes + 15) & ~(size_t)15) + (size_t)n_contigs * 8 : 0;
    auto kern = qk_count_mixed_kernel<uint32_t>;
    // the kernel also holds 8.2 KB of static shared memory (the compaction list of the peer modes): opt in beyond 48 KB in total
    if (smem + ((size_t)9 << 10) > ((size_t)48 << 10)) SIB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SIB_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, QM_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const uint64_t tiles = ((uint64_t)n + QM_THREADS - 1) / QM_THREADS;
    // tables beyond L2 (a whole genome): short-lived CTAs of `rounds` tiles each; tables that fit: persistent CTAs
    size_t table_bytes = 0;
    for (int k = 0; k < n_contigs; ++k)
        if (ixs[k] && ixs[k]->built && ixs[k]->n) table_bytes += ixs[k]->pair_ok ? ((size_t)ixs[k]->cm_pair.cells + 1) * 32 : ixs[k]->cells_total_bytes;
    uint32_t rounds = table_bytes > (host->l2_bytes / 4) * 3 ? 8u : 0u;
    if (const char* e = getenv("SIB_QM_ROUNDS")) rounds = (uint32_t)atoi(e);
    if (peer_mode != QM_PLAIN && rounds > QM_MAX_ROUNDS) rounds = QM_MAX_ROUNDS;   // the kernel's compaction list holds 8 tiles
    const int grid = rounds ? (int)((tiles + rounds - 1) / rounds) : (int)std::min<uint64_t>(tiles, (uint64_t)host->sm_count * per_sm);
    const MixedEntry* d_tab = reinterpret_cast<const MixedEntry*>(host->mixed_tab.p);
    SIB_LAUNCH_T(host, TAG_COUNT_CELLS, kern, grid, QM_THREADS, smem, s, d_tab, (uint32_t)n_contigs, d_contig, d_qs, d_qe,
                 (uint32_t)n, d_counts, d_totals, rounds, peer_mode);
    return 0;
}

// the per-contig descriptor table of a mixed launch, uploaded into host->mixed_tab; foreign[k] != 0 marks a contig without an
// index here that another GPU answers (siCountMixedPeerDevice)
static int mixed_table(siIndex* host, siIndex* const* ixs, int n_contigs, const unsigned char* foreign, cudaStream_t s, size_t* bytes_out) {
    std::vector<MixedEntry> tab((size_t)n_contigs);
    memset(tab.data(), 0, tab.size() * sizeof(MixedEntry));
    for (int k = 0; k < n_contigs; ++k) {
        siIndex* ix = ixs[k];
        MixedEntry& e = tab[(size_t)k];
        if (!ix || !ix->built || ix->n == 0) {
            if (foreign && foreign[k]) e.n_mal = QM_FOREIGN;
            continue;
        }
        const IndexView v = view_of(ix);
        e.cs = v.cells_s; e.ce = v.cells_e; e.pc = v.pair; e.rstarts = v.rstarts; e.eall = v.eall; e.ends = v.ends; e.branch = v.branch;
        e.n = v.n; e.n_mal = v.n_mal;
        for (int a = 0; a < 8; ++a) { e.mal_s[a] = v.mal_s[a]; e.mal_e[a] = v.mal_e[a]; }
    }
    const size_t bytes = tab.size() * sizeof(MixedEntry);
    if (host->mixed_tab.ensure(bytes)) return last_error_code();
    // pageable source: the runtime stages it before returning, so the vector may go out of scope
    SIB_CHECK(cudaMemcpyAsync(host->mixed_tab.p, tab.data(), bytes, cudaMemcpyHostToDevice, s));
    *bytes_out = bytes;
    return 0;
}

// the index that lends its stream, device and table allocation; SI_MIXED_UNSUPPORTED when an index cannot answer from rank cells
static int mixed_host(siIndex* const* ixs, int n_contigs, siIndex** host_out, const char* who) {
    siIndex* host = nullptr;
    for (int k = 0; k < n_contigs; ++k) {
        siIndex* ix = ixs[k];
        if (!ix || !ix->built || ix->n == 0) continue;
        if (count_algo_of(ix) != SI_COUNT_CELLS) return SI_MIXED_UNSUPPORTED;
        if (!host) host = ix;
        if (ix->device != host->device) {
            set_error_msg(cudaErrorInvalidValue, who);
            return cudaErrorInvalidValue;
        }
    }
    *host_out = host;
    return 0;
}

int siCountMixedDevice(siIndex* const* ixs, int n_contigs, const int32_t* d_contig, const int32_t* d_qs, const int32_t* d_qe,
                       size_t n, uint32_t* d_counts, unsigned long long* d_totals, void* stream) {
    if (!ixs || n_contigs < 1 || n_contigs > (1 << 20) || n > 0xFFFFFFFFull) {
        set_error_msg(cudaErrorInvalidValue, "siCountMixedDevice: bad arguments");
        return cudaErrorInvalidValue;
    }
    siIndex* host = nullptr;   // lends its stream, its device and the table's allocation
    int rc = mixed_host(ixs, n_contigs, &host, "siCountMixedDevice: the indexes live on different devices");
    if (rc) return rc;
    if (d_totals) SIB_CHECK(cudaMemsetAsync(d_totals, 0, (size_t)n_contigs * 8, static_cast<cudaStream_t>(stream)));
    if (n == 0) return 0;
    if (!host) {   // no contig has an index: every count is 0
        SIB_CHECK(cudaMemsetAsync(d_counts, 0, n * 4, static_cast<cudaStream_t>(stream)));
        return 0;
    }
    DeviceGuard g(host->device);
    cudaStream_t s = pick_stream(host, stream);
    size_t bytes = 0;
    rc = mixed_table(host, ixs, n_contigs, nullptr, s, &bytes);
    if (rc) return rc;
    return mixed_launch(host, ixs, n_contigs, bytes, d_contig, d_qs, d_qe, n, d_counts, d_totals, QM_PLAIN, s);
}

// Mode B (one index per contig, contigs partitioned over the GPUs) WITHOUT a dispatch: every GPU's slice of the mixed batch
// stays where it is, in memory its peers have mapped (NVLink peer access / CUDA IPC). This GPU walks all n_src slices -- its
// own first, the others read in place over NVLink -- and answers exactly the queries whose contig it holds an index for,
// storing each count into the slice's own counts array (a peer store for a remote slice). foreign[k] != 0 marks a contig
// that another GPU answers; a query of a contig nobody indexes, or with an id outside [0, n_contigs), gets its 0 from
// the GPU the slice lives on (home = that slice's position in the arrays). Per query that crosses: 4 B of contig id read
// by every GPU, 8 B of coordinates read and 4 B of count written by the owner -- no all-to-all, no routing sort, no scatter.
// The caller brackets the call with a barrier between the GPUs on both sides (siPeerBarrierDevice): the slices must be
// complete before peers read them, and every GPU's stores must have landed before a slice's counts are used.
int siCountMixedPeerDevice(siIndex* const* ixs, int n_contigs, const unsigned char* foreign, int n_src, int home,
                           const int32_t* const* d_contig, const int32_t* const* d_qs, const int32_t* const* d_qe, const size_t* n,
                           uint32_t* const* d_counts, unsigned long long* d_totals, void* stream) {
    if (!ixs || n_contigs < 1 || n_contigs > (1 << 20) || n_src < 1 || n_src > 64 || home < 0 || home >= n_src || !d_contig || !d_qs ||
        !d_qe || !n || !d_counts) {
        set_error_msg(cudaErrorInvalidValue, "siCountMixedPeerDevice: bad arguments");
        return cudaErrorInvalidValue;
    }
    for (int k = 0; k < n_src; ++k)
        if (n[k] > 0xFFFFFFFFull) return cudaErrorInvalidValue;
    siIndex* host = nullptr;
    int rc = mixed_host(ixs, n_contigs, &host, "siCountMixedPeerDevice: the indexes live on different devices");
    if (rc) return rc;
    if (d_totals) SIB_CHECK(cudaMemsetAsync(d_totals, 0, (size_t)n_contigs * 8, static_cast<cudaStream_t>(stream)));
    if (!host) {
        // nothing indexed here: only the zeros of the home slice are this GPU's to write -- an id outside the table or a contig
        // nobody indexes; a foreign contig's count comes from its owner. One pass of the kernel over the home slice does that
        // with an empty table, but it needs an index to lend its stream: fall back to a memset when no contig is foreign.
        bool any_foreign = false;
        for (int k = 0; foreign && k < n_contigs; ++k) any_foreign = any_foreign || foreign[k];
        if (!any_foreign && n[home]) SIB_CHECK(cudaMemsetAsync(d_counts[home], 0, n[home] * 4, static_cast<cudaStream_t>(stream)));
        if (any_foreign) return SI_MIXED_UNSUPPORTED;   // a GPU that owns no contig: the caller keeps it out of the partition
        return 0;
    }
    DeviceGuard g(host->device);
    cudaStream_t s = pick_stream(host, stream);
    size_t bytes = 0;
    rc = mixed_table(host, ixs, n_contigs, foreign, s, &bytes);
    if (rc) return rc;
    // the home slice on the caller's stream, every remote slice on a side stream beside it: a remote walk waits on NVLink round
    // trips, the home walk on HBM gathers, so they overlap instead of queueing (SIB_PEER_SERIAL=1: one after the other)
    const bool serial = getenv("SIB_PEER_SERIAL") && atoi(getenv("SIB_PEER_SERIAL")) != 0;
    const int want_side = serial ? 0 : std::min(n_src - 1, (int)siIndex::SI_PEER_SIDE);
    while (host->peer_side_ready < want_side) {
        const int k = host->peer_side_ready;
        SIB_CHECK(cudaStreamCreateWithFlags(&host->peer_side[k], cudaStreamNonBlocking));
        SIB_CHECK(cudaEventCreateWithFlags(&host->peer_join[k], cudaEventDisableTiming));
        ++host->peer_side_ready;
    }
    if (want_side && !host->peer_fork) SIB_CHECK(cudaEventCreateWithFlags(&host->peer_fork, cudaEventDisableTiming));
    if (want_side) SIB_CHECK(cudaEventRecord(host->peer_fork, s));      // after the table upload and the zeroing of the totals
    int used = 0;
    for (int k = 0; k < n_src; ++k) {
        const int src = (home + k) % n_src;      // own slice first; the remote ones start at different peers on every GPU
        if (n[src] == 0) continue;
        cudaStream_t sk = s;
        if (k > 0 && want_side) {
            sk = host->peer_side[(k - 1) % want_side];
            if ((k - 1) < want_side) SIB_CHECK(cudaStreamWaitEvent(sk, host->peer_fork, 0));
            used = std::max(used, std::min(k, want_side));
        }
        rc = mixed_launch(host, ixs, n_contigs, bytes, d_contig[src], d_qs[src], d_qe[src], n[src], d_counts[src], d_totals,
                          src == home ? QM_PEER_HOME : QM_PEER_AWAY, sk);
        if (rc) return rc;
    }
    for (int k = 0; k < used; ++k) {
        SIB_CHECK(cudaEventRecord(host->peer_join[k], host->peer_side[k]));
        SIB_CHECK(cudaStreamWaitEvent(s, host->peer_join[k], 0));
    }
    return 0;
}

// used by c_abi.cu: 32-bit device counts widened to 64 bits on the device
int si_b200_widen_(siIndex* ix, const uint32_t* d_in, size_t n, unsigned long long* d_out, void* stream) {
    if (n == 0) return 0;
    DeviceGuard g(ix->device);
    SIB_LAUNCH(bk_widen_kernel, grid_for(n, BK_THREADS, ix->sm_count * 16), BK_THREADS, 0, static_cast<cudaStream_t>(stream), d_in, (uint64_t)n, d_out);
    return 0;
}

// ---- barrier between GPUs through peer memory --------------------------------------------------------------
// After a fan-out count every GPU has to know that the others' stores have landed before it reads their slots. One warp:
// lane k publishes `seq` in peer k's flag word for this rank (after a system-scope fence: the count kernel's peer stores
// precede it in stream order), then spins on this GPU's own flag word for peer k until it shows `seq` (or later). A peer
// that never arrives trips the timeout (about 10 s of %globaltimer) and sets *timed_out instead of hanging the GPU.
struct PeerFlags { uint32_t* signal[SI_FANOUT_MAX]; const uint32_t* wait[SI_FANOUT_MAX]; int n; };
__global__ void __launch_bounds__(32)
pk_barrier_kernel(const __grid_constant__ PeerFlags f, uint32_t seq, uint32_t* __restrict__ timed_out) {
    const int k = threadIdx.x;
    if (k >= f.n) return;
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(f.signal[k]) = seq;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int32_t)(*reinterpret_cast<const volatile uint32_t*>(f.wait[k]) - seq) < 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) { *timed_out = 1u; break; }
    }
    __threadfence_system();
}

int siPeerBarrierDevice(uint32_t* const* signal_ptrs, const uint32_t* const* wait_ptrs, int n_peers, uint32_t seq, uint32_t* d_timed_out,
                        void* stream) {
    if (n_peers < 0 || n_peers > SI_FANOUT_MAX || (n_peers && (!signal_ptrs || !wait_ptrs)) || !d_timed_out) return cudaErrorInvalidValue;
    if (n_peers == 0) return 0;
    PeerFlags f;
    f.n = n_peers;
    for (int k = 0; k < n_peers; ++k) { f.signal[k] = signal_ptrs[k]; f.wait[k] = wait_ptrs[k]; }
    pk_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(f, seq, d_timed_out);
    SIB_CHECK_LAUNCH();
    note_launch();
    return 0;
}

// ---- buffers shared between the processes of one node (CUDA IPC) --------------------------------------
int siIpcAlloc(size_t bytes, void** d_ptr, unsigned char handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    if (!d_ptr || !handle_out || bytes == 0) return cudaErrorInvalidValue;
    SIB_CHECK(cudaMalloc(d_ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) { cudaFree(*d_ptr); *d_ptr = nullptr; set_error(e, "cudaIpcGetMemHandle", __FILE__, __LINE__); return (int)e; }
    memcpy(handle_out, &h, 64);
    return 0;
}
int siIpcOpen(const unsigned char handle[64], void** d_ptr) {
    if (!d_ptr || !handle) return cudaErrorInvalidValue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    SIB_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int siIpcClose(void* d_ptr) { if (d_ptr) SIB_CHECK(cudaIpcCloseMemHandle(d_ptr)); return 0; }
int siIpcFree(void* d_ptr) { if (d_ptr) SIB_CHECK(cudaFree(d_ptr)); return 0; }

// used by c_abi.cu: resolve SI_ORDER_AUTO once for a count -> fill pair
int si_b200_resolve_order_(siIndex* ix, const int32_t* d_qs, size_t n, void* stream) {
    DeviceGuard g(ix->device);
    int o = resolve_order(ix, d_qs, (uint32_t)n, SI_ORDER_AUTO, pick_stream(ix, stream));
    if (o < 0) set_error(cudaGetLastError(), "resolve_order", __FILE__, __LINE__);
    return o;
}

// Single-query calls (c_abi.cu): one launch, the query as kernel parameters, the answer written into the handle's mapped
// pinned mailbox followed by the call's sequence number; the host spins on that word (a few microseconds) instead of a
// stream synchronise, and synchronises only if it does not show up (which also surfaces a launch or kernel error).
static int single_wait(siIndex* ix, volatile uint32_t* done, uint32_t seq) {
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 0; *done != seq; ++spins) {
        if ((spins & 1023u) == 1023u &&
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > 5.0) {
            SIB_CHECK(cudaStreamSynchronize(ix->own_stream));
            if (*done != seq) { set_error_msg(cudaErrorUnknown, "single-query kernel finished without publishing its answer"); return cudaErrorUnknown; }
            break;
        }
    }
    return 0;
}

// ---- resident single-query kernel (SI_OPT_RESIDENT_QUERIES; qk_single_server_kernel) ----------------------------------
constexpr unsigned long long SRV_IDLE_NS = 200000ull;      // leaves after 0.2 ms without a request ...
constexpr unsigned long long SRV_LIFE_NS = 2000000ull;     // ... and 2 ms after its launch at the latest (bounds what a device-wide synchronise waits)

constexpr int SINGLE_RETRY_BY_LAUNCH = -3;

static void server_post(SingleReq* req, uint32_t seq, int op, int32_t a, int32_t b, uint32_t cap) {
    req->a = a; req->b = b; req->opcap = ((uint32_t)op << 28) | (cap & 0x0FFFFFFFu);
    std::atomic_thread_fence(std::memory_order_release);                     // the query before its sequence number
    *reinterpret_cast<volatile uint32_t*>(&req->seq) = seq;
    std::atomic_thread_fence(std::memory_order_seq_cst);
}

static int server_launch(siIndex* ix, SingleReq* req, uint32_t* out32, unsigned long long* out64, void* out, uint32_t* done) {
    if (!ix->srv_stream) SIB_CHECK(cudaStreamCreateWithFlags(&ix->srv_stream, cudaStreamNonBlocking));
    ix->srv_req = req;
    ix->srv_done = done;
    *reinterpret_cast<volatile uint32_t*>(&req->alive) = 1u;
    const uint32_t last = *reinterpret_cast<volatile uint32_t*>(done);      // the last call that was answered
    SIB_LAUNCH(qk_single_server_kernel, 1, 32, 0, ix->srv_stream, view_of(ix), req, out32, out64, out, done, last, SRV_IDLE_NS, SRV_LIFE_NS);
    return 0;
}

// the resident kernel reads the index view it was launched with: it has to be gone before the index changes
int si_b200_server_stop_(siIndex* ix) {
    if (!ix || !ix->srv_stream || !ix->srv_req) return 0;
    DeviceGuard g(ix->device);
    const uint32_t seq = ++ix->single_seq;
    server_post(ix->srv_req, seq, 15, 0, 0, 0);                              // "leave now" (a kernel that has left already never reads it)
    cudaError_t e = cudaStreamSynchronize(ix->srv_stream);
    *reinterpret_cast<volatile uint32_t*>(ix->srv_done) = seq;              // so that the next launch does not take it for a request
    *reinterpret_cast<volatile uint32_t*>(&ix->srv_req->alive) = 0u;
    if (e != cudaSuccess) { set_error(e, "stopping the resident single-query kernel", __FILE__, __LINE__); return (int)e; }
    return 0;
}

static int single_via_server(siIndex* ix, SingleReq* req, int op, int32_t a, int32_t b, uint32_t cap, uint32_t* out32,
                             unsigned long long* out64, void* out, uint32_t* done) {
    const uint32_t seq = ++ix->single_seq;
    volatile uint32_t* v_done = done;
    volatile uint32_t* v_alive = &req->alive;
    if (!*v_alive) {                                                          // launch first: `last` must be the previous call
        int rc = server_launch(ix, req, out32, out64, out, done);
        if (rc) return rc;
    }
    server_post(req, seq, op, a, b, cap);
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 1; *v_done != seq; ++spins) {
        if ((spins & 255u) == 0u) {
            if (!*v_alive && *v_done != seq) {                                // it left without seeing this request: start another
                int rc = server_launch(ix, req, out32, out64, out, done);
                if (rc) return rc;
            }
            if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > 2000.0) {
                // no answer for two seconds (a saturated GPU may not have scheduled the warp yet): stop the kernel -- which waits for
                // it -- and, if the request was not served on the way out, let the caller answer it with a launch of its own
                int rc = si_b200_server_stop_(ix);
                if (rc) return rc;
                return SINGLE_RETRY_BY_LAUNCH;
            }
        }
    }
    return 0;
}

int si_b200_single_search_(siIndex* ix, int32_t qs, int32_t qe, int what, uint32_t cap, unsigned long long* found, void* out,
                           uint32_t* done, uint32_t* out32, SingleReq* req) {
    if (!ix || !ix->built) {
        set_error_msg(cudaErrorNotReady, "single-query search: index not built");
        return cudaErrorNotReady;
    }
    DeviceGuard g(ix->device);
    if (ix->resident && req) {
        const int rc = single_via_server(ix, req, 3 + what, qs, qe, cap, out32, found, out, done);
        if (rc != SINGLE_RETRY_BY_LAUNCH) return rc;
    }
    cudaStream_t s = ix->own_stream;
    const IndexView v = view_of(ix);
    const uint32_t seq = ++ix->single_seq;
    switch (what) {
        case SI_FILL_VALUES: SIB_LAUNCH((qk_single_search_kernel<FILL_VALUES>), 1, 32, 0, s, v, qs, qe, cap, found, reinterpret_cast<int32_t*>(out), done, seq); break;
        case SI_FILL_IDXS: SIB_LAUNCH((qk_single_search_kernel<FILL_IDXS>), 1, 32, 0, s, v, qs, qe, cap, found, reinterpret_cast<uint32_t*>(out), done, seq); break;
        case SI_FILL_KEYS: SIB_LAUNCH((qk_single_search_kernel<FILL_KEYS>), 1, 32, 0, s, v, qs, qe, cap, found, reinterpret_cast<int2*>(out), done, seq); break;
        default: SIB_LAUNCH((qk_single_search_kernel<FILL_ITEMS>), 1, 32, 0, s, v, qs, qe, cap, found, reinterpret_cast<Item3*>(out), done, seq); break;
    }
    return single_wait(ix, done, seq);
}

// op 0: upper_bound(a) -> *out32; 1: has_overlaps(a, b) -> *out32; 2: count(a, b) -> *out64 (all mailbox words)
int si_b200_single_scalar_(siIndex* ix, int op, int32_t a, int32_t b, uint32_t* out32, unsigned long long* out64, uint32_t* done,
                           void* out, SingleReq* req) {
    if (!ix || !ix->built) {
        set_error_msg(cudaErrorNotReady, "single query: index not built");
        return cudaErrorNotReady;
    }
    DeviceGuard g(ix->device);
    if (ix->resident && req) {
        const int rc = single_via_server(ix, req, op, a, b, 0u, out32, out64, out, done);
        if (rc != SINGLE_RETRY_BY_LAUNCH) return rc;
    }
    const uint32_t seq = ++ix->single_seq;
    SIB_LAUNCH(qk_single_scalar_kernel, 1, 32, 0, ix->own_stream, view_of(ix), op, a, b, out32, out64, done, seq);
    return single_wait(ix, done, seq);
}

}  // extern "C"
