// c_abi.cu -- the reference's C ABI (include/c_superintervals.h) and the host-buffer
// batch entry points (include/superintervals_b200.h section 2), implemented on top
// of the device-resident core in index.cu. Host code only: no kernels here.
//
// Ownership follows the reference (c_superintervals.h:363-400): the handle and its
// arrays are malloc'd by the library; result buffers are caller structs whose
// `data` the library reallocs. The handle is over-allocated: callers see the
// public cSuperIntervals prefix, the library keeps the device index behind it.
#include "../../include/superintervals_b200.h"

#include "common.cuh"
#include "index.cuh"
#include "host_common.cuh"

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

extern "C" int si_b200_single_scalar_(siIndex* ix, int op, int32_t a, int32_t b, uint32_t* out32, unsigned long long* out64, uint32_t* done,
                                      void* out, sib::SingleReq* req);
extern "C" int si_b200_single_search_(siIndex* ix, int32_t qs, int32_t qe, int what, uint32_t cap, unsigned long long* found, void* out,
                                      uint32_t* done, uint32_t* out32, sib::SingleReq* req);
extern "C" int si_b200_widen_(siIndex* ix, const uint32_t* d_in, size_t n, unsigned long long* d_out, void* stream);
extern "C" int si_b200_resolve_order_(siIndex* ix, const int32_t* d_qs, size_t n, void* stream);

namespace sib {

bool handle_ready(Handle* h, const char* who) {
    if (h && h->magic == HANDLE_MAGIC && h->indexed && h->ix) return true;
    char msg[160];
    snprintf(msg, sizeof(msg), "%s: handle not indexed (call indexSuperIntervals first)", who);
    set_error_msg(cudaErrorNotReady, msg);
    return false;
}

// ---- host side of the copies -------------------------------------------------------------
// Caller buffers are usually pageable (malloc / realloc'd result arrays, numpy). A plain
// cudaMemcpy on such memory is staged by the driver through one thread, and a freshly
// realloc'd result array additionally takes a page fault per 4 KB. Copies of more than a
// few MB therefore go through two pinned staging buffers: the DMA of chunk k+1 overlaps a
// multi-threaded memcpy of chunk k between the staging buffer and the caller's memory
// (the page faults are spread over the same threads). Pinned caller buffers are copied directly.
class HostPool {
public:
    static HostPool& get() { static HostPool p; return p; }
    int size() const { return (int)workers_.size() + 1; }
    // runs f(t, n) for t = 0..n-1, n = size(); the caller is worker 0; returns when all are done
    void run(const std::function<void(int, int)>& f) {
        std::lock_guard<std::mutex> serial(run_mu_);
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &f;
            pending_ = (int)workers_.size();
            ++gen_;
        }
        cv_work_.notify_all();
        f(0, size());
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_work_.notify_all();
        for (auto& t : workers_) t.join();
    }
private:
    HostPool() {
        unsigned hw = std::thread::hardware_concurrency();
        int n = hw >= 16 ? 8 : (hw >= 4 ? (int)hw / 2 : 1);
        if (const char* e = getenv("SIB_HOST_THREADS")) n = atoi(e) > 0 ? atoi(e) : n;
        for (int t = 1; t < n; ++t) workers_.emplace_back([this, t] { loop(t); });
    }
    void loop(int t) {
        unsigned seen = 0;
        while (true) {
            const std::function<void(int, int)>* f;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                f = fn_;
            }
            (*f)(t, size());
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) cv_done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int, int)>* fn_ = nullptr;
    unsigned gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

void copy_to_output(void* dst, const void* src, size_t n);   // host_copy.cpp: non-temporal stores

// to_output: dst is a caller's result array (written once) -> stream the stores past the caches
void parallel_memcpy(void* dst, const void* src, size_t bytes, bool to_output = false) {
    if (bytes < ((size_t)1 << 20)) { memcpy(dst, src, bytes); return; }
    HostPool::get().run([&](int t, int n) {
        const size_t per = ((bytes + n - 1) / n + 4095) & ~(size_t)4095;
        const size_t a = (size_t)t * per;
        if (a >= bytes) return;
        const size_t len = bytes - a < per ? bytes - a : per;
        if (to_output) copy_to_output((char*)dst + a, (const char*)src + a, len);
        else memcpy((char*)dst + a, (const char*)src + a, len);
    });
}

// ---- registered result buffers ------------------------------------------------------------------
namespace {
std::mutex g_reg_mu;
std::unordered_map<void*, size_t> g_registered;   // base pointer -> registered bytes
constexpr size_t REGISTER_MIN = (size_t)1 << 20;  // smaller buffers stay plain realloc memory

bool unregister_if_ours(void* p) {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_registered.find(p);
    if (it == g_registered.end()) return false;
    if (cudaHostUnregister(p) != cudaSuccess) (void)cudaGetLastError();
    g_registered.erase(it);
    return true;
}
}  // namespace

void* result_realloc(void* old, size_t old_bytes, size_t new_bytes) {
    if (new_bytes < REGISTER_MIN) {
        if (old && unregister_if_ours(old)) {      // cannot happen (buffers only grow), kept for safety
            void* q = malloc(new_bytes);
            if (q) memcpy(q, old, old_bytes < new_bytes ? old_bytes : new_bytes);
            free(old);
            return q;
        }
        return realloc(old, new_bytes);
    }
    const size_t bytes = (new_bytes + 4095) & ~(size_t)4095;
    void* q = aligned_alloc(4096, bytes);
    if (!q) return nullptr;
    if (old) {
        if (old_bytes) memcpy(q, old, old_bytes);
        unregister_if_ours(old);
        free(old);
    }
    // pins (and so first-touches) the pages; on failure the buffer simply stays pageable
    if (cudaHostRegister(q, bytes, cudaHostRegisterDefault) == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        g_registered[q] = bytes;
    } else {
        (void)cudaGetLastError();
    }
    return q;
}

void result_free(void* p) {
    if (!p) return;
    unregister_if_ours(p);
    free(p);
}

bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

constexpr size_t STAGE_BYTES = (size_t)16 << 20;   // per staging slot
constexpr size_t DIRECT_BYTES = (size_t)4 << 20;   // below this a plain cudaMemcpyAsync is as good

int ensure_stage(siIndex* ix) {
    if (ix->pinned && ix->pinned_bytes >= 2 * STAGE_BYTES) return 0;
    if (ix->pinned) { cudaFreeHost(ix->pinned); ix->pinned = nullptr; ix->pinned_bytes = 0; }
    SIB_CHECK(cudaHostAlloc(&ix->pinned, 2 * STAGE_BYTES, cudaHostAllocDefault));
    ix->pinned_bytes = 2 * STAGE_BYTES;
    if (!ix->e_stage[0]) {
        SIB_CHECK(cudaEventCreateWithFlags(&ix->e_stage[0], cudaEventDisableTiming));
        SIB_CHECK(cudaEventCreateWithFlags(&ix->e_stage[1], cudaEventDisableTiming));
    }
    return 0;
}

// device -> host, returns when the bytes are in dst
int copy_d2h(siIndex* ix, void* dst, const void* src_dev, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return 0;
    if (bytes <= DIRECT_BYTES || is_pinned(dst)) {
        SIB_CHECK(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, s));
        SIB_CHECK(cudaStreamSynchronize(s));
        return 0;
    }
    if (ensure_stage(ix)) return last_error_code();
    char* stage[2] = {(char*)ix->pinned, (char*)ix->pinned + STAGE_BYTES};
    const size_t chunks = (bytes + STAGE_BYTES - 1) / STAGE_BYTES;
    auto len = [&](size_t k) { return k + 1 < chunks ? STAGE_BYTES : bytes - k * STAGE_BYTES; };
    SIB_CHECK(cudaMemcpyAsync(stage[0], src_dev, len(0), cudaMemcpyDeviceToHost, s));
    SIB_CHECK(cudaEventRecord(ix->e_stage[0], s));
    for (size_t k = 0; k < chunks; ++k) {
        if (k + 1 < chunks) {   // slot (k+1)&1 is free: its previous contents were copied out synchronously
            SIB_CHECK(cudaMemcpyAsync(stage[(k + 1) & 1], (const char*)src_dev + (k + 1) * STAGE_BYTES, len(k + 1),
                                      cudaMemcpyDeviceToHost, s));
            SIB_CHECK(cudaEventRecord(ix->e_stage[(k + 1) & 1], s));
        }
        SIB_CHECK(cudaEventSynchronize(ix->e_stage[k & 1]));
        parallel_memcpy((char*)dst + k * STAGE_BYTES, stage[k & 1], len(k), true);
    }
    return 0;
}

// host -> device, stream-ordered on s; src may be reused when the call returns
int copy_h2d(siIndex* ix, void* dst_dev, const void* src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return 0;
    if (bytes <= DIRECT_BYTES || is_pinned(src)) {
        SIB_CHECK(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, s));
        if (bytes > DIRECT_BYTES) return 0;       // pinned: truly asynchronous, the caller synchronises before returning
        return 0;                                  // small pageable: the runtime stages it before returning
    }
    if (ensure_stage(ix)) return last_error_code();
    char* stage[2] = {(char*)ix->pinned, (char*)ix->pinned + STAGE_BYTES};
    const size_t chunks = (bytes + STAGE_BYTES - 1) / STAGE_BYTES;
    for (size_t k = 0; k < chunks; ++k) {
        const size_t l = k + 1 < chunks ? STAGE_BYTES : bytes - k * STAGE_BYTES;
        if (k >= 2) SIB_CHECK(cudaEventSynchronize(ix->e_stage[k & 1]));   // the DMA that last read this slot
        parallel_memcpy(stage[k & 1], (const char*)src + k * STAGE_BYTES, l);
        SIB_CHECK(cudaMemcpyAsync((char*)dst_dev + k * STAGE_BYTES, stage[k & 1], l, cudaMemcpyHostToDevice, s));
        SIB_CHECK(cudaEventRecord(ix->e_stage[k & 1], s));
    }
    // the staging slots may be reused by the next copy only after these DMAs have read them
    SIB_CHECK(cudaEventSynchronize(ix->e_stage[(chunks - 1) & 1]));
    if (chunks > 1) SIB_CHECK(cudaEventSynchronize(ix->e_stage[(chunks - 2) & 1]));
    return 0;
}

// upload a query batch into the index's staging buffers
int stage_queries(siIndex* ix, const int32_t* qs, const int32_t* qe, size_t n) {
    if (ix->h_qs.ensure(n * 4) || ix->h_qe.ensure(n * 4)) return last_error_code();
    int rc = copy_h2d(ix, ix->h_qs.p, qs, n * 4, ix->own_stream);
    if (rc) return rc;
    return copy_h2d(ix, ix->h_qe.p, qe, n * 4, ix->own_stream);
}

// ---- single-query mailbox ------------------------------------------------------------------------------
// The reference's C ABI is one query per call (c.h:537-821) and its drivers loop over it (test/bench.cpp:219-222,
// 240-242): such a call is pure latency. Its query and its answer therefore live in one block of MAPPED PINNED host
// memory that the kernels address directly -- no cudaMemcpy in either direction. By default one resident warp polls the
// request record (SI_OPT_RESIDENT_QUERIES, index.cu / qk_single_server_kernel); with the option off every call is one
// launch of a one-warp kernel. Either way the kernel publishes the call's sequence number after its answer and the
// host spins on that word. coverage() still runs the batch kernel on the mailbox (n = 1).
struct Mailbox {
    int32_t qs, qe;                       // the query, read by the kernels over PCIe
    uint32_t ub, any;                     // upperBound / anyOverlaps answers
    unsigned long long count;             // countOverlaps (64-bit count kernels), search: hits found
    uint32_t cov_count; int32_t cov;      // coverage
    uint32_t done;                        // sequence number of the last call whose answer is complete (written last by its kernel)
    SingleReq req;                        // request block of the resident kernel (SI_OPT_RESIDENT_QUERIES)
    alignas(32) unsigned char out[1];     // search results (MAILBOX_OUT_BYTES)
};
constexpr size_t MAILBOX_OUT_BYTES = (size_t)192 << 10;   // 16 K Interval records / 48 K values; longer lists take the batch path
Mailbox* mailbox_of(siIndex* ix) {
    if (!ix->mailbox) {
        if (cudaHostAlloc(&ix->mailbox, sizeof(Mailbox) + MAILBOX_OUT_BYTES, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            set_error(cudaGetLastError(), "cudaHostAlloc(mailbox)", __FILE__, __LINE__);
            ix->mailbox = nullptr;
            return nullptr;
        }
        memset(ix->mailbox, 0, sizeof(Mailbox));
    }
    return reinterpret_cast<Mailbox*>(ix->mailbox);
}

// one query through the single-search kernel; appended to `found` like the batch call. Returns -1 when the list is
// longer than the mailbox (the caller then takes the batch path), 0 on success, else an error code.
template <typename R>
int search_single(Handle* h, int32_t qs, int32_t qe, R* found, int what, size_t elem) {
    siIndex* ix = h->ix;
    std::lock_guard<std::mutex> lk(ix->api_mu);
    Mailbox* mb = mailbox_of(ix);
    if (!mb) return last_error_code();
    const uint32_t cap = (uint32_t)(MAILBOX_OUT_BYTES / elem);
    int rc = si_b200_single_search_(ix, qs, qe, what, cap, &mb->count, mb->out, &mb->done, &mb->ub, &mb->req);
    if (rc) return rc;
    const size_t total = (size_t)*reinterpret_cast<volatile unsigned long long*>(&mb->count);
    if (total > cap) return -1;
    if (total == 0) return 0;
    if (!grow(found, found->size + total, elem)) return cudaErrorMemoryAllocation;
    memcpy(reinterpret_cast<char*>(found->data) + found->size * elem, mb->out, total * elem);
    found->size += total;
    return 0;
}

// count -> scan -> (host learns total) -> grow -> fill -> copy back, appended to `found`
template <typename R>
int search_batch(Handle* h, const int32_t* qs, const int32_t* qe, size_t n, size_t* offsets_out, R* found,
                 int what, size_t elem) {
    siIndex* ix = h->ix;
    if (n == 0) { if (offsets_out) offsets_out[0] = 0; return 0; }
    std::lock_guard<std::mutex> lk(ix->api_mu);
    int rc = stage_queries(ix, qs, qe, n);
    if (rc) return rc;
    if (ix->h_counts.ensure(n * 4 + 64) || ix->h_offsets.ensure((n + 1) * 8 + 64)) return last_error_code();
    const int32_t* dqs = ix->h_qs.as<int32_t>();
    const int32_t* dqe = ix->h_qe.as<int32_t>();
    // resolve the query order once so that count and fill agree and share one sort
    const int order = si_b200_resolve_order_(ix, dqs, n, ix->own_stream);
    if (order < 0) return last_error_code();
    rc = siCountDevice(ix, dqs, dqe, n, ix->h_counts.as<uint32_t>(), order, ix->own_stream);
    if (rc) return rc;
    rc = siScanDevice(ix, ix->h_counts.as<uint32_t>(), n, ix->h_offsets.as<uint64_t>(), ix->own_stream);
    if (rc) return rc;
    static_assert(sizeof(size_t) == sizeof(uint64_t), "LP64 only");
    uint64_t total = 0;
    SIB_CHECK(cudaMemcpyAsync(&total, ix->h_offsets.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, ix->own_stream));
    SIB_CHECK(cudaStreamSynchronize(ix->own_stream));
    if (total) {
        if (!grow(found, found->size + total, elem)) return cudaErrorMemoryAllocation;
        if (ix->h_out.ensure(total * elem)) return last_error_code();
        // the fill runs while the offsets travel
        rc = siFillDevice(ix, dqs, dqe, n, ix->h_offsets.as<uint64_t>(), what, ix->h_out.p, order, ix->own_stream);
        if (rc) return rc;
    }
    // a pinned result array (result_realloc registers the large ones) takes its DMA now: the values travel while the host
    // stages the offsets below, instead of after it
    char* vdst = total ? reinterpret_cast<char*>(found->data) + found->size * elem : nullptr;
    const bool vals_async = total && total * elem > DIRECT_BYTES && is_pinned(vdst);
    if (vals_async) SIB_CHECK(cudaMemcpyAsync(vdst, ix->h_out.p, total * elem, cudaMemcpyDeviceToHost, ix->own_stream));
    if (offsets_out) {
        if (!ix->pipe_ready_out) {
            SIB_CHECK(cudaStreamCreateWithFlags(&ix->s_out2, cudaStreamNonBlocking));
            ix->pipe_ready_out = true;
        }
        // offsets were final before the fill was launched: copy them on a second stream meanwhile
        rc = copy_d2h(ix, offsets_out, ix->h_offsets.p, (n + 1) * 8, ix->s_out2);
        if (rc) return rc;
    }
    if (total == 0) return 0;
    if (vals_async) SIB_CHECK(cudaStreamSynchronize(ix->own_stream));
    else {
        rc = copy_d2h(ix, vdst, ix->h_out.p, total * elem, ix->own_stream);
        if (rc) return rc;
    }
    found->size += total;
    return 0;
}

}  // namespace sib

using namespace sib;

extern "C" {

// ---- lifecycle (ref c_superintervals.h:363-423) ---------------------------------------
cSuperIntervals* createSuperIntervals(void) {
    Handle* h = (Handle*)calloc(1, sizeof(Handle));
    if (!h) return nullptr;
    h->pub.startSorted = true;
    h->pub.endSorted = true;
    h->magic = HANDLE_MAGIC;
    h->mirror = true;
    return &h->pub;
}

void destroySuperIntervals(cSuperIntervals* si) {
    if (!si) return;   // ref:378
    Handle* h = H(si);
    free(si->starts);
    free(si->ends);
    free(si->data);
    free(si->branch);
    if (h->magic == HANDLE_MAGIC && h->ix) siIndexDestroy(h->ix);
    h->magic = 0;
    free(h);
}

void clearSuperIntervals(cSuperIntervals* si) {   // keeps capacity, ref:386-391
    si->size = 0;
    si->idx = 0;
    si->startSorted = true;
    si->endSorted = true;
    H(si)->indexed = false;
}

void reserveSuperIntervals(cSuperIntervals* si, size_t n) {
    if (n <= si->capacity) return;
    si->capacity = n;
    si->starts = (int32_t*)realloc(si->starts, n * sizeof(int32_t));
    si->ends = (int32_t*)realloc(si->ends, n * sizeof(int32_t));
    si->data = (int32_t*)realloc(si->data, n * sizeof(int32_t));
}

void addInterval(cSuperIntervals* si, int32_t start, int32_t end, int32_t value) {
    if (si->size >= si->capacity) reserveSuperIntervals(si, si->capacity ? si->capacity * 2 : 1);
    // sortedness bookkeeping exactly as ref:408-413
    if (si->startSorted && si->size > 0) {
        const int32_t ps = si->starts[si->size - 1];
        if (start < ps) si->startSorted = false;
        else if (start == ps && end > si->ends[si->size - 1]) si->endSorted = false;
    }
    si->starts[si->size] = start;
    si->ends[si->size] = end;
    si->data[si->size] = value;
    si->size++;
}

void addIntervals(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, const int32_t* values, size_t n) {
    if (n == 0) return;
    if (si->size + n > si->capacity) {
        size_t cap = si->capacity ? si->capacity : 1;
        while (cap < si->size + n) cap *= 2;
        reserveSuperIntervals(si, cap);
    }
    size_t at = si->size;
    for (size_t i = 0; i < n; ++i, ++at) {
        if (si->startSorted && at > 0) {
            const int32_t ps = si->starts[at - 1];
            if (starts[i] < ps) si->startSorted = false;
            else if (starts[i] == ps && ends[i] > si->ends[at - 1]) si->endSorted = false;
        }
        si->starts[at] = starts[i];
        si->ends[at] = ends[i];
        si->data[at] = values ? values[i] : (int32_t)at;
    }
    si->size = at;
}

size_t sizeSuperIntervals(const cSuperIntervals* si) { return si->size; }

void siSetHostMirror(cSuperIntervals* si, bool enabled) { H(si)->mirror = enabled; }
siIndex* siIndexOf(cSuperIntervals* si) {
    Handle* h = H(si);
    return (h && h->magic == HANDLE_MAGIC && h->indexed) ? h->ix : nullptr;
}

// ---- indexing (ref:462-521) ---------------------------------------------------------------
static int build_from_handle(Handle* h) {
    cSuperIntervals* si = &h->pub;
    if (!h->ix) {
        h->ix = siIndexCreate();
        if (!h->ix) return last_error_code();
    }
    return siIndexBuildHost(h->ix, si->starts, si->ends, si->data, si->size);
}

void sortIntervals(cSuperIntervals* si) {
    Handle* h = H(si);
    if (si->size == 0) return;
    if (si->startSorted && si->endSorted) return;   // ref:463-484: nothing to do
    if (build_from_handle(h)) return;
    if (siIndexExport(h->ix, si->starts, si->ends, si->data, nullptr, nullptr)) return;
    si->startSorted = true;
    si->endSorted = true;
}

void indexSuperIntervals(cSuperIntervals* si) {
    Handle* h = H(si);
    if (si->size == 0) return;   // ref:488-490
    h->indexed = false;
    if (build_from_handle(h)) return;
    if (h->mirror) {
        size_t* b = (size_t*)realloc(si->branch, si->size * sizeof(size_t));
        if (!b) { set_error_msg(cudaErrorMemoryAllocation, "realloc(branch) failed"); return; }
        si->branch = b;
        if (siIndexExport(h->ix, si->starts, si->ends, si->data, si->branch, nullptr)) return;
        si->startSorted = true;
        si->endSorted = true;
    }
    si->idx = 0;
    h->indexed = true;
}

// ---- element access (ref:523-535): host mirrors ------------------------------------------
bool intervalAt(const cSuperIntervals* si, size_t index, Interval* out) {
    if (index >= si->size) return false;
    out->start = si->starts[index];
    out->end = si->ends[index];
    out->data = si->data[index];
    return true;
}
int32_t startAt(const cSuperIntervals* si, size_t index) { return si->starts[index]; }
int32_t endAt(const cSuperIntervals* si, size_t index) { return si->ends[index]; }
int32_t dataAt(const cSuperIntervals* si, size_t index) { return si->data[index]; }

}  // extern "C"

// ---- batch queries over host buffers ---------------------------------------------------------
// Large batches are pipelined in chunks over three streams so the PCIe copies of chunk
// k+1 (H2D) and k-1 (D2H) overlap the kernels of chunk k; two buffer slots.
template <typename CountT>
static int count_dev(siIndex* ix, const int32_t* dqs, const int32_t* dqe, size_t n, CountT* dc, int order, cudaStream_t s);
template <>
int count_dev<uint64_t>(siIndex* ix, const int32_t* dqs, const int32_t* dqe, size_t n, uint64_t* dc, int order, cudaStream_t s) {
    return siCountDevice64(ix, dqs, dqe, n, dc, order, s);
}
template <>
int count_dev<uint32_t>(siIndex* ix, const int32_t* dqs, const int32_t* dqe, size_t n, uint32_t* dc, int order, cudaStream_t s) {
    return siCountDevice(ix, dqs, dqe, n, dc, order, s);
}

template <typename CountT>
static int count_batch_pipelined(siIndex* ix, const int32_t* qs, const int32_t* qe, size_t n, CountT* counts_out) {
    // 8 M queries per chunk (64 MB in, 32-64 MB out) on two device slots. Measured on C2 (100 M queries, bare-copy ceiling 16.2-16.6 ms,
    // tools/gpu_r02zl.sh): 8 M x 2 slots 18.3 ms, 8 M x 4 slots 18.5 ms, 4 M x 4 18.3 ms, 2 M x 4 19.3 ms, 16 M x 3 19.4 ms -- smaller
    // chunks and uploads running further ahead buy nothing (SIB_PIPE_CHUNK / SIB_PIPE_SLOTS override).
    static const size_t CHUNK = [] {
        size_t c = (size_t)8 << 20;
        if (const char* e = getenv("SIB_PIPE_CHUNK")) { const long long v = atoll(e); if (v >= 65536) c = (size_t)v; }
        return c;
    }();
    static const int SLOTS = [] {
        int k = 2;
        if (const char* e = getenv("SIB_PIPE_SLOTS")) { const int v = atoi(e); if (v >= 2 && v <= SI_PIPE_SLOTS) k = v; }
        return k;
    }();
    if (!ix->pipe_ready) {
        SIB_CHECK(cudaStreamCreateWithFlags(&ix->s_in, cudaStreamNonBlocking));
        SIB_CHECK(cudaStreamCreateWithFlags(&ix->s_out, cudaStreamNonBlocking));
        for (int k = 0; k < SI_PIPE_SLOTS; ++k) {
            SIB_CHECK(cudaEventCreateWithFlags(&ix->e_in[k], cudaEventDisableTiming));
            SIB_CHECK(cudaEventCreateWithFlags(&ix->e_k[k], cudaEventDisableTiming));
            SIB_CHECK(cudaEventCreateWithFlags(&ix->e_out[k], cudaEventDisableTiming));
        }
        ix->pipe_ready = true;
    }
    if (ix->h_qs.ensure(SLOTS * CHUNK * 4) || ix->h_qe.ensure(SLOTS * CHUNK * 4) || ix->h_counts.ensure(SLOTS * CHUNK * sizeof(CountT)))
        return last_error_code();
    cudaStream_t s_in = ix->s_in, s_k = ix->own_stream, s_out = ix->s_out;
    const size_t nchunks = (n + CHUNK - 1) / CHUNK;
    auto len = [&](size_t j) { return n - j * CHUNK < CHUNK ? n - j * CHUNK : CHUNK; };
    // upload of chunk j into its slot, once the count that last read the slot (chunk j - SLOTS) is done
    auto upload = [&](size_t j) -> int {
        const int slot = (int)(j % SLOTS);
        if (j >= (size_t)SLOTS) SIB_CHECK(cudaStreamWaitEvent(s_in, ix->e_k[slot], 0));
        SIB_CHECK(cudaMemcpyAsync(ix->h_qs.as<int32_t>() + slot * CHUNK, qs + j * CHUNK, len(j) * 4, cudaMemcpyHostToDevice, s_in));
        SIB_CHECK(cudaMemcpyAsync(ix->h_qe.as<int32_t>() + slot * CHUNK, qe + j * CHUNK, len(j) * 4, cudaMemcpyHostToDevice, s_in));
        SIB_CHECK(cudaEventRecord(ix->e_in[slot], s_in));
        return 0;
    };
    int order = SI_ORDER_AUTO;
    size_t uploaded = 0;
    for (size_t k = 0; k < nchunks; ++k) {
        // uploads run ahead of the counts by up to SLOTS chunks (chunk j - SLOTS < k has had its count enqueued)
        while (uploaded < nchunks && uploaded < k + SLOTS)
            if (upload(uploaded++)) return last_error_code();
        const int slot = (int)(k % SLOTS);
        const size_t m = len(k);
        int32_t* dqs = ix->h_qs.as<int32_t>() + slot * CHUNK;
        int32_t* dqe = ix->h_qe.as<int32_t>() + slot * CHUNK;
        CountT* dc = ix->h_counts.as<CountT>() + slot * CHUNK;
        SIB_CHECK(cudaStreamWaitEvent(s_k, ix->e_in[slot], 0));
        if (k >= (size_t)SLOTS) SIB_CHECK(cudaStreamWaitEvent(s_k, ix->e_out[slot], 0));   // slot's counts drained (chunk k - SLOTS)
        if (k == 0) {
            // one device check on the first chunk decides sort-or-not for the whole batch; any
            // order is answered correctly, the choice only affects locality
            order = si_b200_resolve_order_(ix, dqs, m, s_k);
            if (order < 0) return last_error_code();
        }
        int rc = count_dev<CountT>(ix, dqs, dqe, m, dc, order, s_k);
        if (rc) return rc;
        SIB_CHECK(cudaEventRecord(ix->e_k[slot], s_k));
        SIB_CHECK(cudaStreamWaitEvent(s_out, ix->e_k[slot], 0));
        SIB_CHECK(cudaMemcpyAsync(counts_out + k * CHUNK, dc, m * sizeof(CountT), cudaMemcpyDeviceToHost, s_out));
        SIB_CHECK(cudaEventRecord(ix->e_out[slot], s_out));
    }
    SIB_CHECK(cudaStreamSynchronize(s_out));
    SIB_CHECK(cudaStreamSynchronize(s_k));
    return 0;
}

template <typename CountT>
static void count_batch_host(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n, CountT* counts_out, const char* who) {
    Handle* h = H(si);
    if (n == 0) return;
    if (si->size == 0) { memset(counts_out, 0, n * sizeof(CountT)); return; }   // ref:730-732
    if (!handle_ready(h, who)) { memset(counts_out, 0, n * sizeof(CountT)); return; }
    siIndex* ix = h->ix;
    std::lock_guard<std::mutex> lk(ix->api_mu);
    // chunked pipeline: always for large batches; from 4 M queries when the caller's buffers are pinned (pageable ones are
    // staged faster by stage_queries' own threads than by the runtime)
    if (n > ((size_t)12 << 20) || (n > ((size_t)4 << 20) && is_pinned(starts) && is_pinned(ends) && is_pinned(counts_out))) {
        count_batch_pipelined<CountT>(ix, starts, ends, n, counts_out);
        return;
    }
    if (stage_queries(ix, starts, ends, n) || ix->h_counts.ensure(n * sizeof(CountT))) return;
    if (count_dev<CountT>(ix, ix->h_qs.as<int32_t>(), ix->h_qe.as<int32_t>(), n, ix->h_counts.as<CountT>(), SI_ORDER_AUTO, ix->own_stream))
        return;
    if (cudaMemcpyAsync(counts_out, ix->h_counts.p, n * sizeof(CountT), cudaMemcpyDeviceToHost, ix->own_stream) != cudaSuccess ||
        cudaStreamSynchronize(ix->own_stream) != cudaSuccess)
        set_error(cudaGetLastError(), who, __FILE__, __LINE__);
}

extern "C" {

void countOverlapsBatch(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n,
                        size_t* counts_out) {
    static_assert(sizeof(size_t) == 8, "LP64 only");
    count_batch_host<uint64_t>(si, starts, ends, n, reinterpret_cast<uint64_t*>(counts_out), "countOverlapsBatch");
}
// the same with 32-bit counts (a count never exceeds the number of stored intervals, < 2^32): half the D2H bytes
void countOverlapsBatch32(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n,
                          uint32_t* counts_out) {
    count_batch_host<uint32_t>(si, starts, ends, n, counts_out, "countOverlapsBatch32");
}

void anyOverlapsBatch(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n, bool* out) {
    Handle* h = H(si);
    if (n == 0) return;
    if (si->size == 0) { memset(out, 0, n); return; }
    if (!handle_ready(h, "anyOverlapsBatch")) { memset(out, 0, n); return; }
    siIndex* ix = h->ix;
    std::lock_guard<std::mutex> lk(ix->api_mu);
    if (stage_queries(ix, starts, ends, n) || ix->h_counts.ensure(n)) return;
    if (siAnyDevice(ix, ix->h_qs.as<int32_t>(), ix->h_qe.as<int32_t>(), n, ix->h_counts.as<uint8_t>(), ix->own_stream))
        return;
    static_assert(sizeof(bool) == 1, "bool must be one byte");
    if (cudaMemcpyAsync(out, ix->h_counts.p, n, cudaMemcpyDeviceToHost, ix->own_stream) != cudaSuccess ||
        cudaStreamSynchronize(ix->own_stream) != cudaSuccess)
        set_error(cudaGetLastError(), "anyOverlapsBatch copy-back", __FILE__, __LINE__);
}

#define SIB_SEARCH_BATCH(NAME, RTYPE, WHAT, ELEM)                                                         \
    void NAME(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n, size_t* offsets_out, \
              RTYPE* found) {                                                                             \
        Handle* h = H(si);                                                                                \
        if (si->size == 0 || !handle_ready(h, #NAME)) {                                                          \
            if (offsets_out) memset(offsets_out, 0, (n + 1) * sizeof(size_t));                            \
            return;                                                                                       \
        }                                                                                                 \
        search_batch(h, starts, ends, n, offsets_out, found, WHAT, ELEM);                                 \
    }
SIB_SEARCH_BATCH(searchValuesBatch, cIndexResult, SI_FILL_VALUES, sizeof(int32_t))
SIB_SEARCH_BATCH(searchIdxsBatch, cIndexResult, SI_FILL_IDXS, sizeof(int32_t))
SIB_SEARCH_BATCH(searchKeysBatch, cKeyResult, SI_FILL_KEYS, sizeof(KeyPair))
SIB_SEARCH_BATCH(searchItemsBatch, cItemResult, SI_FILL_ITEMS, sizeof(Interval))
#undef SIB_SEARCH_BATCH

void coverageBatch(cSuperIntervals* si, const int32_t* starts, const int32_t* ends, size_t n, size_t* count_out,
                   int32_t* coverage_out) {
    Handle* h = H(si);
    if (n == 0) return;
    if (si->size == 0 || !handle_ready(h, "coverageBatch")) {
        memset(count_out, 0, n * sizeof(size_t));
        memset(coverage_out, 0, n * sizeof(int32_t));
        return;
    }
    siIndex* ix = h->ix;
    std::lock_guard<std::mutex> lk(ix->api_mu);
    if (stage_queries(ix, starts, ends, n) || ix->h_counts.ensure(n * 4) || ix->h_cov.ensure(n * 4) || ix->h_out.ensure(n * 8)) return;
    if (siCoverageDevice(ix, ix->h_qs.as<int32_t>(), ix->h_qe.as<int32_t>(), n, ix->h_counts.as<uint32_t>(),
                         ix->h_cov.as<int32_t>(), ix->own_stream))
        return;
    // the counts are widened to size_t on the device and land in the caller's array directly
    static_assert(sizeof(size_t) == sizeof(unsigned long long), "LP64 only");
    if (si_b200_widen_(ix, ix->h_counts.as<uint32_t>(), n, ix->h_out.as<unsigned long long>(), ix->own_stream)) return;
    if (cudaMemcpyAsync(count_out, ix->h_out.p, n * 8, cudaMemcpyDeviceToHost, ix->own_stream) != cudaSuccess ||
        cudaMemcpyAsync(coverage_out, ix->h_cov.p, n * 4, cudaMemcpyDeviceToHost, ix->own_stream) != cudaSuccess ||
        cudaStreamSynchronize(ix->own_stream) != cudaSuccess)
        set_error(cudaGetLastError(), "coverageBatch copy-back", __FILE__, __LINE__);
}

// ---- single queries (ref:537-821): one launch through the mapped mailbox, no CPU fallback ----------------
size_t upperBound(cSuperIntervals* si, int32_t value) {
    Handle* h = H(si);
    size_t r = SI_NONE;
    if (si->size != 0 && handle_ready(h, "upperBound")) {
        std::lock_guard<std::mutex> lk(h->ix->api_mu);
        if (h->ix->n != 0)
            if (Mailbox* mb = mailbox_of(h->ix))
                if (si_b200_single_scalar_(h->ix, 0, value, 0, &mb->ub, &mb->count, &mb->done, mb->out, &mb->req) == 0) {
                    const uint32_t u = *reinterpret_cast<volatile uint32_t*>(&mb->ub);
                    r = u == 0xFFFFFFFFu ? SI_NONE : (size_t)u;
                }
    }
    si->idx = r;   // ref:539,562,566
    return r;
}

bool anyOverlaps(cSuperIntervals* si, int32_t start, int32_t end) {
    Handle* h = H(si);
    if (si->size == 0 || !handle_ready(h, "anyOverlaps")) return false;
    siIndex* ix = h->ix;
    std::lock_guard<std::mutex> lk(ix->api_mu);
    Mailbox* mb = mailbox_of(ix);
    if (!mb) return false;
    if (si_b200_single_scalar_(ix, 1, start, end, &mb->ub, &mb->count, &mb->done, mb->out, &mb->req)) return false;
    return *reinterpret_cast<volatile uint32_t*>(&mb->ub) != 0;
}

size_t countOverlaps(cSuperIntervals* si, int32_t start, int32_t end) {
    Handle* h = H(si);
    if (si->size == 0 || !handle_ready(h, "countOverlaps")) return 0;   // ref:730-732
    siIndex* ix = h->ix;
    std::lock_guard<std::mutex> lk(ix->api_mu);
    Mailbox* mb = mailbox_of(ix);
    if (!mb) return 0;
    if (si_b200_single_scalar_(ix, 2, start, end, &mb->ub, &mb->count, &mb->done, mb->out, &mb->req)) return 0;
    return (size_t)*reinterpret_cast<volatile unsigned long long*>(&mb->count);
}

#define SIB_SEARCH_SINGLE(NAME, BATCH, RTYPE, WHAT, ELEM)                                           \
    void NAME(cSuperIntervals* si, int32_t start, int32_t end, RTYPE* found) {                      \
        Handle* h = H(si);                                                                          \
        if (si->size == 0 || !handle_ready(h, #NAME)) return;                                       \
        if (search_single(h, start, end, found, WHAT, ELEM) == -1)                                  \
            BATCH(si, &start, &end, 1, nullptr, found);   /* longer than the mailbox */             \
    }
SIB_SEARCH_SINGLE(searchValues, searchValuesBatch, cIndexResult, SI_FILL_VALUES, sizeof(int32_t))
SIB_SEARCH_SINGLE(searchIdxs, searchIdxsBatch, cIndexResult, SI_FILL_IDXS, sizeof(int32_t))
SIB_SEARCH_SINGLE(searchKeys, searchKeysBatch, cKeyResult, SI_FILL_KEYS, sizeof(KeyPair))
SIB_SEARCH_SINGLE(searchItems, searchItemsBatch, cItemResult, SI_FILL_ITEMS, sizeof(Interval))
#undef SIB_SEARCH_SINGLE
void searchPoint(cSuperIntervals* si, int32_t point, cIndexResult* found) {   // ref:725-727
    searchValues(si, point, point, found);
}
void coverage(cSuperIntervals* si, int32_t start, int32_t end, size_t* count_out, int32_t* coverage_out) {
    *count_out = 0;
    *coverage_out = 0;
    Handle* h = H(si);
    if (si->size == 0 || !handle_ready(h, "coverage")) return;
    siIndex* ix = h->ix;
    std::lock_guard<std::mutex> lk(ix->api_mu);
    Mailbox* mb = mailbox_of(ix);
    if (!mb) return;
    mb->qs = start; mb->qe = end;
    if (siCoverageDevice(ix, &mb->qs, &mb->qe, 1, &mb->cov_count, &mb->cov, ix->own_stream)) return;
    if (cudaStreamSynchronize(ix->own_stream) != cudaSuccess) { set_error(cudaGetLastError(), "coverage", __FILE__, __LINE__); return; }
    *count_out = (size_t)*reinterpret_cast<volatile uint32_t*>(&mb->cov_count);
    *coverage_out = *reinterpret_cast<volatile int32_t*>(&mb->cov);
}
void findOverlaps(cSuperIntervals* si, int32_t start, int32_t end, int32_t* found, size_t* found_size) {
    // legacy raw-buffer form (ref:794-821): the caller sized `found`
    cIndexResult tmp = {nullptr, 0, 0};
    searchValues(si, start, end, &tmp);
    if (tmp.size) memcpy(found, tmp.data, tmp.size * sizeof(int32_t));
    *found_size = tmp.size;
    result_free(tmp.data);
}

// ---- result buffers (ref:1066-1089) -----------------------------------------------------------
cIndexResult createIndexResult(void) { cIndexResult r = {nullptr, 0, 0}; return r; }
void clearIndexResult(cIndexResult* r) { r->size = 0; }
void destroyIndexResult(cIndexResult* r) { result_free(r->data); r->data = nullptr; r->size = r->capacity = 0; }
cKeyResult createKeyResult(void) { cKeyResult r = {nullptr, 0, 0}; return r; }
void clearKeyResult(cKeyResult* r) { r->size = 0; }
void destroyKeyResult(cKeyResult* r) { result_free(r->data); r->data = nullptr; r->size = r->capacity = 0; }
cItemResult createItemResult(void) { cItemResult r = {nullptr, 0, 0}; return r; }
void clearItemResult(cItemResult* r) { r->size = 0; }
void destroyItemResult(cItemResult* r) { result_free(r->data); r->data = nullptr; r->size = r->capacity = 0; }

}  // extern "C"
