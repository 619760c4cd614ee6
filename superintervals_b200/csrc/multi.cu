// multi.cu -- several GPUs of one node behind the same C ABI (include/superintervals_b200.h section 4).
//
// one host process, one index replica and one stream per device. The reference has no
// notion of devices; its callers keep one map per chromosome (examples/bed-intersect-si.rs:100-123)
// and loop over queries. Here a host batch is cut into one contiguous range per device:
//   build    the intervals go to the first device once (one H2D), ncclBroadcast carries them to the
//            other devices over NVLink, every device builds its replica from device memory;
//   count    H2D of each range on its own PCIe link, one count launch per device. With NVLink peer access the
//            kernel itself stores each count into every device's gathered vector (fan-out: count and all-gather are
//            one kernel, the streams then wait on each other's events); otherwise ONE ncclAllGather of the
//            per-query counts (4 B x range per rank). Every device then holds the whole count vector and
//            returns its own range to the host;
//   search   the gathered counts are scanned on every device into the GLOBAL 64-bit CSR offsets
//            (no host round trip for the bases), each device fills and returns its own segment.
// NCCL is resolved at run time (dlopen of libnccl.so.2): a single-device process never needs it.
#include "../../include/superintervals_b200.h"

#include "common.cuh"
#include "host_common.cuh"
#include "index.cuh"

#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace sib {
namespace {

// ---- the few NCCL entry points used, resolved from libnccl.so.2 ---------------------------------------
typedef struct ncclComm* ncclComm_t;
enum { NCCL_INT32 = 2, NCCL_UINT32 = 3 };   // ncclDataType_t values (nccl.h: ncclInt32 = 2, ncclUint32 = 3)
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
#define SIB_SYM(field, sym) *(void**)(&field) = dlsym(lib, sym)
        SIB_SYM(CommInitAll, "ncclCommInitAll");
        SIB_SYM(CommDestroy, "ncclCommDestroy");
        SIB_SYM(AllGather, "ncclAllGather");
        SIB_SYM(Broadcast, "ncclBroadcast");
        SIB_SYM(GroupStart, "ncclGroupStart");
        SIB_SYM(GroupEnd, "ncclGroupEnd");
        SIB_SYM(GetVersion, "ncclGetVersion");
        SIB_SYM(GetErrorString, "ncclGetErrorString");
#undef SIB_SYM
        return CommInitAll && CommDestroy && AllGather && Broadcast && GroupStart && GroupEnd;
    }
};

}  // namespace
}  // namespace sib

using namespace sib;

struct siMulti {
    int n = 0;
    std::vector<int> dev;
    std::vector<siIndex*> ix;
    std::vector<cudaStream_t> st;
    std::vector<DevBuf> in_s, in_e, in_v;      // build inputs per device
    std::vector<DevBuf> qs, qe, gather, offsets, out;
    std::vector<cudaEvent_t> ev;               // 5 per device: start, h2d done, count done, gather done, d2h done
    NcclApi nccl;
    std::vector<ncclComm_t> comm;
    bool p2p = false;                          // every device can store into every other's memory (NVLink peer access enabled)
    std::vector<cudaEvent_t> ev_done;          // per device: its fan-out count kernel has finished
    size_t per = 0;                            // queries per device of the last batch (padded range length)
    size_t n_intervals = 0;
    bool built = false;
    siMultiStats stats = {};
    std::mutex mu;
};

namespace {

#define SIB_NCCL(m, expr)                                                                        \
    do {                                                                                         \
        int r_ = (expr);                                                                         \
        if (r_ != 0) {                                                                           \
            char msg_[256];                                                                      \
            snprintf(msg_, sizeof(msg_), "NCCL error %d (%s) in %s", r_,                         \
                     (m)->nccl.GetErrorString ? (m)->nccl.GetErrorString(r_) : "?", #expr);      \
            set_error_msg(cudaErrorUnknown, msg_);                                               \
            return cudaErrorUnknown;                                                             \
        }                                                                                        \
    } while (0)

// run f(i) for every device on its own host thread (pageable copies and the builds' host syncs then overlap)
template <typename F>
int for_each_device(siMulti* m, F f) {
    std::vector<int> rc(m->n, 0);
    std::vector<std::thread> th;
    for (int i = 1; i < m->n; ++i)
        th.emplace_back([&, i] { cudaSetDevice(m->dev[i]); rc[i] = f(i); });
    cudaSetDevice(m->dev[0]);
    rc[0] = f(0);
    for (auto& t : th) t.join();
    for (int i = 0; i < m->n; ++i)
        if (rc[i]) return rc[i];
    return 0;
}

int sync_all(siMulti* m) {
    for (int i = 0; i < m->n; ++i) {
        SIB_CHECK(cudaSetDevice(m->dev[i]));
        SIB_CHECK(cudaStreamSynchronize(m->st[i]));
    }
    return 0;
}

float span_ms(siMulti* m, int a, int b) {   // max over devices of event b - event a
    float worst = 0.f;
    for (int i = 0; i < m->n; ++i) {
        float ms = 0.f;
        cudaSetDevice(m->dev[i]);
        if (cudaEventElapsedTime(&ms, m->ev[5 * i + a], m->ev[5 * i + b]) == cudaSuccess && ms > worst) worst = ms;
    }
    (void)cudaGetLastError();
    return worst;
}

// count every range and gather: on return (stream-ordered) every device holds all counts in gather[i].
// With peer access between the devices (NVLink) the gather is FUSED into the count: device i's kernel stores each count into
// its slot of every device's gathered vector as it is produced (siCountFanoutDevice), and the streams then wait for one
// another's kernels through events -- no collective pass. Without peer access: count, then ncclAllGather in place.
int count_ranges(siMulti* m, const int32_t* qs, const int32_t* qe, size_t nq) {
    const size_t per = (((nq + m->n - 1) / m->n) + 7) & ~(size_t)7;   // 32-byte aligned ranges
    m->per = per;
    for (int i = 0; i < m->n; ++i) {   // every gathered vector exists before any kernel stores into it
        SIB_CHECK(cudaSetDevice(m->dev[i]));
        if (m->qs[i].ensure(per * 4) || m->qe[i].ensure(per * 4) || m->gather[i].ensure((size_t)m->n * per * 4)) return last_error_code();
    }
    const bool fused = m->p2p && m->n > 1;
    int rc = for_each_device(m, [&](int i) -> int {
        const size_t lo = (size_t)i * per < nq ? (size_t)i * per : nq;
        const size_t hi = lo + per < nq ? lo + per : nq;
        const size_t len = hi - lo;
        cudaStream_t s = m->st[i];
        uint32_t* mine = m->gather[i].as<uint32_t>() + (size_t)i * per;
        uint32_t* peers[16];
        int np = 0;
        if (fused)
            for (int k = 0; k < m->n; ++k)
                if (k != i) peers[np++] = m->gather[k].as<uint32_t>() + (size_t)i * per;
        SIB_CHECK(cudaEventRecord(m->ev[5 * i + 0], s));
        if (len) {
            SIB_CHECK(cudaMemcpyAsync(m->qs[i].p, qs + lo, len * 4, cudaMemcpyHostToDevice, s));
            SIB_CHECK(cudaMemcpyAsync(m->qe[i].p, qe + lo, len * 4, cudaMemcpyHostToDevice, s));
        }
        SIB_CHECK(cudaEventRecord(m->ev[5 * i + 1], s));
        if (len < per) {                                                             // padding counts as zero hits
            SIB_CHECK(cudaMemsetAsync(mine + len, 0, (per - len) * 4, s));
            for (int k = 0; k < np; ++k) SIB_CHECK(cudaMemsetAsync(peers[k] + len, 0, (per - len) * 4, s));
        }
        if (len) {
            int r = siCountFanoutDevice(m->ix[i], m->qs[i].as<int32_t>(), m->qe[i].as<int32_t>(), len, mine, peers, np, SI_ORDER_AUTO, (void*)s);
            if (r) return r;
        }
        SIB_CHECK(cudaEventRecord(m->ev[5 * i + 2], s));
        if (fused) SIB_CHECK(cudaEventRecord(m->ev_done[i], s));
        return 0;
    });
    if (rc) return rc;
    if (fused) {
        for (int i = 0; i < m->n; ++i) {   // device i may read the others' slots once their kernels are done
            SIB_CHECK(cudaSetDevice(m->dev[i]));
            for (int k = 0; k < m->n; ++k)
                if (k != i) SIB_CHECK(cudaStreamWaitEvent(m->st[i], m->ev_done[k], 0));
        }
        m->stats.peer_bytes += (unsigned long long)per * 4 * (m->n - 1) * m->n;
    } else if (m->n > 1) {
        SIB_NCCL(m, m->nccl.GroupStart());
        for (int i = 0; i < m->n; ++i) {
            uint32_t* all = m->gather[i].as<uint32_t>();
            SIB_NCCL(m, m->nccl.AllGather(all + (size_t)i * per, all, per, NCCL_UINT32, m->comm[i], m->st[i]));   // in place
        }
        SIB_NCCL(m, m->nccl.GroupEnd());
        m->stats.nccl_bytes += (unsigned long long)per * 4 * (m->n - 1) * m->n;   // received by all ranks
    }
    for (int i = 0; i < m->n; ++i) {
        SIB_CHECK(cudaSetDevice(m->dev[i]));
        SIB_CHECK(cudaEventRecord(m->ev[5 * i + 3], m->st[i]));
    }
    return 0;
}

}  // namespace

extern "C" {

siMulti* siMultiCreate(const int* devices, int n_devices) {
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible <= 0) { set_error(e, "cudaGetDeviceCount (no CUDA device?)", __FILE__, __LINE__); return nullptr; }
    if (n_devices <= 0) n_devices = visible;
    siMulti* m = new siMulti();
    m->n = n_devices;
    int prev = 0;
    cudaGetDevice(&prev);
    for (int i = 0; i < n_devices; ++i) {
        const int d = devices ? devices[i] : i;
        if (d < 0 || d >= visible) { set_error_msg(cudaErrorInvalidDevice, "siMultiCreate: device ordinal out of range"); siMultiDestroy(m); return nullptr; }
        m->dev.push_back(d);
        if (cudaSetDevice(d) != cudaSuccess) { set_error(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__); siMultiDestroy(m); return nullptr; }
        siIndex* ix = siIndexCreate();
        cudaStream_t s = nullptr;
        if (!ix || cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
            if (ix) siIndexDestroy(ix);
            set_error_msg(cudaErrorUnknown, "siMultiCreate: could not create the per-device index or stream");
            siMultiDestroy(m);
            return nullptr;
        }
        m->ix.push_back(ix);
        m->st.push_back(s);
        for (int k = 0; k < 5; ++k) {
            cudaEvent_t ev = nullptr;
            cudaEventCreate(&ev);
            m->ev.push_back(ev);
        }
    }
    m->in_s.resize(n_devices); m->in_e.resize(n_devices); m->in_v.resize(n_devices);
    m->qs.resize(n_devices); m->qe.resize(n_devices); m->gather.resize(n_devices);
    m->offsets.resize(n_devices); m->out.resize(n_devices);
    if (n_devices > 1) {
        if (!m->nccl.load()) {
            set_error_msg(cudaErrorUnknown, "siMultiCreate: libnccl.so.2 not found (needed for more than one device)");
            siMultiDestroy(m);
            return nullptr;
        }
        m->comm.assign(n_devices, nullptr);
        int r = m->nccl.CommInitAll(m->comm.data(), n_devices, m->dev.data());
        if (r != 0) {
            char msg[200];
            snprintf(msg, sizeof(msg), "ncclCommInitAll failed: %d (%s)", r, m->nccl.GetErrorString ? m->nccl.GetErrorString(r) : "?");
            set_error_msg(cudaErrorUnknown, msg);
            m->comm.clear();
            siMultiDestroy(m);
            return nullptr;
        }
        if (m->nccl.GetVersion) m->nccl.GetVersion(&m->stats.nccl_version);
        // peer access between every pair (NVLink / NVSwitch): the count kernels then store straight into the other devices'
        // gathered vectors. SIB_MULTI_P2P=0 keeps the NCCL all-gather (also taken when a pair cannot reach each other).
        const char* env = getenv("SIB_MULTI_P2P");
        bool p2p = n_devices <= 16 && !(env && atoi(env) == 0);
        for (int i = 0; i < n_devices && p2p; ++i)
            for (int k = 0; k < n_devices && p2p; ++k) {
                int can = 0;
                if (i != k && (cudaDeviceCanAccessPeer(&can, m->dev[i], m->dev[k]) != cudaSuccess || !can)) p2p = false;
            }
        for (int i = 0; i < n_devices && p2p; ++i) {
            cudaSetDevice(m->dev[i]);
            for (int k = 0; k < n_devices; ++k) {
                if (i == k) continue;
                cudaError_t pe = cudaDeviceEnablePeerAccess(m->dev[k], 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) p2p = false;
                (void)cudaGetLastError();
            }
        }
        m->p2p = p2p;
        m->stats.peer_access = p2p ? 1 : 0;
        for (int i = 0; i < n_devices; ++i) {
            cudaSetDevice(m->dev[i]);
            cudaEvent_t ev = nullptr;
            cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            m->ev_done.push_back(ev);
        }
    }
    cudaSetDevice(prev);
    return m;
}

void siMultiDestroy(siMulti* m) {
    if (!m) return;
    int prev = 0;
    cudaGetDevice(&prev);
    for (size_t i = 0; i < m->comm.size(); ++i)
        if (m->comm[i]) m->nccl.CommDestroy(m->comm[i]);
    for (size_t i = 0; i < m->ix.size(); ++i) {
        cudaSetDevice(m->dev[i]);
        for (std::vector<DevBuf>* v : {&m->in_s, &m->in_e, &m->in_v, &m->qs, &m->qe, &m->gather, &m->offsets, &m->out})
            if (i < v->size()) (*v)[i].release();
        for (int k = 0; k < 5; ++k)
            if (5 * i + k < m->ev.size() && m->ev[5 * i + k]) cudaEventDestroy(m->ev[5 * i + k]);
        if (i < m->ev_done.size() && m->ev_done[i]) cudaEventDestroy(m->ev_done[i]);
        if (m->st[i]) cudaStreamDestroy(m->st[i]);
        siIndexDestroy(m->ix[i]);
    }
    cudaSetDevice(prev);
    delete m;
}

int siMultiDeviceCount(const siMulti* m) { return m ? m->n : 0; }
siIndex* siMultiIndexOf(siMulti* m, int rank) { return (m && rank >= 0 && rank < m->n) ? m->ix[rank] : nullptr; }

int siMultiBuildReplicated(siMulti* m, const int32_t* starts, const int32_t* ends, const int32_t* values, size_t n) {
    if (!m) return cudaErrorInvalidValue;
    std::lock_guard<std::mutex> lk(m->mu);
    int prev = 0;
    cudaGetDevice(&prev);
    m->built = false;
    m->n_intervals = n;
    int rc = 0;
    if (n) {
        // one upload, then NVLink: broadcast of the three input columns from the first device
        for (int i = 0; i < m->n && !rc; ++i) {
            cudaSetDevice(m->dev[i]);
            if (m->in_s[i].ensure(n * 4) || m->in_e[i].ensure(n * 4) || (values && m->in_v[i].ensure(n * 4))) rc = last_error_code();
        }
        if (!rc) {
            cudaSetDevice(m->dev[0]);
            cudaStream_t s0 = m->st[0];
            if (cudaMemcpyAsync(m->in_s[0].p, starts, n * 4, cudaMemcpyHostToDevice, s0) != cudaSuccess ||
                cudaMemcpyAsync(m->in_e[0].p, ends, n * 4, cudaMemcpyHostToDevice, s0) != cudaSuccess ||
                (values && cudaMemcpyAsync(m->in_v[0].p, values, n * 4, cudaMemcpyHostToDevice, s0) != cudaSuccess)) {
                set_error(cudaGetLastError(), "siMultiBuildReplicated upload", __FILE__, __LINE__);
                rc = last_error_code();
            }
        }
        if (!rc && m->n > 1) {
            auto bcast = [&]() -> int {
                SIB_NCCL(m, m->nccl.GroupStart());
                for (int i = 0; i < m->n; ++i) {
                    SIB_NCCL(m, m->nccl.Broadcast(m->in_s[0].p, m->in_s[i].p, n, NCCL_INT32, 0, m->comm[i], m->st[i]));
                    SIB_NCCL(m, m->nccl.Broadcast(m->in_e[0].p, m->in_e[i].p, n, NCCL_INT32, 0, m->comm[i], m->st[i]));
                    if (values) SIB_NCCL(m, m->nccl.Broadcast(m->in_v[0].p, m->in_v[i].p, n, NCCL_INT32, 0, m->comm[i], m->st[i]));
                }
                SIB_NCCL(m, m->nccl.GroupEnd());
                return 0;
            };
            rc = bcast();
            m->stats.nccl_bytes += (unsigned long long)n * 4 * (values ? 3 : 2) * (m->n - 1);
        }
    }
    if (!rc)
        rc = for_each_device(m, [&](int i) -> int {
            return siIndexBuildDevice(m->ix[i], n ? m->in_s[i].as<int32_t>() : nullptr, n ? m->in_e[i].as<int32_t>() : nullptr,
                                      (n && values) ? m->in_v[i].as<int32_t>() : nullptr, n, (void*)m->st[i]);
        });
    for (int i = 0; i < m->n; ++i) {   // the inputs were only needed for the build
        cudaSetDevice(m->dev[i]);
        m->in_s[i].release(); m->in_e[i].release(); m->in_v[i].release();
    }
    cudaSetDevice(prev);
    m->built = rc == 0;
    return rc;
}

int siMultiCountBatch(siMulti* m, const int32_t* qs, const int32_t* qe, size_t nq, uint32_t* counts_out) {
    if (!m || !m->built) { set_error_msg(cudaErrorNotReady, "siMultiCountBatch: build first"); return cudaErrorNotReady; }
    if (nq == 0) return 0;
    std::lock_guard<std::mutex> lk(m->mu);
    int prev = 0;
    cudaGetDevice(&prev);
    const auto t0 = std::chrono::steady_clock::now();
    int rc = count_ranges(m, qs, qe, nq);
    if (!rc) {
        const size_t per = m->per;
        for (int i = 0; i < m->n && !rc; ++i) {   // every device returns its own range on its own PCIe link
            const size_t lo = (size_t)i * per < nq ? (size_t)i * per : nq;
            const size_t hi = lo + per < nq ? lo + per : nq;
            cudaSetDevice(m->dev[i]);
            if (hi > lo && cudaMemcpyAsync(counts_out + lo, m->gather[i].as<uint32_t>() + lo, (hi - lo) * 4, cudaMemcpyDeviceToHost, m->st[i]) != cudaSuccess) {
                set_error(cudaGetLastError(), "siMultiCountBatch copy-back", __FILE__, __LINE__);
                rc = last_error_code();
            }
            cudaEventRecord(m->ev[5 * i + 4], m->st[i]);
        }
        if (!rc) rc = sync_all(m);
    }
    if (!rc) {
        m->stats.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        m->stats.ms_h2d = span_ms(m, 0, 1);
        m->stats.ms_count = span_ms(m, 1, 2);
        m->stats.ms_gather = span_ms(m, 2, 3);
        m->stats.ms_d2h = span_ms(m, 3, 4);
    }
    cudaSetDevice(prev);
    return rc;
}

// the gathered counts of the last batch as they sit on one device: n_devices ranges of `per` entries,
// range r holding queries [r * per, min(nq, (r + 1) * per)) and zeros after them
int siMultiDeviceCounts(siMulti* m, int rank, const uint32_t** d_counts, size_t* per) {
    if (!m || rank < 0 || rank >= m->n || !d_counts || !per) return cudaErrorInvalidValue;
    *d_counts = m->gather[rank].as<uint32_t>();
    *per = m->per;
    return 0;
}

int siMultiSearchValuesBatch(siMulti* m, const int32_t* qs, const int32_t* qe, size_t nq, size_t* offsets_out, cIndexResult* found) {
    if (!m || !m->built) { set_error_msg(cudaErrorNotReady, "siMultiSearchValuesBatch: build first"); return cudaErrorNotReady; }
    if (nq == 0) { if (offsets_out) offsets_out[0] = 0; return 0; }
    std::lock_guard<std::mutex> lk(m->mu);
    int prev = 0;
    cudaGetDevice(&prev);
    const auto t0 = std::chrono::steady_clock::now();
    int rc = count_ranges(m, qs, qe, nq);
    const size_t per = m->per, all = per * (size_t)m->n;
    std::vector<uint64_t> base(m->n + 1, 0);
    if (!rc)
        rc = for_each_device(m, [&](int i) -> int {
            // global CSR offsets on every device: exclusive scan over the gathered counts (padding counts are zero)
            if (m->offsets[i].ensure((all + 1) * 8 + 64)) return last_error_code();
            int r = siScanDevice(m->ix[i], m->gather[i].as<uint32_t>(), all, m->offsets[i].as<uint64_t>(), (void*)m->st[i]);
            if (r) return r;
            SIB_CHECK(cudaMemcpyAsync(&base[i], m->offsets[i].as<uint64_t>() + (size_t)i * per, 8, cudaMemcpyDeviceToHost, m->st[i]));
            if (i == m->n - 1) SIB_CHECK(cudaMemcpyAsync(&base[m->n], m->offsets[i].as<uint64_t>() + all, 8, cudaMemcpyDeviceToHost, m->st[i]));
            SIB_CHECK(cudaStreamSynchronize(m->st[i]));
            return 0;
        });
    const uint64_t total = base[m->n];
    if (!rc && total && !grow(found, found->size + total, sizeof(int32_t))) rc = cudaErrorMemoryAllocation;
    if (!rc)
        rc = for_each_device(m, [&](int i) -> int {
            const size_t lo = (size_t)i * per < nq ? (size_t)i * per : nq;
            const size_t hi = lo + per < nq ? lo + per : nq;
            cudaStream_t s = m->st[i];
            const uint64_t seg = base[i + 1] - base[i];
            if (hi > lo && seg) {
                if (m->out[i].ensure(seg * 4)) return last_error_code();
                // the fill writes at absolute CSR offsets: hand it this device's segment shifted back by its base
                int32_t* shifted = m->out[i].as<int32_t>() - base[i];
                int r = siFillDevice(m->ix[i], m->qs[i].as<int32_t>(), m->qe[i].as<int32_t>(), hi - lo,
                                     m->offsets[i].as<uint64_t>() + lo, SI_FILL_VALUES, shifted, SI_ORDER_AUTO, (void*)s);
                if (r) return r;
                SIB_CHECK(cudaMemcpyAsync(found->data + found->size + base[i], m->out[i].p, seg * 4, cudaMemcpyDeviceToHost, s));
            }
            if (offsets_out && hi > lo)
                SIB_CHECK(cudaMemcpyAsync(offsets_out + lo, m->offsets[i].as<uint64_t>() + lo, (hi - lo) * 8, cudaMemcpyDeviceToHost, s));
            SIB_CHECK(cudaEventRecord(m->ev[5 * i + 4], s));
            SIB_CHECK(cudaStreamSynchronize(s));
            return 0;
        });
    if (!rc) {
        if (offsets_out) offsets_out[nq] = total;
        found->size += total;
        m->stats.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        m->stats.ms_h2d = span_ms(m, 0, 1);
        m->stats.ms_count = span_ms(m, 1, 2);
        m->stats.ms_gather = span_ms(m, 2, 3);
        m->stats.ms_d2h = span_ms(m, 3, 4);
    }
    cudaSetDevice(prev);
    return rc;
}

int siMultiLastStats(const siMulti* m, siMultiStats* out) {
    if (!m || !out) return cudaErrorInvalidValue;
    *out = m->stats;
    return 0;
}

}  // extern "C"
