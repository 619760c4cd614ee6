// setops.cu -- the set-algebra half of the reference's C ABI (c_superintervals.h:243-324
// declarations, :823-1064 definitions) on the device: the callers either side of the batch
// query path. Each function returns a NEW, un-indexed handle whose arrays are in the
// reference's emission order. Kernels: setops_kernels.cuh. The query-shaped operations
// (intersection / difference / symmetricDifference) run A's stored intervals as one batch
// through B's index: count -> scan -> fill, exactly the hot path.
#include "../../include/superintervals_b200.h"

#include "common.cuh"
#include "host_common.cuh"
#include "index.cuh"
#include "setops_kernels.cuh"

#include <cstdlib>
#include <cstring>

using namespace sib;

extern "C" int si_b200_resolve_order_(siIndex* ix, const int32_t* d_qs, size_t n, void* stream);

namespace {

#define SO_LAUNCH(kernel, n, stream, ...)                                                       \
    do {                                                                                        \
        kernel<<<(unsigned)(((uint64_t)(n) + SO_THREADS - 1) / SO_THREADS), SO_THREADS, 0, (stream)>>>(__VA_ARGS__); \
        SIB_CHECK_LAUNCH();                                                                     \
        note_launch();                                                                          \
    } while (0)

// a scratch index: stream, scan workspace and staging for operations that query nothing
struct TempIndex {
    siIndex* ix;
    TempIndex() : ix(siIndexCreate()) {}
    ~TempIndex() { if (ix) siIndexDestroy(ix); }
    TempIndex(const TempIndex&) = delete;
    TempIndex& operator=(const TempIndex&) = delete;
};

// device buffers released at scope exit
struct Bufs {
    DevBuf b[12];
    ~Bufs() { for (auto& x : b) x.release(); }
};

// an interval list on the device (not owning)
struct DevSet {
    const int32_t* s = nullptr;
    const int32_t* e = nullptr;
    const int32_t* d = nullptr;
    size_t n = 0;
};

int upload(siIndex* ix, const int32_t* s, const int32_t* e, const int32_t* d, size_t n, DevBuf& bs, DevBuf& be,
           DevBuf& bd, DevSet* out) {
    if (bs.ensure(n * 4 + 4) || be.ensure(n * 4 + 4) || bd.ensure(n * 4 + 4)) return last_error_code();
    int rc = copy_h2d(ix, bs.p, s, n * 4, ix->own_stream);
    if (!rc) rc = copy_h2d(ix, be.p, e, n * 4, ix->own_stream);
    if (!rc) rc = copy_h2d(ix, bd.p, d, n * 4, ix->own_stream);
    *out = DevSet{bs.as<int32_t>(), be.as<int32_t>(), bd.as<int32_t>(), n};
    return rc;
}

// A new handle holding n intervals copied from the device, flags as addInterval x n would
// have left them (c.h:402-419: endSorted is only tracked while startSorted holds).
cSuperIntervals* adopt(siIndex* ix, const int32_t* ds, const int32_t* de, const int32_t* dd, size_t n) {
    cSuperIntervals* out = createSuperIntervals();
    if (!out || n == 0) return out;
    reserveSuperIntervals(out, n);
    if (!out->starts || !out->ends || !out->data) { set_error_msg(cudaErrorMemoryAllocation, "set operation: out of host memory"); return out; }
    if (copy_d2h(ix, out->starts, ds, n * 4, ix->own_stream) || copy_d2h(ix, out->ends, de, n * 4, ix->own_stream) ||
        copy_d2h(ix, out->data, dd, n * 4, ix->own_stream))
        return out;
    out->size = n;
    for (size_t i = 1; i < n && out->startSorted; ++i) {
        if (out->starts[i] < out->starts[i - 1]) out->startSorted = false;
        else if (out->starts[i] == out->starts[i - 1] && out->ends[i] > out->ends[i - 1]) out->endSorted = false;
    }
    return out;
}

int read_u64(const uint64_t* dev, uint64_t* host, cudaStream_t s) {
    SIB_CHECK(cudaMemcpyAsync(host, dev, 8, cudaMemcpyDeviceToHost, s));
    SIB_CHECK(cudaStreamSynchronize(s));
    return 0;
}

// counts (uint32, n entries) -> offsets (uint64, n + 1 entries) on the device, total read back
int scan_counts(siIndex* ix, DevBuf& cnt, size_t n, DevBuf& off, uint64_t* total) {
    if (off.ensure((n + 1) * 8 + 64)) return last_error_code();
    int rc = siScanDevice(ix, cnt.as<uint32_t>(), n, off.as<uint64_t>(), ix->own_stream);
    if (rc) return rc;
    return read_u64(off.as<uint64_t>() + n, total, ix->own_stream);
}

// ---- sorted sweep: merge_overlaps / unique on the device ------------------------------------
// Sorts (s, e, d) the way build() does (stable, start asc / end desc; untouched when add()'s
// flags already say so, c.h:831-852), marks cluster heads, scans, scatters. Results stay on the
// device in B.b[4..6]; with `fold_heads` the sorted payloads and the head flags are also
// brought to the host so that a caller-supplied combine() can be folded over each cluster.
struct Clusters {
    DevSet set;                 // merged / unique intervals on the device
    bool wellformed = false;
    int32_t* sorted_data = nullptr;   // host copies for the combine fold (malloc'd), or nullptr
    uint32_t* heads = nullptr;
    ~Clusters() { free(sorted_data); free(heads); }
};

int cluster_device(siIndex* ix, const int32_t* s, const int32_t* e, const int32_t* d, size_t n, bool unique,
                   bool fold_heads, Bufs& B, Clusters* out) {
    int rc = siIndexBuildHost(ix, s, e, d, n);   // the sort (and, for merge, the prefix maxima of the ends)
    if (rc) return rc;
    siDeviceView v;
    if (siIndexDeviceView(ix, &v)) return last_error_code();
    cudaStream_t st = ix->own_stream;
    const uint32_t n32 = (uint32_t)n;
    DevBuf &head = B.b[0], &runmax = B.b[1], &off = B.b[2];
    if (head.ensure(n * 4 + 64) || (!unique && runmax.ensure(n * 4))) return last_error_code();
    out->wellformed = ix->wellformed;
    if (unique) {
        SO_LAUNCH(so_unique_heads_kernel, n, st, v.starts, v.ends, n32, head.as<uint32_t>());
    } else if (ix->wellformed) {
        SO_LAUNCH(so_merge_heads_kernel, n, st, v.starts, v.ends, ix->pmax32, n32, head.as<uint32_t>(), runmax.as<int32_t>());
    } else {
        so_merge_heads_seq_kernel<<<1, 32, 0, st>>>(v.starts, v.ends, n32, head.as<uint32_t>(), runmax.as<int32_t>());
        SIB_CHECK_LAUNCH();
        note_launch();
    }
    uint64_t m = 0;
    rc = scan_counts(ix, head, n, off, &m);
    if (rc) return rc;
    DevBuf &os = B.b[4], &oe = B.b[5], &od = B.b[6];
    if (os.ensure(m * 4 + 4) || oe.ensure(m * 4 + 4) || od.ensure(m * 4 + 4)) return last_error_code();
    SO_LAUNCH(so_cluster_scatter_kernel, n, st, v.starts, v.ends, v.values, head.as<uint32_t>(), off.as<uint64_t>(),
              unique ? (const int32_t*)nullptr : runmax.as<int32_t>(), n32, os.as<int32_t>(), oe.as<int32_t>(), od.as<int32_t>());
    out->set = DevSet{os.as<int32_t>(), oe.as<int32_t>(), od.as<int32_t>(), (size_t)m};
    if (fold_heads) {
        out->sorted_data = (int32_t*)malloc(n * 4);
        out->heads = (uint32_t*)malloc(n * 4);
        if (!out->sorted_data || !out->heads) { set_error_msg(cudaErrorMemoryAllocation, "set operation: out of host memory"); return cudaErrorMemoryAllocation; }
        if (copy_d2h(ix, out->sorted_data, v.values, n * 4, st) || copy_d2h(ix, out->heads, head.p, n * 4, st)) return last_error_code();
    }
    return 0;
}

cSuperIntervals* cluster_op(const int32_t* s, const int32_t* e, const int32_t* d, size_t n, cCombineFn combine, bool unique) {
    if (n == 0) return createSuperIntervals();
    if (n > 0xFFFFFFF0ull) { set_error_msg(cudaErrorInvalidValue, "set operation: too many intervals"); return createSuperIntervals(); }
    TempIndex t;
    if (!t.ix) return createSuperIntervals();
    Bufs B;
    Clusters c;
    if (cluster_device(t.ix, s, e, d, n, unique, combine != nullptr, B, &c)) return createSuperIntervals();
    cSuperIntervals* out = adopt(t.ix, c.set.s, c.set.e, c.set.d, c.set.n);
    if (combine && out->size == c.set.n) {
        // the callback is host code: fold it over each cluster, left to right (c.h:868,1055)
        size_t cl = (size_t)-1;
        int32_t acc = 0;
        for (size_t i = 0; i < n; ++i) {
            if (c.heads[i]) { if (cl != (size_t)-1) out->data[cl] = acc; ++cl; acc = c.sorted_data[i]; }
            else acc = combine(acc, c.sorted_data[i]);
        }
        out->data[cl] = acc;
    }
    return out;
}

// ---- CSR consumers ---------------------------------------------------------------------------
// A's stored intervals as one query batch against B's index; leaves the batch in B's staging
// buffers (h_qs, h_qe), the CSR offsets in h_offsets and the hit payloads in h_out.
int query_stage(Handle* hb, const cSuperIntervals* a, int what, size_t elem, uint64_t* total) {
    siIndex* ix = hb->ix;
    const size_t n = a->size;
    int rc = stage_queries(ix, a->starts, a->ends, n);
    if (rc) return rc;
    if (ix->h_counts.ensure(n * 4 + 64) || ix->h_offsets.ensure((n + 1) * 8 + 64)) return last_error_code();
    const int32_t* dqs = ix->h_qs.as<int32_t>();
    const int32_t* dqe = ix->h_qe.as<int32_t>();
    // stored intervals of an indexed A arrive position-sorted; one device check decides
    const int order = si_b200_resolve_order_(ix, dqs, n, ix->own_stream);
    if (order < 0) return last_error_code();
    rc = siCountDevice(ix, dqs, dqe, n, ix->h_counts.as<uint32_t>(), order, ix->own_stream);
    if (rc) return rc;
    rc = siScanDevice(ix, ix->h_counts.as<uint32_t>(), n, ix->h_offsets.as<uint64_t>(), ix->own_stream);
    if (rc) return rc;
    rc = read_u64(ix->h_offsets.as<uint64_t>() + n, total, ix->own_stream);
    if (rc) return rc;
    if (ix->h_out.ensure(*total * elem + 16)) return last_error_code();
    if (*total) rc = siFillDevice(ix, dqs, dqe, n, ix->h_offsets.as<uint64_t>(), what, ix->h_out.p, order, ix->own_stream);
    return rc;
}

bool both_usable(const cSuperIntervals* si, cSuperIntervals* other, const char* who) {
    if (si->size > 0xFFFFFFF0ull) { set_error_msg(cudaErrorInvalidValue, "set operation: too many intervals"); return false; }
    if (other->size == 0) return true;              // nothing to query: handled by the callers
    return handle_ready(H(other), who);
}

}  // namespace

extern "C" {

// ---- c.h:854-881 -----------------------------------------------------------------------------
cSuperIntervals* mergeOverlaps(const cSuperIntervals* si, cCombineFn combine) {
    return cluster_op(si->starts, si->ends, si->data, si->size, combine, false);
}

// ---- c.h:1042-1064 ---------------------------------------------------------------------------
cSuperIntervals* uniqueIntervals(const cSuperIntervals* si, cCombineFn combine) {
    return cluster_op(si->starts, si->ends, si->data, si->size, combine, true);
}

// ---- c.h:907-919: concatenate (this first), then merge ----------------------------------------
cSuperIntervals* unionWith(const cSuperIntervals* si, const cSuperIntervals* other, cCombineFn combine) {
    const size_t n1 = si->size, n2 = other->size, n = n1 + n2;
    if (n == 0) return createSuperIntervals();
    int32_t* s = (int32_t*)malloc(n * 4);
    int32_t* e = (int32_t*)malloc(n * 4);
    int32_t* d = (int32_t*)malloc(n * 4);
    cSuperIntervals* out = nullptr;
    if (s && e && d) {
        if (n1) { memcpy(s, si->starts, n1 * 4); memcpy(e, si->ends, n1 * 4); memcpy(d, si->data, n1 * 4); }
        if (n2) { memcpy(s + n1, other->starts, n2 * 4); memcpy(e + n1, other->ends, n2 * 4); memcpy(d + n1, other->data, n2 * 4); }
        out = cluster_op(s, e, d, n, combine, false);
    } else {
        set_error_msg(cudaErrorMemoryAllocation, "unionWith: out of host memory");
        out = createSuperIntervals();
    }
    free(s); free(e); free(d);
    return out;
}

// ---- c.h:883-905 -----------------------------------------------------------------------------
cSuperIntervals* intervalGaps(const cSuperIntervals* si, int32_t lo, int32_t hi, int32_t fill) {
    TempIndex t;
    if (!t.ix) return createSuperIntervals();
    siIndex* ix = t.ix;
    cudaStream_t st = ix->own_stream;
    Bufs B;
    Clusters c;
    if (si->size) {
        if (si->size > 0xFFFFFFF0ull) { set_error_msg(cudaErrorInvalidValue, "intervalGaps: too many intervals"); return createSuperIntervals(); }
        if (cluster_device(ix, si->starts, si->ends, si->data, si->size, false, false, B, &c)) return createSuperIntervals();
    }
    const size_t m = c.set.n;
    DevBuf &os = B.b[7], &oe = B.b[8], &od = B.b[9], &cnt = B.b[10], &off = B.b[11];
    if (os.ensure((m + 1) * 4) || oe.ensure((m + 1) * 4) || od.ensure((m + 1) * 4) || ix->small.ensure(256)) return createSuperIntervals();
    uint64_t pieces = 0;
    if (m && !c.wellformed) {
        unsigned long long* d_count = reinterpret_cast<unsigned long long*>(ix->small.as<uint32_t>() + 32);
        so_gaps_seq_kernel<<<1, 32, 0, st>>>(c.set.s, c.set.e, (uint32_t)m, lo, hi, fill, os.as<int32_t>(), oe.as<int32_t>(),
                                            od.as<int32_t>(), d_count);
        const cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) { set_error(le, "so_gaps_seq_kernel", __FILE__, __LINE__); return createSuperIntervals(); }
        note_launch();
        if (read_u64(reinterpret_cast<const uint64_t*>(d_count), &pieces, st)) return createSuperIntervals();
    } else {
        uint32_t any_in = 0;
        if (m) {
            uint32_t* d_any = ix->small.as<uint32_t>() + 34;
            if (cnt.ensure(m * 4 + 64) || cudaMemsetAsync(d_any, 0, 4, st) != cudaSuccess) return createSuperIntervals();
            auto run = [&]() -> int {
                SO_LAUNCH((so_gaps_kernel<false>), m, st, c.set.s, c.set.e, (uint32_t)m, lo, hi, fill, cnt.as<uint32_t>(),
                          (const uint64_t*)nullptr, d_any, (int32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr);
                int rc = scan_counts(ix, cnt, m, off, &pieces);
                if (rc) return rc;
                SIB_CHECK(cudaMemcpyAsync(&any_in, d_any, 4, cudaMemcpyDeviceToHost, st));
                SIB_CHECK(cudaStreamSynchronize(st));
                if (pieces)
                    SO_LAUNCH((so_gaps_kernel<true>), m, st, c.set.s, c.set.e, (uint32_t)m, lo, hi, fill, (uint32_t*)nullptr,
                              off.as<uint64_t>(), (uint32_t*)nullptr, os.as<int32_t>(), oe.as<int32_t>(), od.as<int32_t>());
                return 0;
            };
            if (run()) return createSuperIntervals();
        }
        if (!any_in) {
            // nothing stored inside the span: the whole span is one gap (cursor never moved)
            cSuperIntervals* out = createSuperIntervals();
            if (lo <= hi) addInterval(out, lo, hi, fill);
            return out;
        }
    }
    return adopt(ix, os.as<int32_t>(), oe.as<int32_t>(), od.as<int32_t>(), (size_t)pieces);
}

// ---- c.h:921-939 -----------------------------------------------------------------------------
static cSuperIntervals* intersection_impl(const cSuperIntervals* si, cSuperIntervals* other, cCombineFn combine,
                                          cIndexResult* other_data) {
    if (si->size == 0 || other->size == 0 || !both_usable(si, other, "intersection")) return createSuperIntervals();
    Handle* hb = H(other);
    siIndex* ix = hb->ix;
    cudaStream_t st = ix->own_stream;
    const size_t n = si->size;
    uint64_t hits = 0, pieces = 0;
    if (query_stage(hb, si, SI_FILL_ITEMS, sizeof(Interval), &hits)) return createSuperIntervals();
    Bufs B;
    DevBuf &qd = B.b[0], &cnt = B.b[1], &off = B.b[2], &os = B.b[3], &oe = B.b[4], &oa = B.b[5], &ob = B.b[6];
    if (qd.ensure(n * 4) || cnt.ensure(n * 4 + 64) || copy_h2d(ix, qd.p, si->data, n * 4, st)) return createSuperIntervals();
    const Item3i* items = ix->h_out.as<Item3i>();
    auto run = [&]() -> int {
        SO_LAUNCH((so_intersection_kernel<false>), n, st, ix->h_qs.as<int32_t>(), ix->h_qe.as<int32_t>(), qd.as<int32_t>(), (uint32_t)n,
                  ix->h_offsets.as<uint64_t>(), items, cnt.as<uint32_t>(), (const uint64_t*)nullptr, (int32_t*)nullptr,
                  (int32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr);
        int rc = scan_counts(ix, cnt, n, off, &pieces);
        if (rc) return rc;
        if (os.ensure(pieces * 4 + 4) || oe.ensure(pieces * 4 + 4) || oa.ensure(pieces * 4 + 4) || ob.ensure(pieces * 4 + 4)) return last_error_code();
        if (pieces)
            SO_LAUNCH((so_intersection_kernel<true>), n, st, ix->h_qs.as<int32_t>(), ix->h_qe.as<int32_t>(), qd.as<int32_t>(), (uint32_t)n,
                      ix->h_offsets.as<uint64_t>(), items, (uint32_t*)nullptr, off.as<uint64_t>(), os.as<int32_t>(),
                      oe.as<int32_t>(), oa.as<int32_t>(), ob.as<int32_t>());
        return 0;
    };
    if (run()) return createSuperIntervals();
    cSuperIntervals* out = adopt(ix, os.as<int32_t>(), oe.as<int32_t>(), oa.as<int32_t>(), (size_t)pieces);
    if (other_data && out->size == pieces && pieces) {
        // the other side's data of every piece, appended for the caller (intersectionPairs)
        if (grow(other_data, other_data->size + pieces, sizeof(int32_t)) &&
            !copy_d2h(ix, other_data->data + other_data->size, ob.p, pieces * 4, st))
            other_data->size += pieces;
    }
    if (combine && out->size == pieces && pieces) {
        int32_t* b = (int32_t*)malloc(pieces * 4);
        if (b && !copy_d2h(ix, b, ob.p, pieces * 4, st))
            for (size_t i = 0; i < pieces; ++i) out->data[i] = combine(out->data[i], b[i]);
        free(b);
    }
    return out;
}

cSuperIntervals* intersection(const cSuperIntervals* si, cSuperIntervals* other, cCombineFn combine) {
    return intersection_impl(si, other, combine, nullptr);
}
cSuperIntervals* intersectionPairs(const cSuperIntervals* si, cSuperIntervals* other, cIndexResult* other_data) {
    return intersection_impl(si, other, nullptr, other_data);
}

// ---- c.h:950-974 -----------------------------------------------------------------------------
cSuperIntervals* difference(const cSuperIntervals* si, cSuperIntervals* other) {
    if (si->size == 0) return createSuperIntervals();
    if (other->size == 0) {   // nothing covers anything: every stored interval survives whole (cursor <= end)
        cSuperIntervals* out = createSuperIntervals();
        for (size_t k = 0; k < si->size; ++k)
            if (si->starts[k] <= si->ends[k]) addInterval(out, si->starts[k], si->ends[k], si->data[k]);
        return out;
    }
    if (!both_usable(si, other, "difference")) return createSuperIntervals();
    Handle* hb = H(other);
    siIndex* ix = hb->ix;
    cudaStream_t st = ix->own_stream;
    const size_t n = si->size;
    uint64_t hits = 0, pieces = 0;
    if (query_stage(hb, si, SI_FILL_KEYS, sizeof(KeyPair), &hits)) return createSuperIntervals();
    Bufs B;
    DevBuf &qd = B.b[0], &cnt = B.b[1], &off = B.b[2], &os = B.b[3], &oe = B.b[4], &od = B.b[5];
    if (qd.ensure(n * 4) || cnt.ensure(n * 4 + 64) || copy_h2d(ix, qd.p, si->data, n * 4, st)) return createSuperIntervals();
    const int2* keys = ix->h_out.as<int2>();
    auto run = [&]() -> int {
        SO_LAUNCH((so_difference_kernel<false>), n, st, ix->h_qs.as<int32_t>(), ix->h_qe.as<int32_t>(), qd.as<int32_t>(), (uint32_t)n,
                  ix->h_offsets.as<uint64_t>(), keys, cnt.as<uint32_t>(), (const uint64_t*)nullptr, (int32_t*)nullptr,
                  (int32_t*)nullptr, (int32_t*)nullptr);
        int rc = scan_counts(ix, cnt, n, off, &pieces);
        if (rc) return rc;
        if (os.ensure(pieces * 4 + 4) || oe.ensure(pieces * 4 + 4) || od.ensure(pieces * 4 + 4)) return last_error_code();
        if (pieces)
            SO_LAUNCH((so_difference_kernel<true>), n, st, ix->h_qs.as<int32_t>(), ix->h_qe.as<int32_t>(), qd.as<int32_t>(), (uint32_t)n,
                      ix->h_offsets.as<uint64_t>(), keys, (uint32_t*)nullptr, off.as<uint64_t>(), os.as<int32_t>(),
                      oe.as<int32_t>(), od.as<int32_t>());
        return 0;
    };
    if (run()) return createSuperIntervals();
    return adopt(ix, os.as<int32_t>(), oe.as<int32_t>(), od.as<int32_t>(), (size_t)pieces);
}

// ---- c.h:976-983 -----------------------------------------------------------------------------
cSuperIntervals* symmetricDifference(cSuperIntervals* si, cSuperIntervals* other) {
    cSuperIntervals* a_minus_b = difference(si, other);
    cSuperIntervals* b_minus_a = difference(other, si);
    cSuperIntervals* out = unionWith(a_minus_b, b_minus_a, nullptr);
    destroySuperIntervals(a_minus_b);
    destroySuperIntervals(b_minus_a);
    return out;
}

// ---- c.h:985-998 ----------------------------------------------------------------------------
bool intervalSpan(const cSuperIntervals* si, int32_t* lo_out, int32_t* hi_out) {
    const size_t n = si->size;
    if (n == 0) return false;
    TempIndex t;
    if (!t.ix) return false;
    siIndex* ix = t.ix;
    cudaStream_t st = ix->own_stream;
    Bufs B;
    DevSet a;
    if (n > 0xFFFFFFF0ull || upload(ix, si->starts, si->ends, si->data, n, B.b[0], B.b[1], B.b[2], &a) || ix->small.ensure(256)) return false;
    int32_t* d_res = reinterpret_cast<int32_t*>(ix->small.as<uint32_t>() + 36);
    int32_t res[2] = {INT_MAX, INT_MIN};
    if (cudaMemcpyAsync(d_res, res, 8, cudaMemcpyHostToDevice, st) != cudaSuccess) return false;
    const unsigned grid = (unsigned)((n + SO_THREADS - 1) / SO_THREADS < (size_t)ix->sm_count * 8 ? (n + SO_THREADS - 1) / SO_THREADS : (size_t)ix->sm_count * 8);
    so_span_kernel<<<grid, SO_THREADS, 0, st>>>(a.s, a.e, (uint32_t)n, d_res);
    const cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) { set_error(le, "so_span_kernel", __FILE__, __LINE__); return false; }
    note_launch();
    if (cudaMemcpyAsync(res, d_res, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        set_error(cudaGetLastError(), "intervalSpan", __FILE__, __LINE__);
        return false;
    }
    *lo_out = res[0];
    *hi_out = res[1];
    return true;
}

// ---- c.h:1000-1040 ---------------------------------------------------------------------------
static cSuperIntervals* resize_op(const cSuperIntervals* si, int32_t left, int32_t right, int32_t lo, int32_t hi, bool flank) {
    const size_t n = si->size;
    if (n == 0) return createSuperIntervals();
    TempIndex t;
    if (!t.ix) return createSuperIntervals();
    siIndex* ix = t.ix;
    cudaStream_t st = ix->own_stream;
    Bufs B;
    DevSet a;
    DevBuf &cnt = B.b[3], &off = B.b[4], &os = B.b[5], &oe = B.b[6], &od = B.b[7];
    if (n > 0xFFFFFFF0ull || upload(ix, si->starts, si->ends, si->data, n, B.b[0], B.b[1], B.b[2], &a) || cnt.ensure(n * 4 + 64))
        return createSuperIntervals();
    uint64_t pieces = 0;
    auto run = [&]() -> int {
        if (flank)
            SO_LAUNCH((so_resize_kernel<true, false>), n, st, a.s, a.e, a.d, (uint32_t)n, left, right, lo, hi, cnt.as<uint32_t>(),
                      (const uint64_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr);
        else
            SO_LAUNCH((so_resize_kernel<false, false>), n, st, a.s, a.e, a.d, (uint32_t)n, left, right, lo, hi, cnt.as<uint32_t>(),
                      (const uint64_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr);
        int rc = scan_counts(ix, cnt, n, off, &pieces);
        if (rc) return rc;
        if (os.ensure(pieces * 4 + 4) || oe.ensure(pieces * 4 + 4) || od.ensure(pieces * 4 + 4)) return last_error_code();
        if (!pieces) return 0;
        if (flank)
            SO_LAUNCH((so_resize_kernel<true, true>), n, st, a.s, a.e, a.d, (uint32_t)n, left, right, lo, hi, (uint32_t*)nullptr,
                      off.as<uint64_t>(), os.as<int32_t>(), oe.as<int32_t>(), od.as<int32_t>());
        else
            SO_LAUNCH((so_resize_kernel<false, true>), n, st, a.s, a.e, a.d, (uint32_t)n, left, right, lo, hi, (uint32_t*)nullptr,
                      off.as<uint64_t>(), os.as<int32_t>(), oe.as<int32_t>(), od.as<int32_t>());
        return 0;
    };
    if (run()) return createSuperIntervals();
    return adopt(ix, os.as<int32_t>(), oe.as<int32_t>(), od.as<int32_t>(), (size_t)pieces);
}

cSuperIntervals* expandIntervals(const cSuperIntervals* si, int32_t left, int32_t right, int32_t lo, int32_t hi) {
    return resize_op(si, left, right, lo, hi, false);
}
cSuperIntervals* flankIntervals(const cSuperIntervals* si, int32_t left, int32_t right, int32_t lo, int32_t hi) {
    return resize_op(si, left, right, lo, hi, true);
}

}  // extern "C"
