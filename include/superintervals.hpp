// superintervals.hpp -- C++ si::IntervalMap<S,T> backed by libsuperintervals_b200.so.
//
// Host-side mirror of the reference class template for the batch overlap-query path
// (reference src/superintervals.hpp:45-150 storage/add/build/at, :501-1035 queries):
// same namespace, class name, public members (starts, ends, branch, data,
// start_sorted, end_sorted), method names, argument meaning, append-to-output
// behaviour and result order. All index construction and every query execute on the
// GPU through the C ABI (c_superintervals.h / superintervals_b200.h); payloads of
// any type T stay on the host and are addressed by position.
//
//   #include "superintervals.hpp"        // this file instead of the reference's
//   g++ -std=c++17 app.cpp -Iinclude -Lsuperintervals_b200 -lsuperintervals_b200
//
// Differences from the reference, all outside the accelerated path:
//   * S must be a 32-bit integer type (the reference's AVX2 path has the same
//     restriction, hpp:688; its C ABI is int32 only).
//   * The range classes IndexRange / KeyRange / ValueRange / ItemRange (hpp:153-494) exist with the
//     reference's names and iteration order; begin() runs the query on the device (one round trip
//     for the whole hit list) and the iterators then step through it, reading the host mirrors.
//     IntervalMapEytz is the same implementation under the reference's name.
//   * const queries may be called from several threads at once, as in the reference: per-thread
//     scratch here, and the library serialises the calls that reach one handle.
//   * The set algebra (hpp:1037-1390) is provided on top of the C ABI's device set
//     operations; `other` arguments must have been built.
//   * New: count_batch / search_values_batch / search_idxs_batch / search_keys_batch
//     (CSR), the throughput path.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>

#include "superintervals_b200.h"

namespace si {

template <typename S, typename T>
struct Interval {
    S start, end;
    T data;
    Interval() = default;
    Interval(S s, S e, T d) : start(s), end(e), data(d) {}
};

template <typename S, typename T>
class IntervalMap {
    static_assert(std::is_integral<S>::value && sizeof(S) == 4,
                  "superintervals_b200: coordinates must be a 32-bit integer type");

   public:
    std::vector<S> starts;
    std::vector<S> ends;
    std::vector<size_t> branch;
    std::vector<T> data;
    bool start_sorted, end_sorted;

    IntervalMap() : start_sorted(true), end_sorted(true), h_(nullptr) {}
    virtual ~IntervalMap() { destroySuperIntervals(h_); }
    IntervalMap(const IntervalMap&) = delete;
    IntervalMap& operator=(const IntervalMap&) = delete;
    IntervalMap(IntervalMap&& o) noexcept
        : starts(std::move(o.starts)), ends(std::move(o.ends)), branch(std::move(o.branch)), data(std::move(o.data)),
          start_sorted(o.start_sorted), end_sorted(o.end_sorted), h_(o.h_) {
        o.h_ = nullptr;
    }
    IntervalMap& operator=(IntervalMap&& o) noexcept {
        if (this != &o) {
            destroySuperIntervals(h_);
            starts = std::move(o.starts); ends = std::move(o.ends);
            branch = std::move(o.branch); data = std::move(o.data);
            start_sorted = o.start_sorted; end_sorted = o.end_sorted;
            h_ = o.h_; o.h_ = nullptr;
        }
        return *this;
    }

    void clear() noexcept { data.clear(); starts.clear(); ends.clear(); branch.clear(); }   // hpp:69-71
    void reserve(size_t n) { data.reserve(n); starts.reserve(n); ends.reserve(n); }
    size_t size() { return starts.size(); }

    void add(S start, S end, const T& value) {   // hpp:95-105
        if (start_sorted && !starts.empty()) {
            start_sorted = !(start < starts.back());
            if (start_sorted && start == starts.back() && end > ends.back()) end_sorted = false;
        }
        starts.push_back(start);
        ends.push_back(end);
        data.emplace_back(value);
    }

    // hpp:112-130: sort by (start asc, end desc) + branch array -- on the device.
    virtual void build() {
        if (starts.empty()) return;
        if (!h_) h_ = createSuperIntervals();
        clearSuperIntervals(h_);
        addIntervals(h_, reinterpret_cast<const int32_t*>(starts.data()),
                     reinterpret_cast<const int32_t*>(ends.data()), nullptr, starts.size());
        indexSuperIntervals(h_);
        // success is "this handle now carries a device index" -- not the process-wide sticky error flag,
        // which an unrelated earlier failure may have left set
        if (siIndexOf(h_) == nullptr) return;   // inspect si_b200_last_error_string(); the map stays not ready
        const size_t n = starts.size();
        std::vector<T> sorted;
        sorted.reserve(n);
        for (size_t i = 0; i < n; ++i) sorted.emplace_back(std::move(data[(size_t)h_->data[i]]));
        data.swap(sorted);
        std::copy(h_->starts, h_->starts + n, starts.begin());
        std::copy(h_->ends, h_->ends + n, ends.begin());
        branch.assign(h_->branch, h_->branch + n);
        start_sorted = end_sorted = true;
    }

    Interval<S, T> at(size_t index) const { return Interval<S, T>{starts[index], ends[index], data[index]}; }
    void at(size_t index, Interval<S, T>& itv) {
        itv.start = starts[index]; itv.end = ends[index]; itv.data = data[index];
    }

    // hpp:501-513: largest index with starts[index] <= value, SIZE_MAX if none.
    virtual size_t upper_bound(const S value) const noexcept {
        if (!ready()) return SIZE_MAX;
        return upperBound(h_, (int32_t)value);
    }

    // ---- single queries: appended to `found`, descending position order -------------------
    void search_values(const S start, const S end, std::vector<T>& found) const {   // hpp:551-579
        if (!ready()) return;
        const cIndexResult& r = hit_positions(start, end);
        for (size_t k = 0; k < r.size; ++k) found.push_back(data[(size_t)(uint32_t)r.data[k]]);
    }
    void search_values_large(const S start, const S end, std::vector<T>& found) const {   // hpp:588: same output
        search_values(start, end, found);
    }
    void search_point(const S point, std::vector<T>& found) const { search_values(point, point, found); }   // hpp:1013

    size_t count(const S start, const S end) const noexcept {   // hpp:651-825
        return ready() ? countOverlaps(h_, (int32_t)start, (int32_t)end) : 0;
    }
    size_t count_linear(const S start, const S end) const noexcept { return count(start, end); }   // hpp:623
    size_t count_large(const S start, const S end) const noexcept { return count(start, end); }    // hpp:834

    bool has_overlaps(const S start, const S end) const noexcept {   // hpp:865-871 (last candidate only)
        return ready() && anyOverlaps(h_, (int32_t)start, (int32_t)end);
    }

    // hpp:879-905. The reference inserts its first contiguous run ASCENDING and the rest
    // descending; reproduced here from the descending device result.
    void search_idxs(const S start, const S end, std::vector<size_t>& found) const {
        if (!ready()) return;
        const cIndexResult& r = hit_positions(start, end);
        const size_t first = found.size();
        for (size_t k = 0; k < r.size; ++k) found.push_back((size_t)(uint32_t)r.data[k]);
        if (r.size == 0) return;
        const size_t ub = (size_t)(std::upper_bound(starts.begin(), starts.end(), end) - starts.begin()) - 1;
        if (found[first] != ub) return;
        size_t run = 1;
        while (first + run < found.size() && found[first + run] + 1 == found[first + run - 1]) ++run;
        std::reverse(found.begin() + first, found.begin() + first + run);
    }

    void search_keys(const S start, const S end, std::vector<std::pair<S, S>>& found) const {   // hpp:913-938
        if (!ready()) return;
        cKeyResult r = createKeyResult();
        searchKeys(h_, (int32_t)start, (int32_t)end, &r);
        for (size_t k = 0; k < r.size; ++k) found.emplace_back((S)r.data[k].start, (S)r.data[k].end);
        destroyKeyResult(&r);
    }

    void search_items(const S start, const S end, std::vector<Interval<S, T>>& found) const {   // hpp:946-971
        if (!ready()) return;
        const cIndexResult& r = hit_positions(start, end);
        for (size_t k = 0; k < r.size; ++k) {
            const size_t j = (size_t)(uint32_t)r.data[k];
            found.emplace_back(starts[j], ends[j], data[j]);
        }
    }

    void coverage(const S start, const S end, std::pair<size_t, S>& cov_result) const {   // hpp:979-1006 (accumulates)
        if (!ready()) return;
        size_t c = 0;
        int32_t v = 0;
        ::coverage(h_, (int32_t)start, (int32_t)end, &c, &v);
        cov_result.first += c;
        cov_result.second += (S)v;
    }

    // ---- range objects (hpp:153-494): `for (auto x : map.search_idxs(s, e))` ------------------
    // Same class names, same descending iteration order, same "end() compares by exhaustion"
    // convention as the reference. The reference's iterators walk the branch array one hit at a
    // time on the host; here begin() asks the device for the query's whole hit list once and the
    // iterator steps through those positions, projecting each on the host mirrors.
   private:
    struct ProjIndex { using value_type = size_t;
        static void get(const IntervalMap* m, size_t j, value_type& v) { (void)m; v = j; } };
    struct ProjKey { using value_type = std::pair<S, S>;
        static void get(const IntervalMap* m, size_t j, value_type& v) { v.first = m->starts[j]; v.second = m->ends[j]; } };
    struct ProjValue { using value_type = T;
        static void get(const IntervalMap* m, size_t j, value_type& v) { v = m->data[j]; } };
    struct ProjItem { using value_type = Interval<S, T>;
        static void get(const IntervalMap* m, size_t j, value_type& v) { v.start = m->starts[j]; v.end = m->ends[j]; v.data = m->data[j]; } };

    template <typename Proj>
    class HitIterator {
       public:
        using value_type = typename Proj::value_type;
        HitIterator() = default;   // exhausted
        HitIterator(const IntervalMap* parent, std::shared_ptr<const std::vector<uint32_t>> hits)
            : parent_(parent), hits_(std::move(hits)) { load(); }
        const value_type& operator*() const { return value_; }
        HitIterator& operator++() { ++at_; load(); return *this; }
        bool operator!=(const HitIterator& o) const { return has_value_ != o.has_value_; }
        bool operator==(const HitIterator& o) const { return has_value_ == o.has_value_; }
       private:
        void load() {
            has_value_ = hits_ && at_ < hits_->size();
            if (has_value_) Proj::get(parent_, (size_t)(*hits_)[at_], value_);
        }
        const IntervalMap* parent_ = nullptr;
        std::shared_ptr<const std::vector<uint32_t>> hits_;
        size_t at_ = 0;
        bool has_value_ = false;
        value_type value_{};
    };
    template <typename Proj>
    class HitRange {
       public:
        HitRange(const IntervalMap* parent, S start, S end) : parent_(parent), start_(start), end_(end) {}
        HitIterator<Proj> begin() const {
            auto hits = std::make_shared<std::vector<uint32_t>>();
            if (parent_->ready()) {
                const cIndexResult& r = parent_->hit_positions(start_, end_);
                hits->assign(reinterpret_cast<const uint32_t*>(r.data), reinterpret_cast<const uint32_t*>(r.data) + r.size);
            }
            return HitIterator<Proj>(parent_, std::move(hits));
        }
        HitIterator<Proj> end() const { return HitIterator<Proj>(); }
       private:
        const IntervalMap* parent_;
        S start_, end_;
    };

   public:
    using IndexIterator = HitIterator<ProjIndex>;   // hpp:153
    using IndexRange = HitRange<ProjIndex>;         // hpp:211
    using KeyIterator = HitIterator<ProjKey>;       // hpp:237
    using KeyRange = HitRange<ProjKey>;             // hpp:297
    using ValueIterator = HitIterator<ProjValue>;   // hpp:324
    using ValueRange = HitRange<ProjValue>;         // hpp:382
    using ItemIterator = HitIterator<ProjItem>;     // hpp:408
    using ItemRange = HitRange<ProjItem>;           // hpp:470
    IndexRange search_idxs(S start, S end) const noexcept { return IndexRange(this, start, end); }     // hpp:233
    KeyRange search_keys(S start, S end) const noexcept { return KeyRange(this, start, end); }         // hpp:319
    ValueRange search_values(S start, S end) const noexcept { return ValueRange(this, start, end); }   // hpp:404
    ItemRange search_items(S start, S end) const noexcept { return ItemRange(this, start, end); }      // hpp:492

    // ---- batch queries (the throughput path; shaped after intervalmap.pyx:363-494) --------
    void count_batch(const S* qs, const S* qe, size_t n, std::vector<size_t>& counts) const {
        counts.assign(n, 0);
        if (!ready() || n == 0) return;
        countOverlapsBatch(h_, reinterpret_cast<const int32_t*>(qs), reinterpret_cast<const int32_t*>(qe), n, counts.data());
    }
    // CSR: query i's hits are positions[offsets[i] .. offsets[i+1]) (descending)
    void search_idxs_batch(const S* qs, const S* qe, size_t n, std::vector<size_t>& offsets,
                           std::vector<uint32_t>& positions) const {
        offsets.assign(n + 1, 0);
        positions.clear();
        if (!ready() || n == 0) return;
        cIndexResult r = createIndexResult();
        searchIdxsBatch(h_, reinterpret_cast<const int32_t*>(qs), reinterpret_cast<const int32_t*>(qe), n, offsets.data(), &r);
        positions.assign(reinterpret_cast<uint32_t*>(r.data), reinterpret_cast<uint32_t*>(r.data) + r.size);
        destroyIndexResult(&r);
    }
    void search_values_batch(const S* qs, const S* qe, size_t n, std::vector<size_t>& offsets,
                             std::vector<T>& values) const {
        std::vector<uint32_t> pos;
        search_idxs_batch(qs, qe, n, offsets, pos);
        values.clear();
        values.reserve(pos.size());
        for (uint32_t j : pos) values.push_back(data[j]);
    }
    void search_keys_batch(const S* qs, const S* qe, size_t n, std::vector<size_t>& offsets,
                           std::vector<std::pair<S, S>>& keys) const {
        offsets.assign(n + 1, 0);
        keys.clear();
        if (!ready() || n == 0) return;
        cKeyResult r = createKeyResult();
        searchKeysBatch(h_, reinterpret_cast<const int32_t*>(qs), reinterpret_cast<const int32_t*>(qe), n, offsets.data(), &r);
        keys.reserve(r.size);
        for (size_t k = 0; k < r.size; ++k) keys.emplace_back((S)r.data[k].start, (S)r.data[k].end);
        destroyKeyResult(&r);
    }

    cSuperIntervals* handle() const { return h_; }   // for the device-resident API: siIndexOf(handle())

    // ---- set algebra (hpp:1037-1390) ---------------------------------------------------------
    // Each returns a NEW map that is NOT indexed (call build() before querying it), like the
    // reference. The geometry -- the sort, the coalescing sweep, the batch of overlap queries
    // against `other` -- runs on the device through the C ABI's set operations
    // (c_superintervals.h "set operations"); payloads of type T are carried or folded on the
    // host in the reference's visiting order (hpp:1363-1390: stored order when both
    // sortedness flags hold, else (start, end) ascending).
    template <typename Combine>
    IntervalMap merge_overlaps(Combine combine) const {   // hpp:1057-1087
        IntervalMap out;
        if (starts.empty()) return out;
        CHandle t;
        to_handle(t, false);
        CHandle m{mergeOverlaps(t.h, nullptr)};
        const std::vector<size_t> order = visiting_order();
        size_t c = 0;
        bool open = false;
        T acc{};
        for (size_t idx : order) {
            while (open && c < m.h->size && starts[idx] > (S)m.h->ends[c]) {   // past the open cluster: flush
                out.add((S)m.h->starts[c], (S)m.h->ends[c], acc);
                ++c;
                open = false;
            }
            if (!open) { acc = data[idx]; open = true; }
            else acc = combine(acc, data[idx]);
        }
        if (open && c < m.h->size) out.add((S)m.h->starts[c], (S)m.h->ends[c], acc);
        return out;
    }
    IntervalMap merge_overlaps() const {
        return merge_overlaps([](const T& a, const T&) { return a; });
    }

    IntervalMap gaps(S lo, S hi, const T& fill = T{}) const {   // hpp:1104-1127
        IntervalMap out;
        CHandle t;
        to_handle(t, false);
        CHandle g{intervalGaps(t.h, (int32_t)lo, (int32_t)hi, 0)};
        for (size_t i = 0; i < g.h->size; ++i) out.add((S)g.h->starts[i], (S)g.h->ends[i], fill);
        return out;
    }

    template <typename Combine>
    IntervalMap union_with(const IntervalMap& other, Combine combine) const {   // hpp:1137-1147
        IntervalMap combined;
        combined.reserve(starts.size() + other.starts.size());
        for (size_t k = 0; k < starts.size(); ++k) combined.add(starts[k], ends[k], data[k]);
        for (size_t k = 0; k < other.starts.size(); ++k) combined.add(other.starts[k], other.ends[k], other.data[k]);
        return combined.merge_overlaps(combine);
    }
    IntervalMap union_with(const IntervalMap& other) const {
        return union_with(other, [](const T& a, const T&) { return a; });
    }

    // `other` must be built (it is queried through its index). Pieces are not coalesced.
    template <typename Combine>
    IntervalMap intersection(const IntervalMap& other, Combine combine) const {   // hpp:1164-1179
        IntervalMap out;
        if (starts.empty() || other.starts.empty()) return out;
        CHandle a, b;
        to_handle(a, false);
        other.to_handle(b, true);
        cIndexResult bd = createIndexResult();
        CHandle r{intersectionPairs(a.h, b.h, &bd)};
        for (size_t i = 0; i < r.h->size && i < bd.size; ++i)
            out.add((S)r.h->starts[i], (S)r.h->ends[i], combine(data[(size_t)r.h->data[i]], other.data[(size_t)bd.data[i]]));
        destroyIndexResult(&bd);
        return out;
    }
    IntervalMap intersection(const IntervalMap& other) const {
        return intersection(other, [](const T& a, const T&) { return a; });
    }

    IntervalMap difference(const IntervalMap& other) const {   // hpp:1189-1211
        IntervalMap out;
        if (starts.empty()) return out;
        CHandle a, b;
        to_handle(a, false);
        other.to_handle(b, true);
        CHandle r{::difference(a.h, b.h)};
        for (size_t i = 0; i < r.h->size; ++i) out.add((S)r.h->starts[i], (S)r.h->ends[i], data[(size_t)r.h->data[i]]);
        return out;
    }

    IntervalMap symmetric_difference(const IntervalMap& other) const {   // hpp:1219-1223
        IntervalMap a_minus_b = difference(other);
        IntervalMap b_minus_a = other.difference(*this);
        return a_minus_b.union_with(b_minus_a);
    }

    bool span(std::pair<S, S>& result) const {   // hpp:1230-1243
        if (starts.empty()) return false;
        CHandle t;
        to_handle(t, false);
        int32_t lo = 0, hi = 0;
        if (!intervalSpan(t.h, &lo, &hi)) return false;
        result = {(S)lo, (S)hi};
        return true;
    }

    IntervalMap expand(S left, S right, S lo = std::numeric_limits<S>::min(),
                       S hi = std::numeric_limits<S>::max()) const {   // hpp:1258-1290
        return resized(left, right, lo, hi, false);
    }
    IntervalMap flank(S left, S right, S lo = std::numeric_limits<S>::min(),
                      S hi = std::numeric_limits<S>::max()) const {   // hpp:1305-1330
        return resized(left, right, lo, hi, true);
    }

    template <typename Combine>
    IntervalMap unique(Combine combine) const {   // hpp:1341-1360
        IntervalMap out;
        if (starts.empty()) return out;
        CHandle t;
        to_handle(t, false);
        CHandle u{uniqueIntervals(t.h, nullptr)};   // how many distinct pairs: the fold below must agree
        const std::vector<size_t> order = visiting_order();
        size_t run = order[0];
        T acc = data[run];
        for (size_t k = 1; k < order.size(); ++k) {
            const size_t idx = order[k];
            if (starts[idx] == starts[run] && ends[idx] == ends[run]) acc = combine(acc, data[idx]);
            else { out.add(starts[run], ends[run], acc); run = idx; acc = data[idx]; }
        }
        out.add(starts[run], ends[run], acc);
        if (out.starts.size() != u.h->size) out.clear();   // device and host disagree: report nothing rather than guess
        return out;
    }
    IntervalMap unique() const {
        return unique([](const T& a, const T&) { return a; });
    }

   private:
    struct CHandle {   // owning wrapper of a C handle
        cSuperIntervals* h = nullptr;
        CHandle() = default;
        explicit CHandle(cSuperIntervals* p) : h(p) {}
        ~CHandle() { destroySuperIntervals(h); }
        CHandle(const CHandle&) = delete;
        CHandle& operator=(const CHandle&) = delete;
    };
    // the stored intervals as a C handle whose payload is the position in starts/ends/data
    void to_handle(CHandle& c, bool index) const {
        c.h = createSuperIntervals();
        if (!starts.empty())
            addIntervals(c.h, reinterpret_cast<const int32_t*>(starts.data()), reinterpret_cast<const int32_t*>(ends.data()),
                         nullptr, starts.size());
        if (index) indexSuperIntervals(c.h);
    }
    IntervalMap resized(S left, S right, S lo, S hi, bool flanks) const {
        IntervalMap out;
        if (starts.empty()) return out;
        CHandle t;
        to_handle(t, false);
        CHandle r{flanks ? flankIntervals(t.h, (int32_t)left, (int32_t)right, (int32_t)lo, (int32_t)hi)
                         : expandIntervals(t.h, (int32_t)left, (int32_t)right, (int32_t)lo, (int32_t)hi)};
        for (size_t i = 0; i < r.h->size; ++i) out.add((S)r.h->starts[i], (S)r.h->ends[i], data[(size_t)r.h->data[i]]);
        return out;
    }
    // hpp:1363-1390: stored order when add()/build() left both flags set, else (start, end) ascending
    std::vector<size_t> visiting_order() const {
        std::vector<size_t> order(starts.size());
        for (size_t k = 0; k < order.size(); ++k) order[k] = k;
        bool sorted = start_sorted && end_sorted;
        if (!sorted) {
            sorted = true;
            for (size_t k = 1; k < starts.size() && sorted; ++k)
                sorted = !(starts[k] < starts[k - 1] || (starts[k] == starts[k - 1] && ends[k] < ends[k - 1]));
        }
        if (!sorted)
            std::sort(order.begin(), order.end(), [this](size_t x, size_t y) {
                return starts[x] < starts[y] || (starts[x] == starts[y] && ends[x] < ends[y]);
            });
        return order;
    }
    bool ready() const { return h_ != nullptr && !starts.empty() && siIndexOf(h_) != nullptr; }
    // positions of the hits of one query, descending: one scratch buffer per calling thread, so that const
    // queries stay re-entrant like the reference's (hpp:551, 651)
    static cIndexResult& scratch() {
        struct Holder { cIndexResult r = {nullptr, 0, 0}; ~Holder() { destroyIndexResult(&r); } };
        thread_local Holder h;
        return h.r;
    }
    cIndexResult& hit_positions(const S start, const S end) const {
        cIndexResult& r = scratch();
        r.size = 0;
        searchIdxs(h_, (int32_t)start, (int32_t)end, &r);
        return r;
    }
    cSuperIntervals* h_ = nullptr;
};

// hpp:1457-1535: IntervalMapEytz keeps the starts in Eytzinger order for a cache-friendlier CPU
// upper_bound; its results are identical to IntervalMap's. On the device that layout question does
// not arise, so the name is provided for source compatibility and is the same implementation.
template <typename S, typename T>
class IntervalMapEytz : public IntervalMap<S, T> {};

}  // namespace si
