// ref_shim.cpp -- thin extern "C" handle around the UNMODIFIED reference
// (si::IntervalMap<int,int> from /root/reference/src/superintervals.hpp, found
// through -I at compile time; no reference source is copied into this repo).
//
// TEST / BASELINE INFRASTRUCTURE ONLY: built by oracle/Makefile into
// oracle/_ref/libsi_ref_{native,v3}.so. Used to (a) pin oracle/si_oracle.c,
// (b) generate tests/golden fixtures, (c) time the reference's own CPU path on
// the GPU box's host cores (bench.py cpu_baseline / --impl reference).
// The reference has no threads of its own; *_mt entry points fan its const,
// re-entrant query methods (hpp:551,651) over std::thread on contiguous query
// chunks, each thread with its own output vector (SURVEY 8b "Threading").
#include "superintervals.hpp"

#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

using Map = si::IntervalMap<int, int>;

namespace {
struct Handle {
    Map map;
    std::vector<int> values;       // last batch search result (CSR values)
    std::vector<int> keys;         // last batch search_keys result (2 per hit)
    std::vector<uint64_t> idxs;    // last batch search_idxs result
};

template <class F>
void fan_out(size_t n, int threads, F&& body) {
    if (threads <= 1 || n < 2) { body(0, size_t(0), n); return; }
    std::vector<std::thread> pool;
    size_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t lo = std::min(n, size_t(t) * chunk), hi = std::min(n, lo + chunk);
        pool.emplace_back([&body, t, lo, hi] { body(t, lo, hi); });
    }
    for (auto& th : pool) th.join();
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

extern "C" {

void* si_ref_create() { return new Handle(); }
void si_ref_destroy(void* h) { delete static_cast<Handle*>(h); }
void si_ref_clear(void* h) {
    auto* H = static_cast<Handle*>(h);
    H->map.clear();
    // NOTE: reference clear() (hpp:69-71) does not reset the sortedness flags.
}
void si_ref_add(void* h, int s, int e, int v) { static_cast<Handle*>(h)->map.add(s, e, v); }
void si_ref_add_many(void* h, const int* s, const int* e, const int* v, size_t n) {
    auto& m = static_cast<Handle*>(h)->map;
    m.reserve(m.size() + n);
    for (size_t i = 0; i < n; ++i) m.add(s[i], e[i], v ? v[i] : int(i));
}
void si_ref_build(void* h) { static_cast<Handle*>(h)->map.build(); }
size_t si_ref_size(void* h) { return static_cast<Handle*>(h)->map.size(); }
int si_ref_flags(void* h) {
    auto& m = static_cast<Handle*>(h)->map;
    return (m.start_sorted ? 1 : 0) | (m.end_sorted ? 2 : 0);
}
// Copies the public member arrays (hpp:49-60). Any pointer may be null.
void si_ref_export(void* h, int* starts, int* ends, int* data, uint64_t* branch) {
    auto& m = static_cast<Handle*>(h)->map;
    size_t n = m.starts.size();
    if (starts) std::memcpy(starts, m.starts.data(), n * sizeof(int));
    if (ends) std::memcpy(ends, m.ends.data(), n * sizeof(int));
    if (data) std::memcpy(data, m.data.data(), n * sizeof(int));
    if (branch) for (size_t i = 0; i < m.branch.size(); ++i) branch[i] = m.branch[i];
}

uint64_t si_ref_upper_bound(void* h, int v) { return static_cast<Handle*>(h)->map.upper_bound(v); }
uint64_t si_ref_count(void* h, int s, int e) { return static_cast<Handle*>(h)->map.count(s, e); }
int si_ref_has_overlaps(void* h, int s, int e) { return static_cast<Handle*>(h)->map.has_overlaps(s, e); }

// mode: 0 = count (SIMD path, hpp:651), 1 = count_linear (hpp:623), 2 = count_large (hpp:834)
void si_ref_count_batch(void* h, const int* qs, const int* qe, size_t nq, uint64_t* out,
                        int mode, int threads) {
    const Map& m = static_cast<Handle*>(h)->map;
    fan_out(nq, threads, [&](int, size_t lo, size_t hi) {
        for (size_t q = lo; q < hi; ++q)
            out[q] = mode == 0 ? m.count(qs[q], qe[q])
                   : mode == 1 ? m.count_linear(qs[q], qe[q]) : m.count_large(qs[q], qe[q]);
    });
}

void si_ref_has_overlaps_batch(void* h, const int* qs, const int* qe, size_t nq, uint8_t* out) {
    const Map& m = static_cast<Handle*>(h)->map;
    for (size_t q = 0; q < nq; ++q) out[q] = m.has_overlaps(qs[q], qe[q]);
}

// Runs search_values (mode 0, hpp:551) or search_values_large (mode 1, hpp:588)
// per query, exactly as intervalmap.pyx:485-492 loops, and keeps the ragged
// result as CSR: offsets_out[nq+1]; values retrievable with si_ref_take_values.
uint64_t si_ref_search_values_batch(void* h, const int* qs, const int* qe, size_t nq,
                                    uint64_t* offsets_out, int mode, int threads) {
    auto* H = static_cast<Handle*>(h);
    const Map& m = H->map;
    int T = threads < 1 ? 1 : threads;
    std::vector<std::vector<int>> parts(T);
    fan_out(nq, T, [&](int t, size_t lo, size_t hi) {
        auto& v = parts[t];
        for (size_t q = lo; q < hi; ++q) {
            size_t before = v.size();
            if (mode == 0) m.search_values(qs[q], qe[q], v);   // appends (Q4)
            else m.search_values_large(qs[q], qe[q], v);
            offsets_out[q + 1] = v.size() - before;            // per-query size for now
        }
    });
    offsets_out[0] = 0;
    for (size_t q = 0; q < nq; ++q) offsets_out[q + 1] += offsets_out[q];
    H->values.clear();
    H->values.reserve(offsets_out[nq]);
    for (auto& v : parts) H->values.insert(H->values.end(), v.begin(), v.end());
    return offsets_out[nq];
}
void si_ref_take_values(void* h, int* out) {
    auto* H = static_cast<Handle*>(h);
    if (!H->values.empty()) std::memcpy(out, H->values.data(), H->values.size() * sizeof(int));
}

uint64_t si_ref_search_keys_batch(void* h, const int* qs, const int* qe, size_t nq,
                                  uint64_t* offsets_out) {
    auto* H = static_cast<Handle*>(h);
    std::vector<std::pair<int, int>> found;
    H->keys.clear();
    offsets_out[0] = 0;
    for (size_t q = 0; q < nq; ++q) {
        found.clear();
        H->map.search_keys(qs[q], qe[q], found);               // hpp:913
        for (auto& p : found) { H->keys.push_back(p.first); H->keys.push_back(p.second); }
        offsets_out[q + 1] = offsets_out[q] + found.size();
    }
    return offsets_out[nq];
}
void si_ref_take_keys(void* h, int* out) {
    auto* H = static_cast<Handle*>(h);
    if (!H->keys.empty()) std::memcpy(out, H->keys.data(), H->keys.size() * sizeof(int));
}

// search_idxs into a vector (hpp:879) -- first run ASCENDING, rest descending (Q2).
uint64_t si_ref_search_idxs_batch(void* h, const int* qs, const int* qe, size_t nq,
                                  uint64_t* offsets_out) {
    auto* H = static_cast<Handle*>(h);
    std::vector<size_t> found;
    H->idxs.clear();
    offsets_out[0] = 0;
    for (size_t q = 0; q < nq; ++q) {
        found.clear();
        H->map.search_idxs(qs[q], qe[q], found);
        for (size_t v : found) H->idxs.push_back(v);
        offsets_out[q + 1] = offsets_out[q] + found.size();
    }
    return offsets_out[nq];
}
void si_ref_take_idxs(void* h, uint64_t* out) {
    auto* H = static_cast<Handle*>(h);
    if (!H->idxs.empty()) std::memcpy(out, H->idxs.data(), H->idxs.size() * sizeof(uint64_t));
}

void si_ref_coverage_batch(void* h, const int* qs, const int* qe, size_t nq,
                           uint64_t* count_out, int* cov_out) {
    const Map& m = static_cast<Handle*>(h)->map;
    for (size_t q = 0; q < nq; ++q) {
        std::pair<size_t, int> c{0, 0};
        m.coverage(qs[q], qe[q], c);                           // hpp:979
        count_out[q] = c.first;
        cov_out[q] = c.second;
    }
}

// ---- timing legs: the loops of test/bench.cpp:217-250, optionally fanned out.
// Return wall seconds (steady_clock) and the bench.cpp "found" checksum.
double si_ref_time_count(void* h, const int* qs, const int* qe, size_t nq, int threads,
                         uint64_t* found_out) {
    const Map& m = static_cast<Handle*>(h)->map;
    int T = threads < 1 ? 1 : threads;
    std::vector<uint64_t> acc(T, 0);
    double t0 = now_s();
    fan_out(nq, T, [&](int t, size_t lo, size_t hi) {
        uint64_t f = 0;
        for (size_t q = lo; q < hi; ++q) f += m.count(qs[q], qe[q]);   // bench.cpp:240-242
        acc[t] = f;
    });
    double dt = now_s() - t0;
    uint64_t f = 0;
    for (auto v : acc) f += v;
    *found_out = f;
    return dt;
}

double si_ref_time_search_values(void* h, const int* qs, const int* qe, size_t nq, int threads,
                                 uint64_t* found_out) {
    const Map& m = static_cast<Handle*>(h)->map;
    int T = threads < 1 ? 1 : threads;
    std::vector<uint64_t> acc(T, 0);
    double t0 = now_s();
    fan_out(nq, T, [&](int t, size_t lo, size_t hi) {
        std::vector<int> a;
        a.reserve(10000);                                              // bench.cpp:203-204
        uint64_t f = 0;
        for (size_t q = lo; q < hi; ++q) {                             // bench.cpp:219-222
            m.search_values(qs[q], qe[q], a);
            f += a.size();
            a.clear();
        }
        acc[t] = f;
    });
    double dt = now_s() - t0;
    uint64_t f = 0;
    for (auto v : acc) f += v;
    *found_out = f;
    return dt;
}

// add() x n + build() as bench.cpp:208-213 does; returns seconds.
double si_ref_time_build(void* h, const int* s, const int* e, size_t n) {
    auto& m = static_cast<Handle*>(h)->map;
    m.clear();
    m.start_sorted = true;
    m.end_sorted = true;
    double t0 = now_s();
    for (size_t i = 0; i < n; ++i) m.add(s[i], e[i], int(i));
    m.build();
    return now_s() - t0;
}

}  // extern "C"
