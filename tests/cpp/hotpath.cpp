// hotpath.cpp -- the reference's hot-path unit tests (reference test/tests.cpp:63-238), restated
// against include/superintervals.hpp (GPU-backed si::IntervalMap), plus the batch API.
// Built and run by tests/test_gpu_cpp.py on the GPU box; compiled (not run) on CPU.
#include "superintervals.hpp"

#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

using Map = si::IntervalMap<int, int>;

#define CHECK(cond)                                                                   \
    do {                                                                              \
        if (!(cond)) {                                                                \
            std::fprintf(stderr, "FAILED %s:%d: %s  [%s]\n", __FILE__, __LINE__, #cond, \
                         si_b200_last_error_string());                                \
            std::exit(1);                                                             \
        }                                                                             \
    } while (0)

static void test_basics() {   // tests.cpp:63-107
    Map itv;
    CHECK(!itv.has_overlaps(0, 1));
    itv.add(10, 20, 0);
    itv.add(11, 12, -1);
    itv.add(13, 14, -1);
    itv.add(15, 16, -1);
    itv.add(25, 29, 4);
    itv.build();
    CHECK(itv.has_overlaps(17, 30));
    CHECK(!itv.has_overlaps(1, 3));
    std::vector<int> values;
    itv.search_values(17, 30, values);
    CHECK(values.size() == 2 && values[0] == 4 && values[1] == 0);
    CHECK(itv.count(17, 30) == 2);
    std::vector<size_t> idxs;
    itv.search_idxs(17, 30, idxs);
    CHECK(idxs.size() == 2 && idxs[0] == 4 && idxs[1] == 0);
    std::vector<std::pair<int, int>> keys;
    itv.search_keys(17, 30, keys);
    CHECK(keys[0].first == 25 && keys[1].second == 20);
    std::pair<size_t, int> cov{0, 0};
    itv.coverage(10, 29, cov);
    CHECK(cov.second == 17);
    CHECK(itv.branch.size() == 5 && itv.branch[0] == SIZE_MAX && itv.branch[1] == 0 && itv.branch[4] == SIZE_MAX);
    CHECK(itv.upper_bound(12) == 1 && itv.upper_bound(9) == SIZE_MAX);
}

static void test_iteration() {   // tests.cpp:113-133
    Map itv;
    itv.add(1, 10, 0);
    itv.build();
    size_t count = 0;
    int last_data = -1;
    for (const size_t i : itv.search_idxs(5, 11)) { last_data = itv.data[i]; ++count; }
    CHECK(count == 1 && last_data == 0);
    for (const auto& interval : itv.search_items(5, 11)) CHECK(interval.start == 1 && interval.end == 10 && interval.data == 0);
}

static void test_overlap_queries() {   // tests.cpp:140-188
    std::vector<int> a;
    {
        Map itv;
        itv.add(0, 250000, 0);
        for (int s : {55, 115, 130, 281, 639, 842, 999, 1094, 1157, 1161, 1265, 1532, 1590, 1665, 1945, 2384, 2515})
            itv.add(s, s + 1000, -1);
        itv.build();
        itv.search_values(1377, 2377, a);
        CHECK(a.back() == 0 && a.size() == 12);
        a.clear();
    }
    {
        Map itv;
        itv.add(3, 40, 0); itv.add(4, 5, 4); itv.add(6, 7, 4); itv.add(10, 31, 5); itv.add(31, 32, 5);
        itv.build();
        itv.search_values(31, 32, a); CHECK(a.size() == 3); a.clear();
        itv.search_values(10, 11, a); CHECK(a.size() == 2); a.clear();
        itv.search_values(4, 7, a);   CHECK(a.size() == 3); a.clear();
    }
    {
        Map itv;
        itv.add(3, 40, 0); itv.add(3, 40, 4); itv.add(3, 40, 4); itv.add(3, 4, 4);
        itv.add(35, 50, 4); itv.add(40, 400, 5); itv.add(40, 400, 4);
        itv.build();
        itv.search_values(38, 41, a); CHECK(a.size() == 6); a.clear();
        itv.search_values(41, 42, a); CHECK(a.size() == 3); a.clear();
    }
}

static void test_coverage() {   // tests.cpp:194-213
    Map itv;
    itv.add(1, 100, 0); itv.add(30, 200, 7); itv.add(40, 50, 6); itv.add(60, 70, 7);
    itv.build();
    std::vector<int> a;
    itv.search_values(55, 65, a);
    CHECK(a.size() == 3);
    std::pair<size_t, int> cov{0, 0};
    itv.coverage(55, 65, cov);
    CHECK(cov.first == 3 && cov.second == 25);
}

static void test_edge_cases() {   // tests.cpp:219-238
    {
        Map itv;
        itv.build();
        CHECK(itv.count(1, 5) == 0);
        CHECK(!itv.has_overlaps(1, 5));
    }
    {
        Map itv;
        itv.add(1, 10, 1);
        itv.build();
        CHECK(itv.count(1, 5) == 1);
        CHECK(itv.count(11, 20) == 0);
    }
}

static void test_quirks() {
    Map itv;   // Q1: has_overlaps looks at the last candidate only (hpp:869-870)
    itv.add(1, 100, 0); itv.add(5, 6, 1);
    itv.build();
    CHECK(!itv.has_overlaps(50, 60) && itv.count(50, 60) == 1);
    Map q2;    // Q2: vector search_idxs has its first run ascending (hpp:892-895)
    q2.add(1, 100, 0); q2.add(2, 3, 1); q2.add(10, 50, 2); q2.add(11, 50, 3); q2.add(12, 50, 4);
    q2.build();
    std::vector<size_t> idx;
    q2.search_idxs(20, 30, idx);
    CHECK((idx == std::vector<size_t>{2, 3, 4, 0}));
    std::vector<int> v;
    q2.search_values(20, 30, v);
    CHECK((v == std::vector<int>{4, 3, 2, 0}));
}

static void test_batch_and_payload_types() {
    // non-int payloads ride on the host; batch CSR equals the per-query calls
    si::IntervalMap<int, std::pair<int, double>> m;
    const int n = 5000;
    unsigned x = 12345;
    auto rnd = [&x]() { x = x * 1664525u + 1013904223u; return (int)(x >> 8); };
    for (int i = 0; i < n; ++i) { int s = rnd() % 100000; m.add(s, s + rnd() % 700, {i, i * 0.5}); }
    m.build();
    std::vector<int> qs, qe;
    for (int i = 0; i < 3000; ++i) { int s = rnd() % 100000; qs.push_back(s); qe.push_back(s + rnd() % 900); }
    std::vector<size_t> counts, off;
    std::vector<std::pair<int, double>> vals;
    m.count_batch(qs.data(), qe.data(), qs.size(), counts);
    m.search_values_batch(qs.data(), qe.data(), qs.size(), off, vals);
    CHECK(off.back() == std::accumulate(counts.begin(), counts.end(), size_t(0)) && vals.size() == off.back());
    for (size_t q = 0; q < qs.size(); q += 97) {
        std::vector<std::pair<int, double>> one;
        m.search_values(qs[q], qe[q], one);
        CHECK(one.size() == counts[q] && m.count(qs[q], qe[q]) == counts[q]);
        for (size_t k = 0; k < one.size(); ++k) CHECK(one[k] == vals[off[q] + k]);
        size_t brute = 0;
        for (size_t j = 0; j < m.starts.size(); ++j) brute += (m.starts[j] <= qe[q] && m.ends[j] >= qs[q]);
        CHECK(brute == counts[q]);
    }
}

// ---- set operations: the reference's known answers (test/tests.cpp:259-378) ------------------
static std::vector<std::pair<int, int>> geometry(Map& m) {
    std::vector<std::pair<int, int>> v;
    for (size_t i = 0; i < m.size(); ++i) v.emplace_back(m.starts[i], m.ends[i]);
    return v;
}

static void test_set_operations() {
    {   // tests.cpp:259-278
        Map a; a.add(1, 5, 0); a.add(3, 8, 1); a.add(20, 30, 2); a.build();
        Map merged = a.merge_overlaps();
        auto v = geometry(merged);
        CHECK(v.size() == 2 && v[0] == std::make_pair(1, 8) && v[1] == std::make_pair(20, 30));
        Map summed = a.merge_overlaps([](const int& x, const int& y) { return x + y; });
        CHECK(summed.at(0).data == 1);
    }
    {   // tests.cpp:280-293
        Map a; a.add(10, 20, 0); a.add(30, 40, 1); a.build();
        Map g = a.gaps(0, 50);
        auto v = geometry(g);
        CHECK(v.size() == 3 && v[0] == std::make_pair(0, 9) && v[1] == std::make_pair(21, 29) && v[2] == std::make_pair(41, 50));
    }
    {   // tests.cpp:295-306
        Map a, b; a.add(1, 10, 0); b.add(5, 25, 1); a.build(); b.build();
        Map u = a.union_with(b);
        auto v = geometry(u);
        CHECK(v.size() == 1 && v[0] == std::make_pair(1, 25));
    }
    {   // tests.cpp:308-329
        Map a, b; a.add(1, 10, 0); a.add(20, 30, 1); b.add(5, 25, 2); a.build(); b.build();
        Map inter = a.intersection(b);
        auto v = geometry(inter);
        CHECK(v.size() == 2 && v[0] == std::make_pair(5, 10) && v[1] == std::make_pair(20, 25));
        inter.build();
        CHECK(inter.has_overlaps(7, 7) && !inter.has_overlaps(15, 15));
    }
    {   // tests.cpp:331-343
        Map a, b; a.add(1, 10, 0); b.add(4, 6, 1); a.build(); b.build();
        Map d = a.difference(b);
        auto v = geometry(d);
        CHECK(v.size() == 2 && v[0] == std::make_pair(1, 3) && v[1] == std::make_pair(7, 10));
    }
    {   // tests.cpp:345-357
        Map a, b; a.add(1, 10, 0); b.add(5, 15, 1); a.build(); b.build();
        Map x = a.symmetric_difference(b);
        auto v = geometry(x);
        CHECK(v.size() == 2 && v[0] == std::make_pair(1, 4) && v[1] == std::make_pair(11, 15));
    }
    {   // tests.cpp:359-373
        Map a; a.add(10, 20, 0); a.add(5, 8, 1); a.add(15, 50, 2); a.build();
        std::pair<int, int> sp;
        CHECK(a.span(sp) && sp.first == 5 && sp.second == 50);
        Map empty;
        CHECK(!empty.span(sp));
    }
    {   // expand / flank / unique (hpp:1258-1360; no reference unit test: checked against their definitions)
        Map a; a.add(10, 20, 7); a.add(10, 20, 8); a.add(30, 31, 9); a.build();
        Map e = a.expand(2, 3, 9, 33);
        auto v = geometry(e);
        CHECK(v.size() == 3 && v[0] == std::make_pair(9, 23) && v[2] == std::make_pair(28, 33));
        Map f = a.flank(2, 2);
        CHECK(f.size() == 6 && f.starts[0] == 8 && f.ends[0] == 9 && f.starts[1] == 21 && f.ends[1] == 22);
        Map u = a.unique([](const int& x, const int& y) { return x + y; });
        CHECK(u.size() == 2 && u.at(0).data == 15 && u.at(1).data == 9);
    }
}

// ---- the four range classes of the reference (hpp:153-494) under their own names ---------------
static void test_range_classes() {
    Map m;
    m.add(1, 10, 100); m.add(3, 4, 101); m.add(5, 20, 102); m.add(30, 40, 103);
    m.build();
    static_assert(std::is_same<decltype(m.search_idxs(5, 11)), Map::IndexRange>::value, "search_idxs(s, e) -> IndexRange");
    static_assert(std::is_same<decltype(m.search_keys(5, 11)), Map::KeyRange>::value, "search_keys(s, e) -> KeyRange");
    static_assert(std::is_same<decltype(m.search_values(5, 11)), Map::ValueRange>::value, "search_values(s, e) -> ValueRange");
    static_assert(std::is_same<decltype(m.search_items(5, 11)), Map::ItemRange>::value, "search_items(s, e) -> ItemRange");
    std::vector<size_t> idx;
    for (size_t i : m.search_idxs(5, 11)) idx.push_back(i);                    // descending positions
    CHECK(idx.size() == 2 && idx[0] == 2 && idx[1] == 0);
    std::vector<int> vals;
    for (const int& v : m.search_values(5, 11)) vals.push_back(v);
    CHECK(vals.size() == 2 && vals[0] == 102 && vals[1] == 100);
    std::vector<std::pair<int, int>> keys;
    for (const auto& k : m.search_keys(4, 4)) keys.push_back(k);
    CHECK(keys.size() == 2 && keys[0] == std::make_pair(3, 4) && keys[1] == std::make_pair(1, 10));
    size_t items = 0;
    for (const auto& it : m.search_items(35, 36)) { CHECK(it.start == 30 && it.end == 40 && it.data == 103); ++items; }
    CHECK(items == 1);
    Map::IndexRange none = m.search_idxs(21, 29);
    CHECK(!(none.begin() != none.end()));
    Map::IndexIterator it = m.search_idxs(0, 100).begin();                      // explicit iterator use, as hpp:153-209 allows
    size_t seen = 0;
    for (; it != m.search_idxs(0, 100).end(); ++it) ++seen;
    CHECK(seen == 4);
}

// ---- ownership: a built map survives a real move (vector growth, std::move) ------------------------
static void test_moves() {
    std::vector<Map> maps;
    for (int k = 0; k < 5; ++k) {   // push_back reallocates and move-constructs the earlier maps
        Map m;
        m.add(10 * k, 10 * k + 5, k); m.add(10 * k + 2, 10 * k + 3, 100 + k);
        m.build();
        maps.push_back(std::move(m));
    }
    for (int k = 0; k < 5; ++k) CHECK(maps[k].count(10 * k + 2, 10 * k + 2) == 2 && maps[k].count(10 * k + 7, 10 * k + 8) == 0);
    Map b = std::move(maps[3]);
    CHECK(b.count(32, 33) == 2);
    Map c;
    c = std::move(b);
    std::vector<int> v;
    c.search_values(30, 31, v);
    CHECK(v.size() == 1 && v[0] == 3);
}

// ---- a stale process-wide error must not make build() skip its host-side reordering ------------------
static void test_build_ignores_a_stale_error() {
    Map broken;
    CHECK(broken.count(1, 2) == 0);
    cIndexResult r = createIndexResult();
    cSuperIntervals* raw = createSuperIntervals();
    addInterval(raw, 1, 2, 0);
    searchValues(raw, 1, 2, &r);                       // not indexed: latches "handle not indexed"
    CHECK(si_b200_last_error() != 0);
    Map m;
    m.add(50, 60, 7); m.add(10, 20, 8); m.add(30, 40, 9);
    m.build();                                          // the sticky error is still set
    CHECK(m.starts[0] == 10 && m.data[0] == 8 && m.data[2] == 7);
    std::vector<int> v;
    m.search_values(55, 56, v);
    CHECK(v.size() == 1 && v[0] == 7);
    si_b200_clear_error();
    destroyIndexResult(&r);
    destroySuperIntervals(raw);
}

// ---- const queries from several threads at once (reference: safe, hpp:551,651) -----------
static void test_concurrent_const_queries() {
    Map m;
    unsigned x = 777;
    auto rnd = [&x]() { x = x * 1664525u + 1013904223u; return (int)(x >> 8); };
    for (int i = 0; i < 20000; ++i) { int s = rnd() % 200000; m.add(s, s + rnd() % 900, i); }
    m.build();
    std::vector<int> qs, qe;
    for (int i = 0; i < 400; ++i) { int s = rnd() % 200000; qs.push_back(s); qe.push_back(s + rnd() % 1500); }
    std::vector<size_t> want(qs.size());
    std::vector<std::vector<int>> want_v(qs.size());
    for (size_t q = 0; q < qs.size(); ++q) { want[q] = m.count(qs[q], qe[q]); m.search_values(qs[q], qe[q], want_v[q]); }
    const Map& cm = m;
    std::vector<int> bad(4, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < 4; ++t)
        th.emplace_back([&, t]() {
            for (int rep = 0; rep < 3; ++rep)
                for (size_t q = t; q < qs.size(); q += 2) {      // threads overlap on the same queries
                    std::vector<int> v;
                    cm.search_values(qs[q], qe[q], v);
                    size_t n_it = 0;
                    for (size_t i : cm.search_idxs(qs[q], qe[q])) { (void)i; ++n_it; }
                    if (cm.count(qs[q], qe[q]) != want[q] || v != want_v[q] || n_it != want[q]) ++bad[t];
                    if (cm.has_overlaps(qs[q], qe[q]) && want[q] == 0) ++bad[t];
                }
        });
    for (auto& t : th) t.join();
    CHECK(bad[0] + bad[1] + bad[2] + bad[3] == 0);
}

int main() {
    test_basics();
    test_iteration();
    test_overlap_queries();
    test_coverage();
    test_edge_cases();
    test_quirks();
    test_batch_and_payload_types();
    test_range_classes();
    test_moves();
    test_concurrent_const_queries();
    CHECK(si_b200_last_error() == 0);
    test_build_ignores_a_stale_error();
    CHECK(si_b200_last_error() == 0);
    std::printf("All query tests passed\n");
    {   // hpp:1457-1535: the Eytzinger variant answers exactly like the plain map
        si::IntervalMapEytz<int, int> ez;
        ez.add(10, 20, 0); ez.add(11, 12, 1); ez.add(25, 29, 2);
        ez.build();
        CHECK(ez.count(12, 26) == 3 && ez.has_overlaps(26, 40) && ez.upper_bound(24) == 1);
    }
    test_set_operations();
    CHECK(si_b200_last_error() == 0);
    std::printf("All set operation tests passed\n");
    return 0;
}
