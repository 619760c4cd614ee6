// bedref_shim.cpp -- the reference's OWN BED loader, compiled in place (TEST INFRASTRUCTURE, not product code).
// The reference has no library entry point for BED text: its loader is Bench::load_intervals in
// test/bench.cpp:67-102 (std::getline per line, '\t' fields, std::stoi, min/max normalisation, chrom "chr1" kept).
// This file includes that translation unit UNMODIFIED (its main() renamed away by the preprocessor) and exposes
// the loader over a C ABI so the BED oracle (oracle/bed_oracle.py) and the device parser (csrc/bed.cu) can be
// pinned to the reference's own output on the same text (tests/test_bed_oracle.py, tests/test_gpu_bed.py).
// Built by oracle/Makefile into oracle/_ref/libsi_bedref.so; nothing under superintervals_b200/ links it.
#define main si_ref_bench_main_unused_
#include "bench.cpp"
#undef main

#include <cstdint>
#include <cstring>

extern "C" {
// Loads both files with the reference loader. Returns the number of intervals; *n_queries = number of queries.
// Up to cap / qcap records are written to the output columns.
size_t si_ref_load_intervals(const char* intervals_path, const char* queries_path, int32_t* starts, int32_t* ends,
                             size_t cap, int32_t* q_starts, int32_t* q_ends, size_t qcap, size_t* n_queries) {
    std::vector<Bench::BedInterval> a, b;
    Bench::load_intervals(intervals_path, queries_path, a, b);
    for (size_t i = 0; i < a.size() && i < cap; ++i) { starts[i] = a[i].start; ends[i] = a[i].end; }
    for (size_t i = 0; i < b.size() && i < qcap; ++i) { q_starts[i] = b[i].start; q_ends[i] = b[i].end; }
    *n_queries = b.size();
    return a.size();
}
}
