// bed_cpu.cpp -- CPU BASELINE for BED ingest (TEST / BENCH INFRASTRUCTURE, not product code).
// Restates the reference's loader loop (reference test/bench.cpp:67-102) so it can be timed on
// the host beside the device tokeniser: std::getline per line, std::istringstream split on '\t',
// std::stoi for start and end, min/max normalisation, every chrom kept (bench.cpp keeps "chr1").
// Reads from memory (std::istringstream over the buffer) so that file I/O is excluded on both sides.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

extern "C" size_t si_bed_parse_cpu(const char* text, size_t bytes, int32_t* starts, int32_t* ends, int32_t* contig,
                                   size_t cap) {
    std::istringstream in(std::string(text, bytes));
    std::unordered_map<std::string, int32_t> ids;
    std::string line, token;
    size_t n = 0;
    while (std::getline(in, line)) {
        std::istringstream iss(line);
        std::getline(iss, token, '\t');
        auto it = ids.find(token);
        const int32_t cid = it == ids.end() ? ids.emplace(token, (int32_t)ids.size()).first->second : it->second;
        std::getline(iss, token, '\t');
        const int start = std::stoi(token);
        std::getline(iss, token, '\t');
        const int end = std::stoi(token);
        if (n < cap) {
            starts[n] = std::min(start, end);
            ends[n] = std::max(start, end);
            contig[n] = cid;
        }
        ++n;
    }
    return n;
}
