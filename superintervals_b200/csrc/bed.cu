// bed.cu -- BED text -> (contig id, start, end) columns on the device: the step BEFORE the
// query path. The reference's callers tokenise BED on one host thread, one
// line at a time (test/bench.cpp:67-102: tab-separated chrom, start, end through std::stoi;
// examples/bed-intersect-si.rs:63-123: a per-contig container keyed by the chrom field). At
// the path's sizes (10^8 .. 10^9 records) that step, not the queries, is the wall clock.
//
// Device pipeline over the whole text buffer (one upload):
//   1. bd_count_lines   each thread scans a 64-byte slice and counts '\n'
//   2. scan             -> where each slice's lines go in the line table
//   3. bd_line_starts   byte offset of every line start
//   4. bd_parse         one thread per line: FNV-1a hash of the chrom token, std::stoi
//                       semantics for fields 2 and 3 (leading blanks, optional sign, digits,
//                       trailing junk ignored; no digits or overflow -> the line is skipped,
//                       where the reference would throw), first sight of each chrom recorded
//                       in a small open-addressing table (atomicCAS / atomicMin)
//   5. host             names the distinct chroms in order of first appearance (a handful)
//   6. bd_finish        valid lines -> scan -> compact (contig id, start, end) in line order
//   7. (optional)       group by contig: stable radix sort of the record numbers by contig id, gather,
//                       per-contig offsets -- the per-chrom containers of the reference's callers
#include "../../include/superintervals_b200.h"

#include "common.cuh"
#include "index.cuh"
#include "host_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits.h>
#include <string>
#include <vector>

using namespace sib;

extern "C" int si_b200_stable_order_(siIndex* ix, const int32_t* d_key, size_t n, int key_bits, uint32_t* d_perm, void* stream);

namespace {

constexpr int BD_THREADS = 256;
constexpr uint32_t BD_SLICE = 64;             // bytes of text per thread in the line passes
constexpr uint32_t BD_TABLE = 1u << 16;       // chrom table slots (open addressing); at most half are used
constexpr uint64_t BD_EMPTY = 0ull;

struct ChromSlot {
    unsigned long long hash;    // 0 = empty (a real hash of 0 is remapped to 1)
    unsigned long long first;   // smallest line number carrying this chrom
};

// one 64-byte slice of the text in registers (vector loads when the slice is whole and aligned)
struct Slice {
    uint32_t w[BD_SLICE / 4];
    uint32_t len;
    __device__ __forceinline__ char at(uint32_t i) const { return (char)((w[i >> 2] >> ((i & 3u) * 8u)) & 0xFFu); }
};
__device__ __forceinline__ Slice load_slice(const char* __restrict__ text, uint64_t a, uint64_t bytes) {
    Slice s;
    s.len = (uint32_t)min((uint64_t)BD_SLICE, bytes - a);
    if (s.len == BD_SLICE && ((uintptr_t)(text + a) & 15u) == 0) {
        const uint4* p = reinterpret_cast<const uint4*>(text + a);
#pragma unroll
        for (int k = 0; k < (int)(BD_SLICE / 16); ++k) {
            const uint4 v = __ldg(p + k);
            s.w[4 * k] = v.x; s.w[4 * k + 1] = v.y; s.w[4 * k + 2] = v.z; s.w[4 * k + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < (int)(BD_SLICE / 4); ++k) s.w[k] = 0;
        for (uint32_t i = 0; i < s.len; ++i) s.w[i >> 2] |= (uint32_t)(unsigned char)text[a + i] << ((i & 3u) * 8u);
    }
    return s;
}

// a line STARTS at byte 0 and after every '\n' that has text behind it
__global__ void __launch_bounds__(BD_THREADS)
bd_count_lines(const char* __restrict__ text, uint64_t bytes, uint64_t slices, uint32_t* __restrict__ cnt) {
    const uint64_t t = (uint64_t)blockIdx.x * BD_THREADS + threadIdx.x;
    if (t >= slices) return;
    const uint64_t a = t * BD_SLICE;
    const Slice s = load_slice(text, a, bytes);
    uint32_t c = (t == 0) ? 1u : 0u;
#pragma unroll
    for (uint32_t i = 0; i < BD_SLICE; ++i) c += (i < s.len && s.at(i) == '\n' && a + i + 1 < bytes) ? 1u : 0u;
    cnt[t] = c;
}

__global__ void __launch_bounds__(BD_THREADS)
bd_line_starts(const char* __restrict__ text, uint64_t bytes, uint64_t slices, const uint64_t* __restrict__ off,
               uint64_t* __restrict__ line_at) {
    const uint64_t t = (uint64_t)blockIdx.x * BD_THREADS + threadIdx.x;
    if (t >= slices) return;
    const uint64_t a = t * BD_SLICE;
    const Slice s = load_slice(text, a, bytes);
    uint64_t o = off[t];
    if (t == 0) line_at[o++] = 0;
#pragma unroll
    for (uint32_t i = 0; i < BD_SLICE; ++i)
        if (i < s.len && s.at(i) == '\n' && a + i + 1 < bytes) line_at[o++] = a + i + 1;
}

// std::stoi on the token [p, e): leading isspace, optional sign, at least one digit, stops at
// the first non-digit; false when there is no digit or the value leaves int32.
__device__ __forceinline__ bool parse_i32(const char* __restrict__ text, uint64_t p, uint64_t e, int32_t* out) {
    while (p < e && (text[p] == ' ' || (text[p] >= '\t' && text[p] <= '\r'))) ++p;
    bool neg = false;
    if (p < e && (text[p] == '+' || text[p] == '-')) { neg = text[p] == '-'; ++p; }
    if (p >= e || text[p] < '0' || text[p] > '9') return false;
    int64_t v = 0;
    for (; p < e && text[p] >= '0' && text[p] <= '9'; ++p) {
        v = v * 10 + (text[p] - '0');
        if (v > (int64_t)INT_MAX + 1) return false;
    }
    if (neg) v = -v;
    if (v > INT_MAX || v < INT_MIN) return false;
    *out = (int32_t)v;
    return true;
}

__global__ void __launch_bounds__(BD_THREADS)
bd_parse(const char* __restrict__ text, uint64_t bytes, const uint64_t* __restrict__ line_at, uint64_t lines,
         int normalize, int32_t end_shift, unsigned long long* __restrict__ hash, int32_t* __restrict__ starts,
         int32_t* __restrict__ ends, uint32_t* __restrict__ valid, ChromSlot* __restrict__ table,
         uint32_t* __restrict__ table_full) {
    const uint64_t l = (uint64_t)blockIdx.x * BD_THREADS + threadIdx.x;
    const bool live = l < lines;
    unsigned long long h = 1469598103934665603ull;                      // FNV-1a
    bool ok = false;
    int32_t s = 0, e = 0;
    if (live) {
        const uint64_t p = line_at[l];
        uint64_t eol = (l + 1 < lines) ? line_at[l + 1] - 1 : bytes;      // the '\n' itself, or the end of the text
        if (l + 1 == lines && eol > p && text[eol - 1] == '\n') --eol;
        // field 1: chrom
        uint64_t q = p;
        for (; q < eol && text[q] != '\t'; ++q) h = (h ^ (unsigned char)text[q]) * 1099511628211ull;
        if (h == BD_EMPTY) h = 1;
        ok = q < eol && q > p;                                           // a chrom and a tab after it
        if (ok) {
            uint64_t f2 = q + 1, f2e = f2;
            while (f2e < eol && text[f2e] != '\t') ++f2e;
            ok = parse_i32(text, f2, f2e, &s) && f2e < eol;
            if (ok) {
                uint64_t f3 = f2e + 1, f3e = f3;
                while (f3e < eol && text[f3e] != '\t') ++f3e;
                ok = parse_i32(text, f3, f3e, &e);
            }
        }
        if (ok) {
            if (normalize) { const int32_t a = min(s, e), b = max(s, e); s = a; e = b; }   // bench.cpp:89
            const int64_t e2 = (int64_t)e + end_shift;                   // half-open -> inclusive: bench.cpp:210 passes end - 1
            ok = e2 >= INT_MIN && e2 <= INT_MAX;
            e = (int32_t)e2;
        }
        valid[l] = ok ? 1u : 0u;
        hash[l] = h;
        starts[l] = s;
        ends[l] = e;
    }
    // first sight of each chrom: neighbouring lines share their chrom, so one lane per group of equal
    // hashes speaks for the warp (its line number is the group's smallest), and a slot that already
    // records an earlier line is left alone -- after the first few warps no atomic is issued at all
    const unsigned long long key = ok ? h : BD_EMPTY;
    const uint32_t peers = __match_any_sync(FULL_MASK, key);
    if (!ok || (uint32_t)(__ffs(peers) - 1) != lane_id()) return;
    uint32_t slot = (uint32_t)(h ^ (h >> 32)) & (BD_TABLE - 1);
    for (uint32_t probe = 0; probe < BD_TABLE; ++probe, slot = (slot + 1) & (BD_TABLE - 1)) {
        unsigned long long cur = *(volatile unsigned long long*)&table[slot].hash;
        if (cur == BD_EMPTY) cur = atomicCAS(&table[slot].hash, BD_EMPTY, h);
        if (cur == BD_EMPTY || cur == h) {
            if (*(volatile unsigned long long*)&table[slot].first > (unsigned long long)l)
                atomicMin(&table[slot].first, (unsigned long long)l);
            return;
        }
    }
    *table_full = 1u;
}

// ids[slot] = contig id of the chrom in that table slot (host-assigned, order of first appearance)
__global__ void __launch_bounds__(BD_THREADS)
// A contig is identified by the 64-bit hash of its name; that the NAME is the same is verified here, byte for byte against
// the name the host took from the contig's first line (names: name k at names + name_at[k], name_at[k + 1] - name_at[k]
// bytes), so that two names sharing a hash are reported (*collision) instead of silently merged.
bd_finish(const unsigned long long* __restrict__ hash, const int32_t* __restrict__ starts, const int32_t* __restrict__ ends,
          const uint32_t* __restrict__ valid, const uint64_t* __restrict__ off, uint64_t lines,
          const ChromSlot* __restrict__ table, const int32_t* __restrict__ ids, int32_t* __restrict__ out_c,
          int32_t* __restrict__ out_s, int32_t* __restrict__ out_e, const char* __restrict__ text,
          const uint64_t* __restrict__ line_at, const char* __restrict__ names, const uint32_t* __restrict__ name_at,
          uint32_t* __restrict__ collision) {
    const uint64_t l = (uint64_t)blockIdx.x * BD_THREADS + threadIdx.x;
    if (l >= lines || !valid[l]) return;
    const unsigned long long h = hash[l];
    uint32_t slot = (uint32_t)(h ^ (h >> 32)) & (BD_TABLE - 1);
    while (table[slot].hash != h) slot = (slot + 1) & (BD_TABLE - 1);
    const uint64_t o = off[l];
    const int32_t id = ids[slot];
    {
        const uint32_t a = name_at[id], len = name_at[id + 1] - a;
        const char* t = text + line_at[l];
        bool same = t[len] == '\t';                                    // a valid line has a tab after its chrom
        for (uint32_t k = 0; k < len && same; ++k) same = t[k] == names[a + k];
        if (!same) *collision = 1u;
    }
    out_c[o] = id;
    out_s[o] = starts[l];
    out_e[o] = ends[l];
}

// group == 1: records reordered by contig (stable), perm from the build's radix sort
__global__ void __launch_bounds__(BD_THREADS)
bd_group_kernel(const uint32_t* __restrict__ perm, const int32_t* __restrict__ c, const int32_t* __restrict__ s,
                const int32_t* __restrict__ e, uint64_t n, int32_t* __restrict__ gc, int32_t* __restrict__ gs,
                int32_t* __restrict__ ge) {
    const uint64_t i = (uint64_t)blockIdx.x * BD_THREADS + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = perm[i];
    gc[i] = c[j];
    gs[i] = s[j];
    ge[i] = e[j];
}

// offsets[k] = first record of contig k in the grouped columns (k = 0..contigs; offsets[contigs] = n)
__global__ void bd_contig_offsets_kernel(const int32_t* __restrict__ gc, uint64_t n, uint32_t contigs,
                                         unsigned long long* __restrict__ offsets) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > contigs) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (gc[mid] < (int32_t)k) lo = mid + 1; else hi = mid;
    }
    offsets[k] = lo;
}

struct Bufs {
    DevBuf b[22];
    ~Bufs() { for (auto& x : b) x.release(); }
};

#define BD_LAUNCH(kernel, n, stream, ...)                                                                   \
    do {                                                                                                    \
        kernel<<<(unsigned)(((uint64_t)(n) + BD_THREADS - 1) / BD_THREADS), BD_THREADS, 0, (stream)>>>(__VA_ARGS__); \
        SIB_CHECK_LAUNCH();                                                                                 \
        note_launch();                                                                                      \
    } while (0)

int scan_u32(siIndex* ix, DevBuf& cnt, uint64_t n, DevBuf& off, uint64_t* total) {
    if (off.ensure((n + 1) * 8 + 64)) return last_error_code();
    int rc = siScanDevice(ix, cnt.as<uint32_t>(), n, off.as<uint64_t>(), ix->own_stream);
    if (rc) return rc;
    SIB_CHECK(cudaMemcpyAsync(total, off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, ix->own_stream));
    SIB_CHECK(cudaStreamSynchronize(ix->own_stream));
    return 0;
}

int parse_impl(siIndex* ix, const char* text, size_t bytes, int normalize, int end_shift, int group, siBedTable* out) {
    cudaStream_t st = ix->own_stream;
    Bufs B;
    DevBuf &d_text = B.b[0], &cnt = B.b[1], &off = B.b[2], &line_at = B.b[3], &hash = B.b[4], &ds = B.b[5], &de = B.b[6],
           &valid = B.b[7], &table = B.b[8], &voff = B.b[9], &ids = B.b[10], &oc = B.b[11], &os = B.b[12], &oe = B.b[13];
    const uint64_t slices = (bytes + BD_SLICE - 1) / BD_SLICE;
    if (slices > 0xFFFFFFFFull) { set_error_msg(cudaErrorInvalidValue, "siParseBed: more than 256 GiB of text in one call"); return cudaErrorInvalidValue; }
    if (d_text.ensure(bytes + 16) || cnt.ensure(slices * 4 + 64)) return last_error_code();
    int rc = copy_h2d(ix, d_text.p, text, bytes, st);
    if (rc) return rc;
    const char* dt = d_text.as<char>();
    BD_LAUNCH(bd_count_lines, slices, st, dt, (uint64_t)bytes, slices, cnt.as<uint32_t>());
    uint64_t lines = 0;
    rc = scan_u32(ix, cnt, slices, off, &lines);
    if (rc) return rc;
    if (lines > 0xFFFFFFFFull) { set_error_msg(cudaErrorInvalidValue, "siParseBed: more than 2^32-1 lines in one call"); return cudaErrorInvalidValue; }
    if (line_at.ensure(lines * 8 + 8) || hash.ensure(lines * 8 + 8) || ds.ensure(lines * 4 + 4) || de.ensure(lines * 4 + 4) ||
        valid.ensure(lines * 4 + 64) || table.ensure((size_t)BD_TABLE * sizeof(ChromSlot)) || ix->small.ensure(256))
        return last_error_code();
    BD_LAUNCH(bd_line_starts, slices, st, dt, (uint64_t)bytes, slices, off.as<uint64_t>(), line_at.as<uint64_t>());
    // table: hash = 0 (empty), first = all ones
    SIB_CHECK(cudaMemsetAsync(table.p, 0xFF, (size_t)BD_TABLE * sizeof(ChromSlot), st));
    SIB_CHECK(cudaMemset2DAsync(table.p, sizeof(ChromSlot), 0, sizeof(unsigned long long), BD_TABLE, st));
    uint32_t* d_full = ix->small.as<uint32_t>() + 40;
    SIB_CHECK(cudaMemsetAsync(d_full, 0, 4, st));
    BD_LAUNCH(bd_parse, lines, st, dt, (uint64_t)bytes, line_at.as<uint64_t>(), lines, normalize, (int32_t)end_shift,
              hash.as<unsigned long long>(), ds.as<int32_t>(), de.as<int32_t>(), valid.as<uint32_t>(), table.as<ChromSlot>(), d_full);
    uint64_t n = 0;
    rc = scan_u32(ix, valid, lines, voff, &n);
    if (rc) return rc;
    // the distinct chroms: read the small table, order by first appearance, name them from the text
    std::vector<ChromSlot> h_table(BD_TABLE);
    uint32_t full = 0;
    SIB_CHECK(cudaMemcpyAsync(h_table.data(), table.p, (size_t)BD_TABLE * sizeof(ChromSlot), cudaMemcpyDeviceToHost, st));
    SIB_CHECK(cudaMemcpyAsync(&full, d_full, 4, cudaMemcpyDeviceToHost, st));
    SIB_CHECK(cudaStreamSynchronize(st));
    std::vector<std::pair<unsigned long long, uint32_t>> seen;   // (first line, slot)
    for (uint32_t sl = 0; sl < BD_TABLE; ++sl)
        if (h_table[sl].hash != BD_EMPTY) seen.emplace_back(h_table[sl].first, sl);
    if (full || seen.size() > BD_TABLE / 2) { set_error_msg(cudaErrorInvalidValue, "siParseBed: more than 32768 distinct contig names"); return cudaErrorInvalidValue; }
    std::sort(seen.begin(), seen.end());
    std::vector<int32_t> h_ids(BD_TABLE, -1);
    std::vector<uint64_t> first_at(seen.size());
    for (size_t k = 0; k < seen.size(); ++k) h_ids[seen[k].second] = (int32_t)k;
    // byte offsets of the first line of every chrom (a handful of 8-byte reads)
    for (size_t k = 0; k < seen.size(); ++k)
        SIB_CHECK(cudaMemcpyAsync(&first_at[k], line_at.as<uint64_t>() + seen[k].first, 8, cudaMemcpyDeviceToHost, st));
    SIB_CHECK(cudaStreamSynchronize(st));
    out->names = (char**)calloc(seen.size() ? seen.size() : 1, sizeof(char*));
    out->n_contigs = seen.size();
    for (size_t k = 0; k < seen.size(); ++k) {
        const char* p = text + first_at[k];
        size_t len = 0;
        while (first_at[k] + len < bytes && p[len] != '\t' && p[len] != '\n') ++len;
        out->names[k] = (char*)malloc(len + 1);
        memcpy(out->names[k], p, len);
        out->names[k][len] = 0;
    }
    out->n = n;
    out->lines = lines;
    out->skipped = lines - n;
    if (n == 0) return 0;
    if (ids.ensure((size_t)BD_TABLE * 4) || oc.ensure(n * 4) || os.ensure(n * 4) || oe.ensure(n * 4)) return last_error_code();
    SIB_CHECK(cudaMemcpyAsync(ids.p, h_ids.data(), (size_t)BD_TABLE * 4, cudaMemcpyHostToDevice, st));
    // the names on the device, for bd_finish's byte-for-byte check of every line's chrom against its contig's name
    std::vector<uint32_t> h_name_at(seen.size() + 1, 0);
    std::string blob;
    for (size_t k = 0; k < seen.size(); ++k) { h_name_at[k] = (uint32_t)blob.size(); blob += out->names[k]; }
    h_name_at[seen.size()] = (uint32_t)blob.size();
    DevBuf &d_names = B.b[19], &d_name_at = B.b[20];
    if (d_names.ensure(blob.size() + 16) || d_name_at.ensure(h_name_at.size() * 4)) return last_error_code();
    SIB_CHECK(cudaMemcpyAsync(d_names.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
    SIB_CHECK(cudaMemcpyAsync(d_name_at.p, h_name_at.data(), h_name_at.size() * 4, cudaMemcpyHostToDevice, st));
    uint32_t* d_collision = d_full + 1;
    SIB_CHECK(cudaMemsetAsync(d_collision, 0, 4, st));
    BD_LAUNCH(bd_finish, lines, st, hash.as<unsigned long long>(), ds.as<int32_t>(), de.as<int32_t>(), valid.as<uint32_t>(),
              voff.as<uint64_t>(), lines, table.as<ChromSlot>(), ids.as<int32_t>(), oc.as<int32_t>(), os.as<int32_t>(), oe.as<int32_t>(),
              dt, line_at.as<uint64_t>(), d_names.as<char>(), d_name_at.as<uint32_t>(), d_collision);
    uint32_t collided = 0;
    SIB_CHECK(cudaMemcpyAsync(&collided, d_collision, 4, cudaMemcpyDeviceToHost, st));
    SIB_CHECK(cudaStreamSynchronize(st));
    if (collided) { set_error_msg(cudaErrorInvalidValue, "siParseBed: two contig names share a 64-bit hash (not merged: rename one)"); return cudaErrorInvalidValue; }
    const void *fc = oc.p, *fs = os.p, *fe = oe.p;
    if (group) {
        // per-contig containers (bed-intersect-si.rs:100-123): a stable sort of the record numbers by
        // contig id, the columns gathered through it, and where each contig begins
        DevBuf &perm = B.b[14], &gc = B.b[15], &gs = B.b[16], &ge = B.b[17], &coff = B.b[18];
        if (perm.ensure(n * 4) || gc.ensure(n * 4) || gs.ensure(n * 4) || ge.ensure(n * 4) || coff.ensure((seen.size() + 1) * 8)) return last_error_code();
        int bits = 1;
        while (((size_t)1 << bits) < seen.size()) ++bits;
        rc = si_b200_stable_order_(ix, oc.as<int32_t>(), n, bits, perm.as<uint32_t>(), st);
        if (rc) return rc;
        BD_LAUNCH(bd_group_kernel, n, st, perm.as<uint32_t>(), oc.as<int32_t>(), os.as<int32_t>(), oe.as<int32_t>(), (uint64_t)n,
                  gc.as<int32_t>(), gs.as<int32_t>(), ge.as<int32_t>());
        bd_contig_offsets_kernel<<<(unsigned)((seen.size() + 1 + 127) / 128), 128, 0, st>>>(gc.as<int32_t>(), (uint64_t)n, (uint32_t)seen.size(),
                                                                                          coff.as<unsigned long long>());
        SIB_CHECK_LAUNCH();
        note_launch();
        out->contig_offsets = (size_t*)malloc((seen.size() + 1) * sizeof(size_t));
        if (!out->contig_offsets) { set_error_msg(cudaErrorMemoryAllocation, "siParseBed: out of host memory"); return cudaErrorMemoryAllocation; }
        static_assert(sizeof(size_t) == 8, "LP64 only");
        SIB_CHECK(cudaMemcpyAsync(out->contig_offsets, coff.p, (seen.size() + 1) * 8, cudaMemcpyDeviceToHost, st));
        fc = gc.p; fs = gs.p; fe = ge.p;
    }
    out->contig = (int32_t*)malloc(n * 4);
    out->starts = (int32_t*)malloc(n * 4);
    out->ends = (int32_t*)malloc(n * 4);
    if (!out->contig || !out->starts || !out->ends) { set_error_msg(cudaErrorMemoryAllocation, "siParseBed: out of host memory"); return cudaErrorMemoryAllocation; }
    if (copy_d2h(ix, out->contig, fc, n * 4, st) || copy_d2h(ix, out->starts, fs, n * 4, st) || copy_d2h(ix, out->ends, fe, n * 4, st))
        return last_error_code();
    return 0;
}

}  // namespace

extern "C" {

int siParseBed(const char* text, size_t bytes, int normalize, int end_shift, int group_by_contig, siBedTable* out) {
    if (!out) return cudaErrorInvalidValue;
    memset(out, 0, sizeof(*out));
    if (!text || bytes == 0) return 0;
    siIndex* ix = siIndexCreate();
    if (!ix) return last_error_code();
    const int rc = parse_impl(ix, text, bytes, normalize, end_shift, group_by_contig, out);
    siIndexDestroy(ix);
    if (rc) siBedTableFree(out);
    return rc;
}

void siBedTableFree(siBedTable* t) {
    if (!t) return;
    free(t->contig);
    free(t->starts);
    free(t->ends);
    free(t->contig_offsets);
    if (t->names) {
        for (size_t k = 0; k < t->n_contigs; ++k) free(t->names[k]);
        free(t->names);
    }
    memset(t, 0, sizeof(*t));
}

}  // extern "C"
