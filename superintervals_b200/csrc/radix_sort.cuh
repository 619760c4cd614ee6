// radix_sort.cuh -- hand-written stable LSD radix sort of (key, uint32 payload) pairs
// for sm_100a: 8-bit digits, ONE histogram pass over the keys for all digits, then
// one single-pass "onesweep" kernel per live digit (tile-local ranking with
// __match_any_sync + chained decoupled look-back across tiles for the global
// digit offsets). Digits that are constant over the whole input (genomic
// coordinates < 2^28 leave the top digit of each 32-bit half constant) are
// detected on the device from the histogram and their passes exit immediately:
// no host synchronisation anywhere, so the sort is stream-ordered / graph-safe.
//
// Replaces std::sort in the reference's build() (superintervals.hpp:1396-1420)
// and, for queries, gives the position-sorted query tiles the count kernel wants.
// Traffic per live pass: read key+payload, write key+payload (24 B/pair for
// 64-bit keys, 16 B/pair for 32-bit) + 2 KB of look-back state per tile.
#pragma once

#include "common.cuh"
#include "index.cuh"

namespace sib {

constexpr int RS_RADIX_BITS = 8;
constexpr int RS_RADIX = 1 << RS_RADIX_BITS;
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAX_PASSES = 8;

// look-back status word: [63:62] flag, [61:0] count
constexpr unsigned long long RS_FLAG_AGG = 1ull << 62;   // tile aggregate available
constexpr unsigned long long RS_FLAG_INC = 2ull << 62;   // inclusive prefix available
constexpr unsigned long long RS_VAL_MASK = (1ull << 62) - 1;

struct RsPassInfo {
    uint32_t trivial;   // 1 -> every key has the same digit in this pass: skip
    uint32_t src_sel;   // 0 -> read buffer A / write B; 1 -> read B / write A
};

template <typename KeyT> struct RsTraits;
// 32-bit keys: 12 items per thread at 4 CTAs/SM (64 registers, no spills) measured best on the 10 M-interval build
// (tools/gpu_r02l.sh: 1.95 ms against 2.07-2.20 ms for 18 items at 2 CTAs/SM); the pass is latency-bound, not bandwidth-bound
#ifndef SIB_RS_ITEMS32
#define SIB_RS_ITEMS32 12
#endif
#ifndef SIB_RS_MINBLOCKS32
#define SIB_RS_MINBLOCKS32 4
#endif
#ifndef SIB_RS_ITEMS64
#define SIB_RS_ITEMS64 14
#endif
#ifndef SIB_RS_MINBLOCKS
#define SIB_RS_MINBLOCKS 1
#endif
template <> struct RsTraits<uint32_t> { static constexpr int ITEMS = SIB_RS_ITEMS32; static constexpr int MINB = SIB_RS_MINBLOCKS32; };
template <> struct RsTraits<uint64_t> { static constexpr int ITEMS = SIB_RS_ITEMS64; static constexpr int MINB = SIB_RS_MINBLOCKS; };

struct RsWorkspace {
    uint32_t* hist;        // [RS_MAX_PASSES][256] counts, turned into exclusive scans in place
    RsPassInfo* info;      // [RS_MAX_PASSES]
    uint32_t* final_sel;   // which buffer holds the sorted output (0 = A, 1 = B)
    uint32_t* tickets;     // [RS_MAX_PASSES] dynamic tile ids
    unsigned long long* status;   // [tiles][256] look-back words (re-zeroed per pass)
};

// Lanes of the warp holding the same 8-bit digit. MATCH.ANY is one instruction but a slow one (r02k: half of the
// onesweep samples wait on it); eight ballots give the same mask from instructions that pipeline (SIB_RS_MATCH=1
// restores the match instruction).
#ifndef SIB_RS_MATCH
#define SIB_RS_MATCH 0
#endif
__device__ __forceinline__ uint32_t rs_same_digit(uint32_t d) {
#if SIB_RS_MATCH
    return __match_any_sync(FULL_MASK, d);
#else
    uint32_t m = FULL_MASK;
#pragma unroll
    for (int b = 0; b < RS_RADIX_BITS; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t v = __ballot_sync(FULL_MASK, bit);
        m &= bit ? v : ~v;
    }
    return m;
#endif
}

template <typename KeyT>
__host__ inline uint32_t rs_num_tiles(uint32_t n) {
    constexpr uint32_t TILE = RS_THREADS * RsTraits<KeyT>::ITEMS;
    return (uint32_t)(((uint64_t)n + TILE - 1) / TILE);
}

template <typename KeyT>
__host__ inline size_t rs_workspace_bytes(uint32_t n) {
    size_t b = 0;
    b += sizeof(uint32_t) * RS_MAX_PASSES * RS_RADIX;
    b += sizeof(RsPassInfo) * RS_MAX_PASSES;
    b += 256;   // final_sel + tickets, padded
    b += sizeof(unsigned long long) * (size_t)rs_num_tiles<KeyT>(n) * RS_RADIX;
    return b + 1024;
}

__host__ inline RsWorkspace rs_carve(void* base) {
    RsWorkspace w;
    char* p = (char*)base;
    w.hist = (uint32_t*)p;            p += sizeof(uint32_t) * RS_MAX_PASSES * RS_RADIX;
    w.info = (RsPassInfo*)p;          p += sizeof(RsPassInfo) * RS_MAX_PASSES;
    w.final_sel = (uint32_t*)p;       p += 64;
    w.tickets = (uint32_t*)p;         p += 192;
    p = (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
    w.status = (unsigned long long*)p;
    return w;
}

// ---- 1. one pass over the keys: histograms of every digit -------------------------
template <typename KeyT, int NPASS>
__global__ void __launch_bounds__(RS_THREADS)
rs_histogram_kernel(const KeyT* __restrict__ keys, uint32_t n, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[NPASS * RS_RADIX];
    for (int i = threadIdx.x; i < NPASS * RS_RADIX; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * RS_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += stride) {
        KeyT k = keys[i];
#pragma unroll
        for (int p = 0; p < NPASS; ++p)
            atomicAdd(&sh[p * RS_RADIX + (uint32_t)((k >> (p * RS_RADIX_BITS)) & (RS_RADIX - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NPASS * RS_RADIX; i += RS_THREADS) {
        uint32_t v = sh[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// ---- 2. one CTA: exclusive scan per digit position, trivial-pass detection ---------
__global__ void __launch_bounds__(RS_RADIX)
rs_prepare_kernel(uint32_t* __restrict__ hist, RsPassInfo* __restrict__ info,
                  uint32_t* __restrict__ final_sel, uint32_t n, int npass) {
    __shared__ uint32_t s_scan[RS_RADIX];
    __shared__ uint32_t s_trivial[RS_MAX_PASSES];
    const int d = threadIdx.x;
    if (d < RS_MAX_PASSES) s_trivial[d] = 0;
    __syncthreads();
    for (int p = 0; p < npass; ++p) {
        uint32_t c = hist[p * RS_RADIX + d];
        if (c == n) s_trivial[p] = 1;
        s_scan[d] = c;
        __syncthreads();
        // Hillis-Steele inclusive scan over 256 bins (tiny, one CTA)
        for (int off = 1; off < RS_RADIX; off <<= 1) {
            uint32_t add = d >= off ? s_scan[d - off] : 0;
            __syncthreads();
            s_scan[d] += add;
            __syncthreads();
        }
        hist[p * RS_RADIX + d] = s_scan[d] - c;   // exclusive
        __syncthreads();
    }
    if (d == 0) {
        uint32_t sel = 0;
        for (int p = 0; p < npass; ++p) {
            info[p].trivial = s_trivial[p];
            info[p].src_sel = sel;
            if (!s_trivial[p]) sel ^= 1u;
        }
        *final_sel = sel;
    }
}

// ---- 3. onesweep pass --------------------------------------------------------------
template <typename KeyT, bool VALS>
__global__ void __launch_bounds__(RS_THREADS, RsTraits<KeyT>::MINB)
rs_onesweep_kernel(KeyT* __restrict__ kA, KeyT* __restrict__ kB,
                   uint32_t* __restrict__ vA, uint32_t* __restrict__ vB, uint32_t n, int pass,
                   const uint32_t* __restrict__ gbase_all, const RsPassInfo* __restrict__ info,
                   unsigned long long* __restrict__ status, uint32_t* __restrict__ ticket) {
    constexpr int ITEMS = RsTraits<KeyT>::ITEMS;
    constexpr uint32_t TILE = RS_THREADS * ITEMS;

    const RsPassInfo pi = info[pass];
    if (pi.trivial) return;
    const KeyT* __restrict__ kin = pi.src_sel ? kB : kA;
    KeyT* __restrict__ kout = pi.src_sel ? kA : kB;
    const uint32_t* __restrict__ vin = pi.src_sel ? vB : vA;
    uint32_t* __restrict__ vout = pi.src_sel ? vA : vB;
    const uint32_t* __restrict__ gbase = gbase_all + pass * RS_RADIX;
    const int shift = pass * RS_RADIX_BITS;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    KeyT* s_keys = reinterpret_cast<KeyT*>(smem_raw);                               // [TILE]
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(smem_raw + sizeof(KeyT) * TILE);  // [TILE]
    uint32_t* s_whist = s_vals + TILE;                                              // [WARPS][256]
    uint32_t* s_dstart = s_whist + RS_WARPS * RS_RADIX;                             // [256]
    uint32_t* s_gofs = s_dstart + RS_RADIX;                                         // [256]
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_wsum[RS_WARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) s_whist[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_base = (uint64_t)tile * TILE;
    const uint64_t wbase = tile_base + (uint64_t)warp * 32u * ITEMS;

    // warp-striped load: item k of lane l is element wbase + k*32 + l (coalesced, order = (k, l))
    KeyT key[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        uint64_t idx = wbase + (uint64_t)k * 32u + lane;
        key[k] = idx < n ? kin[idx] : ~(KeyT)0;   // padding sorts to the very end of the tile
    }

    // stable rank of every key inside its warp's digit bucket
    uint32_t rank[ITEMS];
    uint32_t* wh = s_whist + warp * RS_RADIX;
    const uint32_t lt = lanemask_lt();
    uint32_t same[ITEMS];   // the masks first: they do not depend on the histogram chain below
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) same[k] = rs_same_digit((uint32_t)((key[k] >> shift) & (RS_RADIX - 1)));
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        uint32_t d = (uint32_t)((key[k] >> shift) & (RS_RADIX - 1));
        uint32_t peers = same[k];
        uint32_t leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (lane == leader) {
            pre = wh[d];
            wh[d] = pre + __popc(peers);
        }
        pre = __shfl_sync(FULL_MASK, pre, leader);
        rank[k] = pre + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // thread d owns digit d: exclusive prefix over warps, tile total
    uint32_t total = 0;
    {
        const uint32_t d = tid;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t t = s_whist[w * RS_RADIX + d];
            s_whist[w * RS_RADIX + d] = total;
            total += t;
        }
    }
    // publish the tile aggregate early so successors can make progress
    unsigned long long* my_status = status + (uint64_t)tile * RS_RADIX + tid;
    if (tile == 0) {
        *(volatile unsigned long long*)my_status = RS_FLAG_INC | (unsigned long long)total;
    } else {
        *(volatile unsigned long long*)my_status = RS_FLAG_AGG | (unsigned long long)total;
    }

    // block-wide exclusive scan of `total` over digits -> s_dstart
    {
        uint32_t incl = total;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_up_sync(FULL_MASK, incl, off);
            if (lane >= (uint32_t)off) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t wpre = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) wpre += (w < (int)warp) ? s_wsum[w] : 0;
        s_dstart[tid] = wpre + incl - total;
    }

    // chained look-back: exclusive count of digit `tid` over all earlier tiles
    uint64_t excl = 0;
    if (tile > 0) {
        uint32_t t = tile - 1;
        while (true) {
            unsigned long long v = *(volatile unsigned long long*)(status + (uint64_t)t * RS_RADIX + tid);
            unsigned long long f = v & ~RS_VAL_MASK;
            if (f == 0) continue;   // predecessor not published yet (it holds an earlier ticket: it is running)
            excl += v & RS_VAL_MASK;
            if (f == RS_FLAG_INC) break;
            --t;
        }
        *(volatile unsigned long long*)my_status = RS_FLAG_INC | (excl + total);
    }
    s_gofs[tid] = gbase[tid] + (uint32_t)excl - s_dstart[tid];
    __syncthreads();

    // scatter into shared memory in digit order (keys), remember the slot for the payload
    uint32_t slot[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        uint32_t d = (uint32_t)((key[k] >> shift) & (RS_RADIX - 1));
        slot[k] = s_dstart[d] + wh[d] + rank[k];
        s_keys[slot[k]] = key[k];
    }
    if (VALS) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            uint64_t idx = wbase + (uint64_t)k * 32u + lane;
            s_vals[slot[k]] = idx < n ? vin[idx] : 0u;
        }
    }
    __syncthreads();

    // coalesced write-out: consecutive threads write consecutive addresses inside a digit run
    const uint32_t valid = (uint32_t)(((uint64_t)n - tile_base) < TILE ? ((uint64_t)n - tile_base) : TILE);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        uint32_t li = k * RS_THREADS + tid;
        if (li < valid) {
            KeyT kk = s_keys[li];
            uint32_t d = (uint32_t)((kk >> shift) & (RS_RADIX - 1));
            uint32_t dst = s_gofs[d] + li;
            kout[dst] = kk;
            if (VALS) vout[dst] = s_vals[li];
        }
    }
}

template <typename KeyT>
__host__ inline size_t rs_onesweep_smem_bytes() {
    constexpr size_t TILE = RS_THREADS * RsTraits<KeyT>::ITEMS;
    return sizeof(KeyT) * TILE + sizeof(uint32_t) * TILE + sizeof(uint32_t) * (RS_WARPS * RS_RADIX + 2 * RS_RADIX);
}

// Stable sort of n pairs by the low `key_bits` bits of the key. Input in (kA, vA); (kB, vB)
// are same-sized alternates. On completion *ws.final_sel (device) says which pair of
// buffers holds the result. Returns a cudaError_t as int.
template <typename KeyT>
__host__ inline int radix_sort_pairs(KeyT* kA, KeyT* kB, uint32_t* vA, uint32_t* vB, uint32_t n,
                                     int key_bits, void* workspace, int sm_count, cudaStream_t s) {
    RsWorkspace ws = rs_carve(workspace);
    const int npass = (key_bits + RS_RADIX_BITS - 1) / RS_RADIX_BITS;
    const uint32_t tiles = rs_num_tiles<KeyT>(n);
    SIB_CHECK(cudaMemsetAsync(ws.hist, 0, sizeof(uint32_t) * RS_MAX_PASSES * RS_RADIX, s));
    SIB_CHECK(cudaMemsetAsync(ws.tickets, 0, sizeof(uint32_t) * RS_MAX_PASSES, s));
    if (n == 0) {
        SIB_CHECK(cudaMemsetAsync(ws.final_sel, 0, sizeof(uint32_t), s));
        return 0;
    }
    {
        int grid = sm_count * 8;
        uint32_t need = ceil_div_u32(n, RS_THREADS);
        if ((uint32_t)grid > need) grid = (int)need;
        if (sizeof(KeyT) == 8 && npass > 4)
            rs_histogram_kernel<KeyT, 8><<<grid, RS_THREADS, 0, s>>>(kA, n, ws.hist);
        else
            rs_histogram_kernel<KeyT, 4><<<grid, RS_THREADS, 0, s>>>(kA, n, ws.hist);
        SIB_CHECK_LAUNCH();
        note_launch();
    }
    const int hist_passes = (sizeof(KeyT) == 8 && npass > 4) ? 8 : 4;
    const int run_passes = npass < hist_passes ? npass : hist_passes;
    rs_prepare_kernel<<<1, RS_RADIX, 0, s>>>(ws.hist, ws.info, ws.final_sel, n, run_passes);
    SIB_CHECK_LAUNCH();
    note_launch();

    // vA == nullptr: keys only (no payload read, staged or written: half the bytes of a 32-bit pass)
    const size_t smem = rs_onesweep_smem_bytes<KeyT>();
    auto kern = vA ? rs_onesweep_kernel<KeyT, true> : rs_onesweep_kernel<KeyT, false>;
    SIB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int p = 0; p < run_passes; ++p) {
        SIB_CHECK(cudaMemsetAsync(ws.status, 0, sizeof(unsigned long long) * (size_t)tiles * RS_RADIX, s));
        kern<<<tiles, RS_THREADS, smem, s>>>(kA, kB, vA, vB, n, p, ws.hist, ws.info, ws.status, ws.tickets + p);
        SIB_CHECK_LAUNCH();
        note_launch();
    }
    return 0;
}

}  // namespace sib
