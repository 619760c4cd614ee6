// partition.cuh -- locality partition of a query batch (sm_100a, hand-written).
//
// The query kernels only need LOCALITY, not a total order: queries that sit in the
// same thread block should touch the same window of the index, and their results
// should land in the same L2-sized window of the caller's output array. So instead
// of sorting 32-bit keys completely (4 passes of 8 bits + a gather of `qe` and a
// random scatter of the results through the permutation, 19.6 GB of DRAM traffic per
// 100 M queries, profiles/ncu_summary_r01b.txt) a batch is PARTITIONED by a short key
//
//     key(q) = ( q.index >> wshift ) << P  |  min((q.start - lo) >> shift, 2^P - 1)
//              \__ result window (W bits) _/   \__ position bucket (P bits) ______/
//
// in ceil((W+P)/10) stable LSD passes (2 for the bench workload) that carry the whole
// 12-byte record (start, end, index), so the consumer streams its queries and writes
// its results inside one window of the output at a time (L2 merges the 4-byte stores).
//
// Each pass is a single-sweep ("onesweep") kernel: one histogram read of the starts up
// front gives every pass its global digit offsets; a tile (16384 records, one CTA of 1024 threads) hands out slots
// with shared-memory atomics, publishes its per-digit counts and resolves its global
// offsets by a chained decoupled look-back over 32-bit status words; records are staged
// in shared memory in digit order so that global stores are coalesced runs. Tiles take tickets
// (atomic counter) so every predecessor a tile waits for is already running.
// Traffic per pass: 12 B read + 12 B written per query (first pass reads 8 B) plus
// 4*2^bits B of look-back state per 16384-record tile; HBM-bound.
#pragma once

#include "common.cuh"
#include "index.cuh"

namespace sib {

#ifndef SIB_PT_THREADS
#define SIB_PT_THREADS 1024
#endif
constexpr int PT_THREADS = SIB_PT_THREADS;
constexpr int PT_WARPS = PT_THREADS / 32;
#ifndef SIB_PT_ITEMS
#define SIB_PT_ITEMS 16
#endif
#ifndef SIB_PT_LOOK
#define SIB_PT_LOOK 8
#endif
#ifndef SIB_PT_MINBLOCKS
#define SIB_PT_MINBLOCKS 1
#endif
constexpr int PT_ITEMS = SIB_PT_ITEMS;
constexpr uint32_t PT_TILE = PT_THREADS * PT_ITEMS;   // records per tile
constexpr int PT_MAX_PASSES = 3;
constexpr int PT_MAX_BITS = 10;
constexpr uint32_t PT_MAX_BATCH = 1u << 27;           // records per partition call (30-bit look-back counts)

constexpr uint32_t PT_FLAG_AGG = 1u << 30;   // tile aggregate available
constexpr uint32_t PT_FLAG_INC = 2u << 30;   // inclusive prefix available
constexpr uint32_t PT_VAL_MASK = (1u << 30) - 1;

struct PtKey {
    int32_t lo;        // smallest start of the index
    uint32_t shift;    // bucket = min((qs - lo) >> shift, bmax)
    uint32_t bmax;     // 2^P - 1
    uint32_t pbits;    // P
    uint32_t wshift;   // result window = index >> wshift
};

__device__ __forceinline__ uint32_t pt_key(const PtKey& k, int32_t qs, uint32_t idx) {
    const uint32_t d = qs > k.lo ? (uint32_t)qs - (uint32_t)k.lo : 0u;
    return ((idx >> k.wshift) << k.pbits) | min(d >> k.shift, k.bmax);
}

// A batch of query records in device memory. idx == nullptr means "record i is query i".
struct QueryRecords {
    const int32_t* qs;
    const int32_t* qe;
    const uint32_t* idx;
};

// ---- 1. one read of the starts: digit histograms of every pass ----------------------
__global__ void __launch_bounds__(PT_THREADS)
pt_histogram_kernel(const int32_t* __restrict__ qs, uint32_t n, PtKey key, int npass, int bits,
                    uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t pt_sh[];
    const int total = npass << bits;
    for (int i = threadIdx.x; i < total; i += PT_THREADS) pt_sh[i] = 0;
    __syncthreads();
    const uint32_t mask = (1u << bits) - 1u;
    const uint64_t stride = (uint64_t)gridDim.x * PT_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * PT_THREADS + threadIdx.x; i < n; i += stride) {
        const uint32_t k = pt_key(key, ld_stream(qs + i), (uint32_t)i);
        for (int p = 0; p < npass; ++p) atomicAdd(&pt_sh[(p << bits) + ((k >> (p * bits)) & mask)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += PT_THREADS) {
        const uint32_t v = pt_sh[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// ---- 2. one CTA: exclusive scan of each pass's digit counts ---------------------------
__global__ void __launch_bounds__(1024)
pt_prepare_kernel(uint32_t* __restrict__ hist, int npass, int bits) {
    __shared__ uint32_t s_w[32];
    const uint32_t d = threadIdx.x, lane = d & 31u, warp = d >> 5;
    const uint32_t nb = 1u << bits;
    for (int p = 0; p < npass; ++p) {
        const uint32_t c = d < nb ? hist[(p << bits) + d] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL_MASK, incl, off);
            if (lane >= (uint32_t)off) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_w[lane], wi = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL_MASK, wi, off);
                if (lane >= (uint32_t)off) wi += t;
            }
            s_w[lane] = wi - w;   // exclusive over warps
        }
        __syncthreads();
        if (d < nb) hist[(p << bits) + d] = s_w[warp] + incl - c;
        __syncthreads();
    }
}

// ---- 3. one partition pass ---------------------------------------------------------------
struct PtPass {
    const int32_t* in_qs;
    const int32_t* in_qe;
    const uint32_t* in_idx;   // nullptr on the first pass: record i is query i
    int32_t* out_qs;
    int32_t* out_qe;
    uint32_t* out_idx;
    const uint32_t* gbase;    // [2^BITS] exclusive digit offsets of this pass
    uint32_t* status;         // [tiles][2^BITS] look-back words, zeroed before the launch
    uint32_t* ticket;
    uint32_t n;
    uint32_t shift;           // digit = (key >> shift) & dmask
    uint32_t dmask;           // 2^bits - 1 with bits <= BITS
    PtKey key;
};

constexpr int PT_LOOK = SIB_PT_LOOK;   // predecessors inspected per look-back round

// volatile (L2) load of the N consecutive status words one thread owns
template <int N>
__device__ __forceinline__ void pt_ld_status(const uint32_t* p, uint32_t (&w)[N]) {
    if constexpr (N == 4) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
    } else if constexpr (N == 2) {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(p));
    } else {
        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(w[0]) : "l"(p));
    }
}

template <int BITS>
__host__ __device__ constexpr size_t pt_onesweep_smem_bytes() {
    // record staging (3 arrays) + tile histogram + dstart + gofs
    return sizeof(uint32_t) * (3 * (size_t)PT_TILE + 3 * ((size_t)1 << BITS));
}

// The ranking inside a tile is NOT stable (shared-memory atomics hand out the slots of a
// digit): a pass may permute the records one tile contributes to one digit among
// themselves. That is all the consumers need -- after an LSD pass a tile holds one or two
// adjacent values of the previous digit, so the final order is the key order up to swaps
// between neighbouring buckets inside runs of a few records; results never depend on it.
template <int BITS>
__global__ void __launch_bounds__(PT_THREADS, SIB_PT_MINBLOCKS)
pt_onesweep_kernel(PtPass a) {
    constexpr uint32_t NB = 1u << BITS;
    constexpr int DPT = NB >= PT_THREADS ? NB / PT_THREADS : 1;   // digits owned per thread (blocked)

    extern __shared__ __align__(16) unsigned char pt_smem[];
    int32_t* s_qs = reinterpret_cast<int32_t*>(pt_smem);             // [TILE] staging, digit order
    int32_t* s_qe = s_qs + PT_TILE;                                  // [TILE]
    uint32_t* s_idx = reinterpret_cast<uint32_t*>(s_qe + PT_TILE);   // [TILE]
    uint32_t* s_hist = s_idx + PT_TILE;                              // [NB] records per digit in this tile
    uint32_t* s_dstart = s_hist + NB;                                // [NB] exclusive scan of s_hist
    uint32_t* s_gofs = s_dstart + NB;                                // [NB] global offset of the digit's run - dstart
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_wsum[PT_WARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool owner = tid * DPT < NB;   // threads beyond the digit count own none
    if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
    if (owner) {
#pragma unroll
        for (int j = 0; j < DPT; ++j) s_hist[tid * DPT + j] = 0;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_base = (uint64_t)tile * PT_TILE;
    const uint64_t wbase = tile_base + (uint64_t)warp * 32u * PT_ITEMS;
    const bool first = a.in_idx == nullptr;

    // warp-striped load: item k of lane l is record wbase + 32k + l (coalesced)
    int32_t qs[PT_ITEMS];
    uint32_t qi[PT_ITEMS];
#pragma unroll
    for (int k = 0; k < PT_ITEMS; ++k) {
        const uint64_t e = wbase + (uint64_t)k * 32u + lane;
        qs[k] = e < a.n ? ld_stream(a.in_qs + e) : 0;
        qi[k] = e < a.n ? (first ? (uint32_t)e : ld_stream(a.in_idx + e)) : 0u;
    }
    // slot of every record inside its digit: dr = digit << 16 | rank (rank < TILE <= 2^16)
    uint32_t dr[PT_ITEMS];
#pragma unroll
    for (int k = 0; k < PT_ITEMS; ++k) {
        const uint64_t e = wbase + (uint64_t)k * 32u + lane;
        dr[k] = 0xFFFFFFFFu;
        if (e < a.n) {
            const uint32_t d = (pt_key(a.key, qs[k], qi[k]) >> a.shift) & a.dmask;
            dr[k] = (d << 16) | atomicAdd(&s_hist[d], 1u);
        }
    }
    __syncthreads();

    // digits [tid*DPT, tid*DPT+DPT): publish the tile's counts, then scan them over the block
    uint32_t total[DPT];
    uint32_t tsum = 0;
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        total[j] = owner ? s_hist[tid * DPT + j] : 0u;
        tsum += total[j];
    }
    volatile uint32_t* my_status = a.status + (uint64_t)tile * NB + tid * DPT;
    if (owner) {
#pragma unroll
        for (int j = 0; j < DPT; ++j) my_status[j] = (tile == 0 ? PT_FLAG_INC : PT_FLAG_AGG) | total[j];
    }
    {
        uint32_t incl = tsum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL_MASK, incl, off);
            if (lane >= (uint32_t)off) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t run = incl - tsum;
#pragma unroll
        for (int w = 0; w < PT_WARPS; ++w) run += (w < (int)warp) ? s_wsum[w] : 0u;
        if (owner) {
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                s_dstart[tid * DPT + j] = run;
                run += total[j];
            }
        }
    }
    __syncthreads();

    // stage the records in digit order; the ends are requested now and land after the look-back
#pragma unroll
    for (int k = 0; k < PT_ITEMS; ++k) {
        if (dr[k] != 0xFFFFFFFFu) {
            dr[k] = s_dstart[dr[k] >> 16] + (dr[k] & 0xFFFFu);
            s_qs[dr[k]] = qs[k];
            s_idx[dr[k]] = qi[k];
        }
    }
    int32_t qe[PT_ITEMS];
#pragma unroll
    for (int k = 0; k < PT_ITEMS; ++k) {
        const uint64_t e = wbase + (uint64_t)k * 32u + lane;
        qe[k] = e < a.n ? ld_stream(a.in_qe + e) : 0;
    }

    // decoupled look-back, PT_LOOK predecessors per round: records with my digits in all earlier
    // tiles. The status words of the window are fetched together (independent vector loads, one
    // L2 round trip) and then consumed nearest-first per digit, stopping at the first tile that
    // already holds an inclusive prefix. An unpublished predecessor (it holds an earlier ticket,
    // so it is running) is re-polled after a short sleep.
    uint32_t excl[DPT];
#pragma unroll
    for (int j = 0; j < DPT; ++j) excl[j] = 0;
    if (tile > 0 && owner) {
        uint32_t look[DPT];
        uint32_t pending = (1u << DPT) - 1u;
#pragma unroll
        for (int j = 0; j < DPT; ++j) look[j] = tile - 1;
        while (pending) {
            uint32_t base = 0;
#pragma unroll
            for (int j = 0; j < DPT; ++j)
                if (pending & (1u << j)) base = max(base, look[j]);
            uint32_t w[PT_LOOK][DPT];
#pragma unroll
            for (int p = 0; p < PT_LOOK; ++p) {
                if (base >= (uint32_t)p) {
                    pt_ld_status<DPT>(a.status + (uint64_t)(base - p) * NB + tid * DPT, w[p]);
                } else {
#pragma unroll
                    for (int j = 0; j < DPT; ++j) w[p][j] = 0;
                }
            }
            bool waited = false;
#pragma unroll
            for (int p = 0; p < PT_LOOK; ++p) {
#pragma unroll
                for (int j = 0; j < DPT; ++j) {
                    if ((pending & (1u << j)) && look[j] == base - (uint32_t)p && base >= (uint32_t)p) {
                        const uint32_t v = w[p][j];
                        const uint32_t f = v & ~PT_VAL_MASK;
                        if (f) {
                            excl[j] += v & PT_VAL_MASK;
                            if (f == PT_FLAG_INC) pending &= ~(1u << j);
                            else --look[j];
                        } else {
                            waited = true;
                        }
                    }
                }
            }
            if (waited) __nanosleep(64);
        }
#pragma unroll
        for (int j = 0; j < DPT; ++j) my_status[j] = PT_FLAG_INC | (excl[j] + total[j]);
    }
    if (owner) {
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const uint32_t d = tid * DPT + j;
            s_gofs[d] = a.gbase[d] + excl[j] - s_dstart[d];
        }
    }
#pragma unroll
    for (int k = 0; k < PT_ITEMS; ++k)
        if (dr[k] != 0xFFFFFFFFu) s_qe[dr[k]] = qe[k];
    __syncthreads();

    // coalesced write-out: consecutive threads write consecutive addresses inside a digit run
    const uint32_t valid = (uint32_t)(((uint64_t)a.n - tile_base) < PT_TILE ? ((uint64_t)a.n - tile_base) : PT_TILE);
#pragma unroll 8
    for (int k = 0; k < PT_ITEMS; ++k) {
        const uint32_t li = k * PT_THREADS + tid;
        if (li < valid) {
            const int32_t q = s_qs[li];
            const uint32_t x = s_idx[li];
            const uint32_t d = (pt_key(a.key, q, x) >> a.shift) & a.dmask;
            const uint32_t dst = s_gofs[d] + li;
            a.out_qs[dst] = q;
            a.out_qe[dst] = s_qe[li];
            a.out_idx[dst] = x;
        }
    }
}

// ---- host side ------------------------------------------------------------------------------
struct PtPlan {
    int passes = 0;   // 0: nothing to do (records stay as given)
    int bits = 0;     // digit width of every pass
    PtKey key{};
};

// Locality key for a batch of nq queries against an index of n intervals spanning [lo, hi].
// bucket_intervals: target number of index intervals per position bucket;
// window_shift: log2 of the result window in queries.
__host__ inline PtPlan pt_make_plan(uint32_t nq, uint32_t n, int32_t lo, int32_t hi, uint32_t bucket_intervals,
                                    uint32_t window_shift) {
    PtPlan p;
    const uint64_t range = hi > lo ? (uint64_t)((int64_t)hi - (int64_t)lo) : 0;
    int rbits = 0;
    while (rbits < 32 && (range >> rbits) != 0) ++rbits;   // bits needed for (qs - lo) inside the index span
    if (rbits == 0) rbits = 1;
    int pb = 0;
    while (pb < rbits && ((uint64_t)bucket_intervals << pb) < n) ++pb;   // 2^pb buckets of <= bucket_intervals
    int wb = 0;
    const uint64_t windows = ((uint64_t)nq + ((uint64_t)1 << window_shift) - 1) >> window_shift;
    while (((uint64_t)1 << wb) < windows) ++wb;
    int total = pb + wb;
    if (total == 0) return p;
    const int max_total = PT_MAX_PASSES * PT_MAX_BITS;
    if (total > max_total) { pb -= total - max_total; total = max_total; }
    p.passes = (total + PT_MAX_BITS - 1) / PT_MAX_BITS;
    p.bits = (total + p.passes - 1) / p.passes;
    if (p.bits < 8) p.bits = 8;                       // smallest instantiation; spare bits refine the buckets
    int spare = p.passes * p.bits - total;
    while (spare > 0 && pb < rbits) { ++pb; --spare; }
    p.key.lo = lo;
    p.key.pbits = (uint32_t)pb;
    p.key.bmax = pb ? (((uint32_t)1 << pb) - 1u) : 0u;
    p.key.shift = pb ? (uint32_t)(rbits - pb) : 31u;
    p.key.wshift = window_shift;
    return p;
}

__host__ inline uint32_t pt_num_tiles(uint32_t n) { return (uint32_t)(((uint64_t)n + PT_TILE - 1) / PT_TILE); }

// workspace: hist [MAX_PASSES][2^MAX_BITS] | tickets [MAX_PASSES] (padded) | status [tiles][2^bits]
__host__ inline size_t pt_workspace_bytes(uint32_t n) {
    return sizeof(uint32_t) * PT_MAX_PASSES * ((size_t)1 << PT_MAX_BITS) + 256 +
           sizeof(uint32_t) * (size_t)pt_num_tiles(n) * ((size_t)1 << PT_MAX_BITS) + 256;
}

template <int BITS>
__host__ inline int pt_launch_pass(const PtPass& a, uint32_t tiles, cudaStream_t s, LaunchTimer* timer) {
    constexpr size_t smem = pt_onesweep_smem_bytes<BITS>();
    SIB_CHECK(cudaFuncSetAttribute(pt_onesweep_kernel<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (timer) timer->begin(TAG_PT_PASS, s);
    pt_onesweep_kernel<BITS><<<tiles, PT_THREADS, smem, s>>>(a);
    SIB_CHECK_LAUNCH();
    if (timer) timer->end(s);
    note_launch();
    return 0;
}

// Partition n query records (qs, qe; record i is query i) by plan.key. bufA / bufB are two
// record buffers of 3 arrays of `cap` entries each (qs | qe | idx). On return *out describes
// the partitioned records (inside bufA or bufB). Stream-ordered, no host synchronisation.
__host__ inline int pt_partition(const PtPlan& plan, const int32_t* d_qs, const int32_t* d_qe, uint32_t n,
                                 void* bufA, void* bufB, size_t cap, void* workspace, int sm_count,
                                 cudaStream_t s, QueryRecords* out, LaunchTimer* timer = nullptr) {
    if (plan.passes == 0 || n == 0) {
        out->qs = d_qs; out->qe = d_qe; out->idx = nullptr;
        return 0;
    }
    char* w = (char*)workspace;
    uint32_t* hist = (uint32_t*)w;
    uint32_t* tickets = (uint32_t*)(w + sizeof(uint32_t) * PT_MAX_PASSES * ((size_t)1 << PT_MAX_BITS));
    uint32_t* status = (uint32_t*)((char*)tickets + 256);
    const int bits = plan.bits;
    const int tbits = bits <= 8 ? 8 : (bits == 9 ? 9 : 10);   // kernel instantiation
    const uint32_t tiles = pt_num_tiles(n);
    SIB_CHECK(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * PT_MAX_PASSES * ((size_t)1 << PT_MAX_BITS) + 256, s));
    {
        int grid = sm_count * 8;
        const uint32_t need = ceil_div_u32(n, PT_THREADS);
        if ((uint32_t)grid > need) grid = (int)need;
        if (timer) timer->begin(TAG_PT_HIST, s);
        pt_histogram_kernel<<<grid, PT_THREADS, sizeof(uint32_t) * ((size_t)plan.passes << tbits), s>>>(
            d_qs, n, plan.key, plan.passes, tbits, hist);
        SIB_CHECK_LAUNCH();
        note_launch();
        pt_prepare_kernel<<<1, 1024, 0, s>>>(hist, plan.passes, tbits);
        SIB_CHECK_LAUNCH();
        if (timer) timer->end(s);
        note_launch();
    }
    int32_t* A = (int32_t*)bufA;
    int32_t* B = (int32_t*)bufB;
    PtPass a;
    a.n = n;
    a.key = plan.key;
    a.dmask = ((uint32_t)1 << bits) - 1u;
    a.status = status;
    a.in_qs = d_qs; a.in_qe = d_qe; a.in_idx = nullptr;
    for (int p = 0; p < plan.passes; ++p) {
        int32_t* dst = (p & 1) ? B : A;
        a.out_qs = dst; a.out_qe = dst + cap; a.out_idx = (uint32_t*)(dst + 2 * cap);
        a.gbase = hist + ((size_t)p << tbits);
        a.ticket = tickets + p;
        a.shift = (uint32_t)(p * bits);
        SIB_CHECK(cudaMemsetAsync(status, 0, sizeof(uint32_t) * (size_t)tiles << tbits, s));
        int rc = tbits == 8 ? pt_launch_pass<8>(a, tiles, s, timer) : tbits == 9 ? pt_launch_pass<9>(a, tiles, s, timer)
                                                                                 : pt_launch_pass<10>(a, tiles, s, timer);
        if (rc) return rc;
        a.in_qs = a.out_qs; a.in_qe = a.out_qe; a.in_idx = a.out_idx;
    }
    out->qs = a.in_qs; out->qe = a.in_qe; out->idx = a.in_idx;
    return 0;
}

}  // namespace sib
