// stream_kernels.cuh -- count for POSITION-SORTED query batches at streaming speed (sm_100a).
//
// The reference answers a sorted batch faster than a shuffled one because consecutive
// queries touch the same cache lines (superintervals.hpp:501-513 binary search + 651-825 walk;
// test/generate_test_intervals.py:43-51 runs the sorted case). Here a CTA owns a contiguous
// tile of the sorted batch, so everything its queries need from the index is ONE contiguous
// window per table: the window is copied into shared memory by the TMA engine
// (cp.async.bulk, completion on an mbarrier) and every rank is answered from shared memory.
//
// Closed form, as in qk_count_cells_kernel (well-formed index, qs <= qe):
//     count = #{starts <= qe} - #{ends < qs}
// Table: RankBits, one entry per 32 coordinates of a sorted array A,
//     t[k]  = { #{A < lo + 32k} | flags,  b1 = coordinates of the word holding >= 1 value }
//     d2[k] = coordinates of the word holding >= 2 values        (read only when RB_DUP is set)
//     #{A < x} = prefix + popc(b1 & below(x)) + popc(d2 & below(x))
// -- one 8-byte shared-memory load, a mask and a popc per rank. A word holding a coordinate
// with three or more values (RB_SLOW) is answered from the rank cells in global memory.
// 12 B per 32 coordinates: read once per pass over the batch, sequentially.
//
// Per tile of 2048 queries: 16 KB of queries in (256-bit loads), the two windows in (TMA),
// 8 KB of counts out (256-bit stores). Nothing depends on the batch REALLY being sorted:
// the window is [min, max] over the tile, and a tile whose window does not fit the staging
// buffers (an unsorted or very sparse tile) answers each query from the rank cells instead.
// Queries with qs > qe (quirk Q6) take the branch-array walk, as in every count kernel.
#pragma once

#include "query_kernels.cuh"

namespace sib {

constexpr uint32_t RB_DUP = 0x80000000u;    // d2[k] != 0
constexpr uint32_t RB_SLOW = 0x40000000u;   // some coordinate of the word holds >= 3 values
constexpr uint32_t RB_MASK = 0x3FFFFFFFu;

// ---- build: one thread per word; ranks at the word borders come from the rank cells -------------
__global__ void __launch_bounds__(256)
sk_rank_bits_kernel(RankCells rc, const int32_t* __restrict__ A, uint32_t n, uint32_t nwords, uint32_t nwords_padded,
                    uint2* __restrict__ t, uint32_t* __restrict__ d2, unsigned long long* __restrict__ slow_words) {
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x; k < nwords_padded; k += stride) {
        if (k >= nwords) { t[k] = make_uint2(n, 0u); d2[k] = 0u; continue; }   // beyond A[n-1]: everything is below
        const int64_t v0 = (int64_t)rc.lo + (int64_t)(k << 5);
        const uint32_t p0 = k ? cells_rank_lt(rc, A, v0) : 0u;
        const uint32_t p1 = cells_rank_lt(rc, A, v0 + 32);
        uint32_t b1 = 0, b2 = 0, flags = 0;
        int32_t prev = 0;
        uint32_t run = 0;
        for (uint32_t i = p0; i < p1; ++i) {
            const int32_t a = ld_nc(A + i);
            run = (i > p0 && a == prev) ? run + 1u : 1u;
            prev = a;
            const uint32_t bit = 1u << (uint32_t)((int64_t)a - v0);
            if (run == 1u) b1 |= bit;
            else if (run == 2u) b2 |= bit;
            else flags |= RB_SLOW;
        }
        if (b2) flags |= RB_DUP;
        if (flags & RB_SLOW) atomicAdd(slow_words, 1ull);
        t[k] = make_uint2(p0 | flags, b1);
        d2[k] = b2;
    }
}

// ---- PTX: mbarrier + TMA bulk copy (SASS: SYNCS / UBLKCP) ------------------------------------------
__device__ __forceinline__ uint32_t sk_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sk_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sk_smem(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sk_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sk_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sk_smem(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(sk_smem(bar)) : "memory");
}
__device__ __forceinline__ void sk_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(sk_smem(bar)), "r"(parity) : "memory");
    }
}

struct __align__(32) Vec8 { uint32_t w[8]; };
__device__ __forceinline__ Vec8 sk_ld8(const int32_t* p) {
    Vec8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void sk_st8(uint32_t* p, const uint32_t* c) {
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7]) : "memory");
}
__device__ __forceinline__ void sk_st8(uint64_t* p, const uint32_t* c) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
        __stcs(reinterpret_cast<ulonglong2*>(p) + k, make_ulonglong2((unsigned long long)c[2 * k], (unsigned long long)c[2 * k + 1]));
}

// ---- the kernel ------------------------------------------------------------------------------------------
// Persistent CTAs, software-pipelined over their tiles (tile = blockIdx.x + k * gridDim.x):
//     while tile k is being ranked out of shared memory (stage k & 1),
//       the TMA copies of tile k+1's windows are in flight into the other stage (its mbarrier), and
//       the 256-bit loads of tile k+2's queries are in flight into registers.
// One __syncthreads per tile (the window reduction; it also frees the stage the next copies overwrite);
// the window itself is published through the stage's mbarrier (thread 0 writes it before it arrives).
#ifndef SIB_SK_THREADS
#define SIB_SK_THREADS 256
#endif
constexpr int SK_THREADS = SIB_SK_THREADS;
constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int SK_PER_THREAD = 8;
constexpr uint32_t SK_TILE = SK_THREADS * SK_PER_THREAD;   // 2048 queries: 16 KB in, 8 KB out
#ifndef SIB_SK_SW
#define SIB_SK_SW 1536
#endif
#ifndef SIB_SK_EW
#define SIB_SK_EW 768
#endif
// 3 CTAs/SM (80 registers, 24 bytes of spills, 3 x 55 KB of staging): 0.347 -> 0.312 ms on C2 sorted against 2 CTAs/SM at 114
// registers; 4 CTAs/SM with 36 KB stages 0.308 ms, 3 with 36 KB stages 0.305 ms (tools/gpu_r02zi.sh) -- the wider windows are kept
#ifndef SIB_SK_MINBLOCKS
#define SIB_SK_MINBLOCKS 3
#endif
constexpr uint32_t SK_SW = SIB_SK_SW;   // staged words of the starts table (x 32 coordinates)
constexpr uint32_t SK_EW = SIB_SK_EW;   // staged words of the ends table
constexpr size_t SK_STAGE = (size_t)(SK_SW + SK_EW) * 12;
constexpr size_t SK_SMEM = 2 * SK_STAGE;

struct SkStage {
    uint2* Ts; uint2* Te; uint32_t* Ds; uint32_t* De;
};
__device__ __forceinline__ SkStage sk_stage(unsigned char* base, uint32_t stage) {
    unsigned char* p = base + stage * SK_STAGE;
    SkStage st;
    st.Ts = reinterpret_cast<uint2*>(p);
    st.Te = st.Ts + SK_SW;
    st.Ds = reinterpret_cast<uint32_t*>(st.Te + SK_EW);
    st.De = st.Ds + SK_SW;
    return st;
}
__device__ __forceinline__ void sk_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sk_smem(bar)) : "memory");
}

// fail_list / fail_count: tiles this kernel does not answer -- their window does not fit the staging
// buffers (an unsorted or very sparse tile) or they hold a query with qs > qe (quirk Q6: the closed
// form does not apply) -- are appended here and answered by sk_count_failed_tiles_kernel right after.
template <typename CountT>
__global__ void __launch_bounds__(SK_THREADS, SIB_SK_MINBLOCKS)
sk_count_stream_kernel(IndexView ix, const int32_t* __restrict__ qs_in, const int32_t* __restrict__ qe_in, uint32_t nq,
                       CountT* __restrict__ counts, uint32_t vec_ok, uint32_t ntiles, uint32_t* __restrict__ fail_list,
                       uint32_t* __restrict__ fail_count, const __grid_constant__ Fanout fan) {
    extern __shared__ __align__(128) unsigned char sk_sh[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_red[2][SK_WARPS][5];
    __shared__ uint32_t s_win[2][4];   // per stage: first staged word of each table, staged?

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const RankBits bs = ix.bits_s, be = ix.bits_e;
    const int32_t lo_s = bs.lo, lo_e = be.lo;
    const uint32_t span_s = bs.span, span1_e = be.span + 1u;

    uint32_t tile = blockIdx.x;
    if (tile >= ntiles) return;
    if (tid == 0) { sk_mbar_init(&s_bar[0], 1); sk_mbar_init(&s_bar[1], 1); }

    int32_t qs[SK_PER_THREAD], qe[SK_PER_THREAD];
    // the queries of one tile into registers; lanes past the end of the batch repeat its last query: they stay
    // inside the tile's window (and add nothing to it, to its inverted flag or to the output)
    auto load_queries = [&](uint32_t t) {
        const uint64_t base = (uint64_t)t * SK_TILE + (uint64_t)tid * SK_PER_THREAD;
        if (base + SK_PER_THREAD <= nq && vec_ok) {
            const Vec8 a = sk_ld8(qs_in + base), b = sk_ld8(qe_in + base);
#pragma unroll
            for (int j = 0; j < SK_PER_THREAD; ++j) { qs[j] = (int32_t)a.w[j]; qe[j] = (int32_t)b.w[j]; }
        } else {
#pragma unroll 1
            for (int j = 0; j < SK_PER_THREAD; ++j) {
                const uint64_t q = base + j < nq ? base + j : (uint64_t)nq - 1;
                const int32_t a = ld_stream(qs_in + q), b = ld_stream(qe_in + q);
#pragma unroll
                for (int k = 0; k < SK_PER_THREAD; ++k)
                    if (k == j) { qs[k] = a; qe[k] = b; }
            }
        }
    };
    // where each rank falls (32-bit arithmetic: build() refuses rank bits on a span that would overflow it):
    //   ds = clamp(qe + 1 - lo_s, 0, span_s + 1)      #{starts <= qe} = #{starts < qe + 1}
    //   de = clamp(qs - lo_e, 0, span_e + 1)          #{ends < qs}
    // and this warp's share of the tile's window of words in either table
    auto place = [&](uint32_t (&ds)[SK_PER_THREAD], uint32_t (&de)[SK_PER_THREAD], uint32_t buf) {
        uint32_t smin = 0xFFFFFFFFu, smax = 0, emin = 0xFFFFFFFFu, emax = 0, inv = 0;
#pragma unroll
        for (int j = 0; j < SK_PER_THREAD; ++j) {
            ds[j] = qe[j] < lo_s ? 0u : min((uint32_t)qe[j] - (uint32_t)lo_s, span_s) + 1u;
            de[j] = qs[j] <= lo_e ? 0u : min((uint32_t)qs[j] - (uint32_t)lo_e, span1_e);
            smin = min(smin, ds[j]); smax = max(smax, ds[j]);
            emin = min(emin, de[j]); emax = max(emax, de[j]);
            inv |= qs[j] > qe[j] ? 1u : 0u;
        }
        smin = __reduce_min_sync(FULL_MASK, smin); smax = __reduce_max_sync(FULL_MASK, smax);
        emin = __reduce_min_sync(FULL_MASK, emin); emax = __reduce_max_sync(FULL_MASK, emax);
        inv = __reduce_or_sync(FULL_MASK, inv);
        if (lane == 0) { s_red[buf][warp][0] = smin; s_red[buf][warp][1] = smax; s_red[buf][warp][2] = emin; s_red[buf][warp][3] = emax; s_red[buf][warp][4] = inv; }
    };
    // thread 0, after the barrier: the tile's window, its TMA copies into `stage`, completion on the stage's mbarrier
    auto issue = [&](uint32_t t, uint32_t stage, uint32_t buf) {
        uint32_t smin = 0xFFFFFFFFu, smax = 0, emin = 0xFFFFFFFFu, emax = 0, inv = 0;
#pragma unroll
        for (int w = 0; w < SK_WARPS; ++w) {
            smin = min(smin, s_red[buf][w][0]); smax = max(smax, s_red[buf][w][1]);
            emin = min(emin, s_red[buf][w][2]); emax = max(emax, s_red[buf][w][3]);
            inv |= s_red[buf][w][4];
        }
        // windows in words, first word aligned to 4 (16-byte granules of both arrays), length a multiple of 4
        const uint32_t ks0 = (smin >> 5) & ~3u, ke0 = (emin >> 5) & ~3u;
        const uint32_t ls = (((smax >> 5) - ks0) | 3u) + 1u, le = (((emax >> 5) - ke0) | 3u) + 1u;
        const bool ok = !inv && ls <= SK_SW && le <= SK_EW;
        s_win[stage][0] = ks0; s_win[stage][1] = ke0; s_win[stage][2] = ok ? 1u : 0u;
        uint64_t* bar = &s_bar[stage];
        if (ok) {
            const SkStage st = sk_stage(sk_sh, stage);
            sk_mbar_expect_tx(bar, (ls + le) * 12u);      // release: s_win is visible to whoever sees the phase complete
            sk_bulk_g2s(st.Ts, bs.t + ks0, ls * 8u, bar);
            sk_bulk_g2s(st.Te, be.t + ke0, le * 8u, bar);
            sk_bulk_g2s(st.Ds, bs.d2 + ks0, ls * 4u, bar);
            sk_bulk_g2s(st.De, be.d2 + ke0, le * 4u, bar);
        } else {
            fail_list[atomicAdd(fail_count, 1u)] = t;
            sk_mbar_arrive(bar);                          // nothing to wait for: the phase completes at once
        }
    };
    // every rank of one tile from its staged windows
    auto rank_tile = [&](uint32_t t, uint32_t stage, uint32_t parity, const uint32_t (&ds)[SK_PER_THREAD], const uint32_t (&de)[SK_PER_THREAD]) {
        sk_mbar_wait(&s_bar[stage], parity);
        if (s_win[stage][2] == 0) return;
        const uint32_t ks0 = s_win[stage][0], ke0 = s_win[stage][1];
        const SkStage st = sk_stage(sk_sh, stage);
        uint32_t c[SK_PER_THREAD];
#pragma unroll
        for (int j = 0; j < SK_PER_THREAD; ++j) {
            const uint32_t k1 = (ds[j] >> 5) - ks0, k2 = (de[j] >> 5) - ke0;
            const uint2 e1 = st.Ts[k1], e2 = st.Te[k2];
            const uint32_t m1 = (1u << (ds[j] & 31u)) - 1u, m2 = (1u << (de[j] & 31u)) - 1u;
            uint32_t ns = (e1.x & RB_MASK) + __popc(e1.y & m1);
            uint32_t ne = (e2.x & RB_MASK) + __popc(e2.y & m2);
            if (e1.x & RB_DUP) ns += __popc(st.Ds[k1] & m1);
            if (e2.x & RB_DUP) ne += __popc(st.De[k2] & m2);
            if (__builtin_expect(((e1.x | e2.x) & RB_SLOW) != 0, 0)) {   // a coordinate with >= 3 values in one of the words: rank cells (rare)
                // a slow word lies inside the table's span, where d was not clamped: the coordinate is lo + d
                if (e1.x & RB_SLOW) ns = cells_rank_lt(ix.cells_s, ix.starts, (int64_t)lo_s + ds[j]);
                if (e2.x & RB_SLOW) ne = cells_rank_lt(ix.cells_e, ix.eall, (int64_t)lo_e + de[j]);
            }
            c[j] = ns - ne;
        }
        const uint64_t base = (uint64_t)t * SK_TILE + (uint64_t)tid * SK_PER_THREAD;
        if (base + SK_PER_THREAD <= nq && vec_ok) {
            sk_st8(counts + base, c);
            if (sizeof(CountT) == 4)
                for (int f = 0; f < fan.n; ++f) sk_st8(fan.p[f] + base, c);      // peer copies of the gathered vector (vec_ok covers them)
        } else {
#pragma unroll
            for (int j = 0; j < SK_PER_THREAD; ++j)
                if (base + j < nq) { counts[base + j] = (CountT)c[j]; fan_store<CountT>(fan, base + j, c[j]); }
        }
    };

    uint32_t dsA[SK_PER_THREAD], deA[SK_PER_THREAD], dsB[SK_PER_THREAD], deB[SK_PER_THREAD];
    load_queries(tile);
    place(dsA, deA, 0);
    __syncthreads();                       // s_red[0] complete; the mbarrier inits are visible
    if (tid == 0) issue(tile, 0, 0);
    uint32_t next = tile + gridDim.x;
    if (next < ntiles) load_queries(next);
    for (uint32_t k = 0;; ++k) {
        const uint32_t stage = k & 1u, parity = (k >> 1) & 1u;
        const bool has_next = next < ntiles;
        if (has_next) place(dsB, deB, (k + 1u) & 1u);
        __syncthreads();                   // s_red complete; everyone has left tile k-1, whose stage the next copies overwrite
        if (has_next) {
            if (tid == 0) issue(next, stage ^ 1u, (k + 1u) & 1u);
            if (next + gridDim.x < ntiles) load_queries(next + gridDim.x);
        }
        rank_tile(tile, stage, parity, dsA, deA);
        if (!has_next) break;
#pragma unroll
        for (int j = 0; j < SK_PER_THREAD; ++j) { dsA[j] = dsB[j]; deA[j] = deB[j]; }
        tile = next;
        next += gridDim.x;
    }
}

// The tiles the streaming kernel handed back, by the rank-cells code (any order, inverted queries walk).
// Persistent grid: a sorted batch leaves the list empty and the kernel returns at once.
template <typename CountT>
__global__ void __launch_bounds__(QC_THREADS)
sk_count_failed_tiles_kernel(IndexView ix, const int32_t* __restrict__ qs_in, const int32_t* __restrict__ qe_in, uint32_t nq,
                             CountT* __restrict__ counts, const uint32_t* __restrict__ fail_list,
                             const uint32_t* __restrict__ fail_count, const __grid_constant__ Fanout fan) {
    const uint32_t nfail = *fail_count;
    const QueryRecords rec{qs_in, qe_in, nullptr};
    constexpr uint32_t PARTS = SK_TILE / QC_TILE;
    for (uint32_t w = blockIdx.x; w < nfail * PARTS; w += gridDim.x) {
        const uint64_t base = (uint64_t)fail_list[w / PARTS] * SK_TILE + (uint64_t)(w % PARTS) * QC_TILE;
        if (base < nq) count_cells_tile<CountT, false>(ix, rec, base, nq, counts, fan);
    }
}

}  // namespace sib
