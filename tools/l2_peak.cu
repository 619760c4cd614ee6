// l2_peak.cu -- measured ceilings for the count kernels' rooflines (test/measurement tool, not product).
//
//   gather   random 32-byte sector reads (ld.global.nc.v8.b32, one sector per lane, all lanes of a warp in
//            different lines) from a table of S MB, K independent loads in flight per thread: the access
//            pattern of qk_count_cells_kernel. Reported as G sectors/s and GB/s per table size, so the knee
//            at the L2 capacity and the L1TEX/L2 request-rate ceiling are both visible.
//   stream   coalesced 256-bit read of 8 B/query + 4 B/query write (the count kernels' compulsory stream).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/l2_peak tools/l2_peak.cu
// Run:   tools/bin/l2_peak > gpurun_out/l2_peak.json
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct __align__(32) Rec { uint32_t w[8]; };

__device__ __forceinline__ Rec ld_sector(const Rec* p) {
    Rec r;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t mix(uint32_t x) {   // cheap integer hash: indices cost no memory traffic
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int K, int BYTES>
__global__ void __launch_bounds__(256) gather_kernel(const Rec* __restrict__ tab, uint32_t nrec, uint32_t rounds, uint32_t* __restrict__ sink) {
    const uint32_t gid = blockIdx.x * 256 + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t r = 0; r < rounds; ++r) {
        uint32_t idx[K];
#pragma unroll
        for (int k = 0; k < K; ++k) idx[k] = (uint32_t)(((uint64_t)mix(gid * 977u + r * 131071u + k * 7919u) * nrec) >> 32);
        if (BYTES == 32) {
            Rec v[K];
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] = ld_sector(tab + idx[k]);
#pragma unroll
            for (int k = 0; k < K; ++k) acc += v[k].w[0] ^ v[k].w[7];
        } else {
            uint32_t v[K];
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] = __ldg(reinterpret_cast<const uint32_t*>(tab + idx[k]));
#pragma unroll
            for (int k = 0; k < K; ++k) acc += v[k];
        }
    }
    if (acc == 0x12345u) sink[0] = acc;
}

__global__ void __launch_bounds__(256) stream_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, size_t n4) {
    // per 4 queries: 16 B of a, 16 B of b in; 16 B out
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
        uint4 x, y;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(a + i));
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(y.x), "=r"(y.y), "=r"(y.z), "=r"(y.w) : "l"(b + i));
        __stcs(out + i, make_uint4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w));
    }
}

// the count kernels' own shape: one sector per thread, one CTA of 128 threads per 128 requests, no loop
__global__ void __launch_bounds__(128) flat_kernel(const Rec* __restrict__ tab, uint32_t nrec, uint32_t salt, uint32_t* __restrict__ sink) {
    const uint32_t gid = blockIdx.x * 128 + threadIdx.x;
    const Rec v = ld_sector(tab + (uint32_t)(((uint64_t)mix(gid * 977u + salt) * nrec) >> 32));
    if ((v.w[0] ^ v.w[7]) == 0x12345u) sink[0] = v.w[0];
}
static double run_flat(const Rec* tab, uint32_t nrec, uint32_t* sink, uint64_t requests) {
    const uint32_t grid = (uint32_t)(requests / 128);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    flat_kernel<<<grid / 4, 128>>>(tab, nrec, 1u, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        flat_kernel<<<grid, 128>>>(tab, nrec, 7u + it, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return (double)grid * 128 / (best * 1e-3);
}

template <int K, int BYTES>
static double run_gather(const Rec* tab, uint32_t nrec, int ctas_per_sm, int sms, uint32_t* sink, uint64_t target_loads) {
    const int grid = sms * ctas_per_sm;
    uint32_t rounds = (uint32_t)(target_loads / ((uint64_t)grid * 256 * K));
    if (rounds < 1) rounds = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    gather_kernel<K, BYTES><<<grid, 256>>>(tab, nrec, rounds / 4 + 1, sink);   // warm: table into L2
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        gather_kernel<K, BYTES><<<grid, 256>>>(tab, nrec, rounds, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return (double)grid * 256 * K * rounds / (best * 1e-3);   // loads per second
}

int main() {
    int dev = 0, sms = 0, l2 = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
    uint32_t* sink; CK(cudaMalloc(&sink, 64));
    const size_t max_bytes = (size_t)512 << 20;
    Rec* tab; CK(cudaMalloc(&tab, max_bytes)); CK(cudaMemset(tab, 1, max_bytes));
    printf("{\"sms\": %d, \"l2_bytes\": %d, \"gather\": [", sms, l2);
    const double sizes_mb[] = {8, 16, 32, 48, 62.5, 80, 96, 128, 256, 512};
    const int occs[] = {2, 4, 8};
    bool first = true;
    const uint64_t target = 400000000ull;
    for (double mb : sizes_mb) {
        const uint32_t nrec = (uint32_t)(mb * 1048576.0 / 32);
        for (int occ : occs) {
            const double r4 = run_gather<4, 32>(tab, nrec, occ, sms, sink, target);
            const double r8 = run_gather<8, 32>(tab, nrec, occ, sms, sink, target);
            const double w8 = run_gather<8, 4>(tab, nrec, occ, sms, sink, target);
            const double fl = occ == 2 ? run_flat(tab, nrec, sink, target) : 0.0;
            printf("%s\n {\"table_mb\": %.1f, \"ctas_per_sm\": %d, \"sectors_per_s_flat\": %.4g, \"sectors_per_s_k4\": %.4g, \"sectors_per_s_k8\": %.4g, \"gbs_k8\": %.1f, \"words_per_s_k8\": %.4g}",
                   first ? "" : ",", mb, occ, fl, r4, r8, r8 * 32 / 1e9, w8);
            first = false;
        }
    }
    printf("],\n");
    // compulsory stream of a count: 8 B in + 4 B out per query
    const size_t nq = (size_t)100000000, n4 = nq / 4;
    uint4 *a, *b, *o;
    CK(cudaMalloc(&a, n4 * 16)); CK(cudaMalloc(&b, n4 * 16)); CK(cudaMalloc(&o, n4 * 16));
    CK(cudaMemset(a, 1, n4 * 16)); CK(cudaMemset(b, 2, n4 * 16));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        CK(cudaEventRecord(e0));
        stream_kernel<<<sms * 8, 256>>>(a, b, o, n4);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it && ms < best) best = ms;
    }
    printf(" \"stream\": {\"queries\": %zu, \"bytes\": %zu, \"ms\": %.4f, \"gbs\": %.1f}}\n", nq, nq * 12, best, nq * 12 / (best * 1e-3) / 1e9);
    return 0;
}
