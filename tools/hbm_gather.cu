// hbm_gather.cu -- what HBM delivers for the rank-cell gathers of an index larger than L2 (measurement tool, not product).
//
// qk_count_cells_kernel reads two 32-byte sectors per query: one rank cell of the starts table, one of the ends
// table. When the tables do not fit L2 (C4, C5) every sector is a DRAM access. This tool measures, per table size,
//   single   one random 32-byte sector per request                         (today's cost per rank)
//   pair     two ADJACENT sectors (one 64-byte aligned pair) per request   (both ranks of a stabbing / short query
//            from an interleaved table: starts cell i and ends cell i side by side)
//   line     four adjacent sectors (one 128-byte line) per request
// as requests/s, so that "is the cost per sector or per DRAM access" is a number.
//
// Build: make -C tools      Run: tools/bin/hbm_gather > gpurun_out/hbm_gather.json
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct __align__(32) Rec { uint32_t w[8]; };

__device__ __forceinline__ Rec ld_sector(const Rec* p) {
    Rec r;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// SECTORS adjacent sectors per request, K requests in flight per thread
template <int K, int SECTORS>
__global__ void __launch_bounds__(128) gather_kernel(const Rec* __restrict__ tab, uint32_t ngroups, uint32_t rounds, uint32_t* __restrict__ sink) {
    const uint32_t gid = blockIdx.x * 128 + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t r = 0; r < rounds; ++r) {
        size_t idx[K];
#pragma unroll
        for (int k = 0; k < K; ++k)
            idx[k] = (size_t)(((uint64_t)mix(gid * 977u + r * 131071u + k * 7919u) * ngroups) >> 32) * SECTORS;
        Rec v[K][SECTORS];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int s = 0; s < SECTORS; ++s) v[k][s] = ld_sector(tab + idx[k] + s);
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int s = 0; s < SECTORS; ++s) acc += v[k][s].w[0] ^ v[k][s].w[7];
    }
    if (acc == 0x12345u) sink[0] = acc;
}

// the count kernels' own shape: ONE request per thread, one CTA of 128 threads per 128 requests, no loop
__global__ void __launch_bounds__(128) flat_kernel(const Rec* __restrict__ tab, uint32_t ngroups, uint32_t salt, uint32_t* __restrict__ sink) {
    const uint32_t gid = blockIdx.x * 128 + threadIdx.x;
    const size_t idx = (size_t)(((uint64_t)mix(gid * 977u + salt) * ngroups) >> 32);
    const Rec v = ld_sector(tab + idx);
    if ((v.w[0] ^ v.w[7]) == 0x12345u) sink[0] = v.w[0];
}
static double run_flat(const Rec* tab, size_t nrec, uint32_t* sink, uint64_t requests) {
    const uint32_t grid = (uint32_t)(requests / 128);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    flat_kernel<<<grid / 4, 128>>>(tab, (uint32_t)nrec, 1u, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        flat_kernel<<<grid, 128>>>(tab, (uint32_t)nrec, 7u + it, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return (double)grid * 128 / (best * 1e-3);
}

template <int K, int SECTORS>
static double run(const Rec* tab, size_t nrec, int sms, uint32_t* sink, uint64_t target_requests) {
    const int grid = sms * 16;
    const uint32_t ngroups = (uint32_t)(nrec / SECTORS);
    uint32_t rounds = (uint32_t)(target_requests / ((uint64_t)grid * 128 * K));
    if (rounds < 1) rounds = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    gather_kernel<K, SECTORS><<<grid, 128>>>(tab, ngroups, rounds / 4 + 1, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        gather_kernel<K, SECTORS><<<grid, 128>>>(tab, ngroups, rounds, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return (double)grid * 128 * K * rounds / (best * 1e-3);
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint32_t* sink; CK(cudaMalloc(&sink, 64));
    const size_t max_bytes = (size_t)8 << 30;
    Rec* tab; CK(cudaMalloc(&tab, max_bytes)); CK(cudaMemset(tab, 1, max_bytes));
    printf("{\"sms\": %d, \"gather\": [", sms);
    const double sizes_gb[] = {0.25, 1, 4, 8};
    bool first = true;
    const uint64_t target = 200000000ull;
    // the L2 fetch granularity is a device-wide hint (cudaLimitMaxL2FetchGranularity, default 64): 0 = leave it alone
    const size_t grans[] = {0, 32};
    for (size_t gran : grans) {
        size_t got = 0;
        if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
        CK(cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity));
        for (double gb : sizes_gb) {
            const size_t nrec = (size_t)(gb * 1073741824.0 / 32);
            const double sf = run_flat(tab, nrec, sink, target);
            const double s0 = run<1, 1>(tab, nrec, sms, sink, target);
            const double s1 = run<2, 1>(tab, nrec, sms, sink, target);
            const double s1b = run<4, 1>(tab, nrec, sms, sink, target);
            const double s1c = run<8, 1>(tab, nrec, sms, sink, target);
            const double s2 = run<2, 2>(tab, nrec, sms, sink, target);
            const double s2b = run<4, 2>(tab, nrec, sms, sink, target);
            const double s4 = run<2, 4>(tab, nrec, sms, sink, target);
            printf("%s\n {\"l2_fetch_granularity\": %zu, \"table_gb\": %.2f, \"single_per_s_flat\": %.4g, \"single_per_s_k1\": %.4g, \"single_per_s_k2\": %.4g, \"single_per_s_k4\": %.4g, \"single_per_s_k8\": %.4g, "
                   "\"pair_per_s_k2\": %.4g, \"pair_per_s_k4\": %.4g, \"line_per_s_k2\": %.4g}",
                   first ? "" : ",", got, gb, sf, s0, s1, s1b, s1c, s2, s2b, s4);
            first = false;
        }
    }
    printf("]}\n");
    return 0;
}
