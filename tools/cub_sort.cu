// cub_sort.cu -- YARDSTICK, not product: the time cub::DeviceRadixSort::SortPairs (the CUDA toolkit's tuned onesweep)
// needs for the same job build() does with csrc/radix_sort.cuh -- N pairs of (64-bit key, 32-bit payload), all 64 key
// bits, and N 32-bit keys alone (the `eall` sort). bench.py prints it beside the build's own kernel times
// (SURVEY 7 hard part 8). Nothing in superintervals_b200/ includes CUB.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/cub_sort tools/cub_sort.cu
// Run:   tools/bin/cub_sort [n = 10000000] [reps = 5]      -> one JSON line
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void fill(uint64_t* k, uint32_t* v, uint32_t* k32, size_t n, uint64_t seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t x = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
        x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
        // like build()'s keys: a start below 250e6 in the high word, an end a few kb above it in the low word
        const uint32_t s = (uint32_t)(x % 250000000ull), e = s + 150u + (uint32_t)((x >> 32) % 9850ull);
        k[i] = ((uint64_t)(s ^ 0x80000000u) << 32) | (uint32_t)~(e ^ 0x80000000u);
        v[i] = (uint32_t)i;
        k32[i] = e ^ 0x80000000u;
    }
}

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? (size_t)atoll(argv[1]) : 10000000;
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    uint64_t *k0, *k1; uint32_t *v0, *v1, *a0, *a1;
    CK(cudaMalloc(&k0, n * 8)); CK(cudaMalloc(&k1, n * 8)); CK(cudaMalloc(&v0, n * 4)); CK(cudaMalloc(&v1, n * 4));
    CK(cudaMalloc(&a0, n * 4)); CK(cudaMalloc(&a1, n * 4));
    size_t tb = 0, tb2 = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, v1, n, 0, 64));
    CK(cub::DeviceRadixSort::SortKeys(nullptr, tb2, a0, a1, n, 0, 32));
    void* tmp; CK(cudaMalloc(&tmp, tb > tb2 ? tb : tb2));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best_pairs = 1e30f, best_keys = 1e30f;
    for (int r = 0; r < reps + 1; ++r) {
        fill<<<1184, 256>>>(k0, v0, a0, n, 17 + r);
        CK(cudaEventRecord(e0));
        CK(cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, v0, v1, n, 0, 64));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r && ms < best_pairs) best_pairs = ms;
        CK(cudaEventRecord(e0));
        CK(cub::DeviceRadixSort::SortKeys(tmp, tb2, a0, a1, n, 0, 32));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r && ms < best_keys) best_keys = ms;
    }
    printf("{\"n\": %zu, \"cub_sort_pairs_u64_u32_ms\": %.4f, \"cub_sort_keys_u32_ms\": %.4f, \"reps\": %d, "
           "\"what\": \"cub::DeviceRadixSort (CUDA toolkit), best of reps, CUDA events; yardstick only\"}\n",
           n, best_pairs, best_keys, reps);
    return 0;
}
