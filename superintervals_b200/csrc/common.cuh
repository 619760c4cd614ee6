// common.cuh -- shared device/host helpers for libsuperintervals_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace sib {

constexpr uint32_t NONE32 = 0xFFFFFFFFu;   // device twin of SI_NONE / SIZE_MAX
constexpr uint32_t FULL_MASK = 0xFFFFFFFFu;

// Sticky error side channel (the reference ABI has no error returns).
void set_error(cudaError_t e, const char* what, const char* file, int line);
void set_error_msg(int code, const char* msg);
int last_error_code();

#define SIB_CHECK(expr)                                                        \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) {                                               \
            ::sib::set_error(_e, #expr, __FILE__, __LINE__);                   \
            return (int)_e;                                                    \
        }                                                                      \
    } while (0)

#define SIB_CHECK_LAUNCH() SIB_CHECK(cudaGetLastError())

__host__ __device__ inline uint32_t ceil_div_u32(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

// ---- cache-hinted loads -----------------------------------------------------------
// Index arrays (starts/ends/branch/values) are read-only for the life of a query
// kernel: read them through the non-coherent path so L1 keeps the hot window.
__device__ __forceinline__ int32_t ld_nc(const int32_t* p) { return __ldg(p); }
__device__ __forceinline__ uint32_t ld_nc(const uint32_t* p) { return __ldg(p); }
__device__ __forceinline__ int4 ld_nc4(const int32_t* p) { return __ldg(reinterpret_cast<const int4*>(p)); }

// Streaming (touch-once) loads/stores for query and result arrays: keep them out of L1.
__device__ __forceinline__ int32_t ld_stream(const int32_t* p) {
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint64_t ld_stream(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void st_stream(uint32_t* p, uint32_t v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(uint64_t* p, uint64_t v) {
    __stcs(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Order-preserving maps between signed coordinates and unsigned radix keys.
__host__ __device__ inline uint32_t flip_i32(int32_t v) { return (uint32_t)v ^ 0x80000000u; }
__host__ __device__ inline int32_t unflip_i32(uint32_t k) { return (int32_t)(k ^ 0x80000000u); }

}  // namespace sib
