// host_common.cuh -- host-side pieces shared by c_abi.cu and setops.cu: the handle behind a
// cSuperIntervals*, result-buffer growth, and the pinned-staged host <-> device copies.
#pragma once

#include "../../include/superintervals_b200.h"
#include "common.cuh"
#include "index.cuh"

#include <cstdlib>

namespace sib {

constexpr uint64_t HANDLE_MAGIC = 0x53495F4232303021ull;   // "SI_B200!"

// The handle is over-allocated: callers see the public cSuperIntervals prefix
// (c_superintervals.h:81-91), the library keeps the device index behind it.
struct Handle {
    cSuperIntervals pub;   // must stay first: &handle->pub is what callers hold
    uint64_t magic;
    siIndex* ix;
    bool mirror;
    bool indexed;
};

inline Handle* H(cSuperIntervals* si) { return reinterpret_cast<Handle*>(si); }
inline const Handle* H(const cSuperIntervals* si) { return reinterpret_cast<const Handle*>(si); }

// true when the handle carries a built device index; latches an error otherwise
bool handle_ready(Handle* h, const char* who);

// Result buffers (cIndexResult / cKeyResult / cItemResult .data) are library-owned memory that the
// reference grows by realloc and frees in destroy*Result (c.h:585-588, 1066-1089). Large ones are
// page-aligned malloc memory REGISTERED with the CUDA driver (cudaHostRegister), so that the device
// copies results straight into them by DMA instead of through a pinned staging slot and a host memcpy.
// They remain ordinary malloc memory to the caller; result_free() un-registers before freeing.
void* result_realloc(void* old, size_t old_bytes, size_t new_bytes);
void result_free(void* p);

template <typename R>
bool grow(R* r, size_t need_total, size_t elem) {
    if (need_total <= r->capacity) return true;
    size_t cap = r->capacity ? r->capacity : 16;   // reference growth: x2 from 16 (c.h:585-588)
    while (cap < need_total) cap *= 2;
    void* p = result_realloc(r->data, r->size * elem, cap * elem);
    if (!p) { set_error_msg(cudaErrorMemoryAllocation, "realloc of result buffer failed"); return false; }
    r->data = reinterpret_cast<decltype(r->data)>(p);
    r->capacity = cap;
    return true;
}

// device -> host; returns when the bytes are in dst (pageable buffers go through pinned staging)
int copy_d2h(siIndex* ix, void* dst, const void* src_dev, size_t bytes, cudaStream_t s);
// host -> device, stream-ordered on s; src may be reused when the call returns
int copy_h2d(siIndex* ix, void* dst_dev, const void* src, size_t bytes, cudaStream_t s);
// upload a query batch into the index's staging buffers (h_qs, h_qe)
int stage_queries(siIndex* ix, const int32_t* qs, const int32_t* qe, size_t n);

}  // namespace sib
