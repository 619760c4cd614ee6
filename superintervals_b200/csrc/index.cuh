// index.cuh -- the device-resident index object behind siIndex* (host-side C++).
#pragma once

#include "common.cuh"

struct siIndex;   // C-visible opaque name

namespace sib {

// cudaMalloc'd buffer that only ever grows (reallocation drops old contents)
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

void note_launch(unsigned n = 1);
unsigned long long launches();

}  // namespace sib

struct siIndex {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;

    // ---- the index, position order -------------------------------------------------
    uint32_t n = 0, n_padded = 0;
    bool built = false;
    bool wellformed = false;   // every interval has start <= end (enables the count shortcut)
    sib::DevBuf starts, ends, values, branch, perm;
    sib::DevBuf esort;         // ends, each aligned 32-block sorted ascending (count sweep)
    sib::DevBuf tree;          // 32-ary max tree levels + prefix-max levels
    const int32_t* pmax32 = nullptr;   // inside `tree`: exclusive prefix max of ends per 32-block

    // ---- build scratch (released after build for large n) --------------------------
    sib::DevBuf b_in_s, b_in_e, b_in_v;        // staged host inputs
    sib::DevBuf b_kA, b_kB, b_vA, b_vB, b_ws;  // 64-bit key sort

    // ---- query scratch ---------------------------------------------------------------
    sib::DevBuf small;                          // device scalars: [0] flags/sorted, ...
    sib::DevBuf q_kA, q_kB, q_vA, q_vB, q_ws;   // 32-bit key sort of a query batch
    sib::DevBuf scan_status;
    const int32_t* plan_qs = nullptr;           // query batch the cached sort belongs to
    size_t plan_n = 0;
    bool plan_valid = false;
    bool plan_armed = false;                    // set by siSortQueriesDevice: next count may reuse

    // ---- staging for the host-buffer API ---------------------------------------------
    sib::DevBuf h_qs, h_qe, h_counts, h_offsets, h_out, h_cov;
    bool pipe_ready = false;                    // chunked host-batch pipeline (c_abi.cu)
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t e_in[2] = {nullptr, nullptr}, e_k[2] = {nullptr, nullptr}, e_out[2] = {nullptr, nullptr};
    void* pinned = nullptr;                     // small pinned scratch for scalars
    size_t pinned_bytes = 0;

    size_t device_bytes() const;
};
