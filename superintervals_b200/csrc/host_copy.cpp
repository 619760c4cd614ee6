// host_copy.cpp -- the host half of a staged device -> host copy: pinned staging buffer -> the
// caller's (pageable) result array. The destination is written once and not read by this
// library again, so the copy uses non-temporal stores where the CPU has AVX2: no read-for-ownership
// of the destination lines, one third less DRAM traffic than a cached memcpy, which is what bounds
// searchValuesBatch's copy-out (several threads already share the work, c_abi.cu).
#include <cstddef>
#include <cstdint>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace sib {

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void copy_stream_avx2(char* d, const char* s, size_t n) {
    size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
    if (head > n) head = n;
    memcpy(d, s, head);
    d += head; s += head; n -= head;
    const size_t blocks = n / 128;
    for (size_t i = 0; i < blocks; ++i) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 96), e);
        s += 128; d += 128;
    }
    _mm_sfence();
    memcpy(d, s, n - blocks * 128);
}
#endif

// dst is write-once output: stream it past the caches when that is possible and worth it
void copy_to_output(void* dst, const void* src, size_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 4096) { copy_stream_avx2(static_cast<char*>(dst), static_cast<const char*>(src), n); return; }
#endif
    memcpy(dst, src, n);   // aarch64 hosts (Grace + B200): the libc copy
}

}  // namespace sib
