/* multi_count.c -- every GPU of the node behind the C ABI (include/superintervals_b200.h section 4).
 *
 * Builds one interval set, replicates its index on all visible devices (siMultiBuildReplicated: one upload,
 * NCCL broadcast over NVLink), then answers ONE host batch split across the devices (siMultiCountBatch /
 * siMultiSearchValuesBatch: per-query counts all-gathered over NCCL, global CSR offsets on device) and checks
 * every answer against the single-device calls of the reference-shaped ABI (countOverlapsBatch /
 * searchValuesBatch). Prints one line per check and "multi ok".
 *
 *   cc examples/multi_count.c -Iinclude -Lsuperintervals_b200 -lsuperintervals_b200 \
 *      -Wl,-rpath,$PWD/superintervals_b200 -o multi_count && ./multi_count [n_devices] [intervals] [queries]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "c_superintervals.h"
#include "superintervals_b200.h"

static unsigned long long rng_state = 88172645463325252ull;
static unsigned rnd(void) {   /* xorshift64 */
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (unsigned)(rng_state >> 11);
}

int main(int argc, char** argv) {
    int n_dev = argc > 1 ? atoi(argv[1]) : 0;
    size_t n = argc > 2 ? (size_t)atoll(argv[2]) : 300000, nq = argc > 3 ? (size_t)atoll(argv[3]) : 1000003;
    const int axis = 8000000;
    int32_t *s = malloc(n * 4), *e = malloc(n * 4), *qs = malloc(nq * 4), *qe = malloc(nq * 4);
    for (size_t i = 0; i < n; ++i) { s[i] = (int32_t)(rnd() % axis); e[i] = s[i] + (int32_t)(150 + rnd() % 9000); }
    for (size_t i = 0; i < nq; ++i) { qs[i] = (int32_t)(rnd() % axis); qe[i] = qs[i] + (int32_t)(rnd() % 10000); }

    /* single device, the reference-shaped ABI */
    cSuperIntervals* one = createSuperIntervals();
    addIntervals(one, s, e, NULL, n);
    indexSuperIntervals(one);
    size_t* want = malloc(nq * sizeof(size_t));
    countOverlapsBatch(one, qs, qe, nq, want);
    size_t* want_off = malloc((nq + 1) * sizeof(size_t));
    cIndexResult want_vals = createIndexResult();
    searchValuesBatch(one, qs, qe, nq, want_off, &want_vals);
    if (si_b200_last_error()) { fprintf(stderr, "single-device path failed: %s\n", si_b200_last_error_string()); return 1; }

    siMulti* m = siMultiCreate(NULL, n_dev);
    if (!m) { fprintf(stderr, "siMultiCreate: %s\n", si_b200_last_error_string()); return 1; }
    n_dev = siMultiDeviceCount(m);
    if (siMultiBuildReplicated(m, s, e, NULL, n)) { fprintf(stderr, "build: %s\n", si_b200_last_error_string()); return 1; }
    uint32_t* got = malloc(nq * 4);
    if (siMultiCountBatch(m, qs, qe, nq, got)) { fprintf(stderr, "count: %s\n", si_b200_last_error_string()); return 1; }
    size_t bad = 0;
    unsigned long long hits = 0;
    for (size_t i = 0; i < nq; ++i) { bad += (size_t)got[i] != want[i]; hits += got[i]; }
    siMultiStats st;
    siMultiLastStats(m, &st);
    printf("count: devices %d queries %zu hits %llu mismatches %zu | h2d %.3f count %.3f gather %.3f d2h %.3f ms, nccl %d bytes %llu\n",
           n_dev, nq, hits, bad, st.ms_h2d, st.ms_count, st.ms_gather, st.ms_d2h, st.nccl_version, st.nccl_bytes);
    if (bad) return 1;
    if (n_dev > 1 && st.nccl_bytes == 0) { fprintf(stderr, "no NCCL traffic on %d devices\n", n_dev); return 1; }

    size_t* off = malloc((nq + 1) * sizeof(size_t));
    cIndexResult vals = createIndexResult();
    if (siMultiSearchValuesBatch(m, qs, qe, nq, off, &vals)) { fprintf(stderr, "search: %s\n", si_b200_last_error_string()); return 1; }
    bad = vals.size != want_vals.size;
    for (size_t i = 0; i <= nq && !bad; ++i) bad += off[i] != want_off[i];
    if (!bad) bad = memcmp(vals.data, want_vals.data, vals.size * sizeof(int32_t)) != 0;
    printf("search_values: total %zu (want %zu) %s\n", vals.size, want_vals.size, bad ? "MISMATCH" : "identical offsets and values");
    if (bad) return 1;

    destroyIndexResult(&vals); destroyIndexResult(&want_vals);
    siMultiDestroy(m);
    destroySuperIntervals(one);
    free(s); free(e); free(qs); free(qe); free(want); free(want_off); free(got); free(off);
    if (si_b200_last_error()) { fprintf(stderr, "latched error: %s\n", si_b200_last_error_string()); return 1; }
    printf("multi ok\n");
    return 0;
}
