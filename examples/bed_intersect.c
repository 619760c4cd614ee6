/* bed_intersect.c -- the flow of the reference's examples/bed-intersect-si.rs (and of
 * test/bench.cpp:200-252) on libsuperintervals_b200: read two BED files, keep one interval set per
 * chrom, then for every chrom search and count all queries -- as ONE batch per chrom instead of a
 * query at a time. Prints, per chrom and in total, the number of overlaps found by searchValuesBatch
 * and by countOverlapsBatch (they must agree).
 *
 *   cc examples/bed_intersect.c -Iinclude -Lsuperintervals_b200 -lsuperintervals_b200 \
 *      -Wl,-rpath,$PWD/superintervals_b200 -o bed_intersect && ./bed_intersect ref.bed queries.bed
 *
 * BED ends are half-open; the index stores inclusive ends, so both sides use end - 1
 * (reference test/bench.cpp:210,220; the Rust example does the same in parse_bed_line). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "c_superintervals.h"
#include "superintervals_b200.h"

static char* slurp(const char* path, size_t* n) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* buf = (char*)malloc(sz > 0 ? (size_t)sz : 1);
    *n = fread(buf, 1, sz > 0 ? (size_t)sz : 0, f);
    fclose(f);
    return buf;
}

static int find_contig(const siBedTable* t, const char* name) {
    for (size_t k = 0; k < t->n_contigs; ++k)
        if (strcmp(t->names[k], name) == 0) return (int)k;
    return -1;
}

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s intervals.bed queries.bed\n", argv[0]); return 2; }
    size_t na = 0, nb = 0;
    char* ta = slurp(argv[1], &na);
    char* tb = slurp(argv[2], &nb);
    siBedTable A, Q;   /* tokenised on the device, grouped per chrom */
    if (siParseBed(ta, na, 1, -1, 1, &A) || siParseBed(tb, nb, 1, -1, 1, &Q)) {
        fprintf(stderr, "BED ingest failed: %s\n", si_b200_last_error_string());
        return 1;
    }
    free(ta);
    free(tb);
    unsigned long long total_found = 0, total_counted = 0;
    for (size_t c = 0; c < A.n_contigs; ++c) {
        const int qc = find_contig(&Q, A.names[c]);
        if (qc < 0) continue;
        const size_t a0 = A.contig_offsets[c], an = A.contig_offsets[c + 1] - a0;
        const size_t q0 = Q.contig_offsets[qc], qn = Q.contig_offsets[qc + 1] - q0;
        cSuperIntervals* si = createSuperIntervals();
        addIntervals(si, A.starts + a0, A.ends + a0, NULL, an);     /* payload = position in the file's chrom */
        indexSuperIntervals(si);
        size_t* offsets = (size_t*)malloc((qn + 1) * sizeof(size_t));
        size_t* counts = (size_t*)malloc((qn ? qn : 1) * sizeof(size_t));
        cIndexResult found = createIndexResult();
        searchValuesBatch(si, Q.starts + q0, Q.ends + q0, qn, offsets, &found);   /* CSR: query i -> found.data[offsets[i] .. offsets[i+1]) */
        countOverlapsBatch(si, Q.starts + q0, Q.ends + q0, qn, counts);
        unsigned long long counted = 0;
        for (size_t i = 0; i < qn; ++i) counted += counts[i];
        printf("%s\t%zu intervals\t%zu queries\t%zu found\t%llu counted\n", A.names[c], an, qn, found.size, counted);
        total_found += found.size;
        total_counted += counted;
        destroyIndexResult(&found);
        free(offsets);
        free(counts);
        destroySuperIntervals(si);
    }
    printf("total\t%llu found\t%llu counted\n", total_found, total_counted);
    siBedTableFree(&A);
    siBedTableFree(&Q);
    if (si_b200_last_error()) { fprintf(stderr, "error: %s\n", si_b200_last_error_string()); return 1; }
    return total_found == total_counted ? 0 : 1;
}
