/* FASTA (plain / gzip) -> BED of perfect tandem repeats through the C ABI alone (include/crf.h): no Python, no torch.
 * What the reference does in main() for a FASTA input (perfect_repeat_finder.py:117-151: pyfastx records -> detect_repeats()
 * per record -> BED rows), here for every record of the file in one load and one scan.
 *
 *   gcc -std=c99 -I include examples/fasta_to_bed.c -o fasta_to_bed \
 *       -L colab-repeat-finder_b200/crf_b200 -l:libcrf.so -Wl,-rpath,$PWD/colab-repeat-finder_b200/crf_b200
 *   ./fasta_to_bed genome.fa.gz genome.bed [min_motif_size max_motif_size min_repeats min_span]     (defaults 1 50 3 9)
 *
 * tests/test_cabi_cpu.py builds it with -pedantic -Werror and runs it (without a GPU: up to the context, which must fail
 * with a message, not crash). */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "crf.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != CRF_OK) {                                                         \
            fprintf(stderr, "%s: status %d: %s\n", #call, rc_, crf_last_error());    \
            status = 1;                                                              \
            goto done;                                                               \
        }                                                                            \
    } while (0)

int main(int argc, char **argv) {
    crf_fasta *fa = NULL;
    crf_ctx *ctx = NULL;
    crf_seq *seq = NULL;
    uint32_t *record = NULL, *start = NULL, *end = NULL, *motif_size = NULL;
    int status = 0;
    if (argc != 3 && argc != 7) {
        fprintf(stderr, "usage: %s in.fa[.gz] out.bed [min_motif_size max_motif_size min_repeats min_span]\n", argv[0]);
        return 2;
    }
    crf_scan_params params;
    memset(&params, 0, sizeof params);               /* knobs 0 = library defaults */
    params.min_motif_size = argc == 7 ? (uint32_t)atoi(argv[3]) : 1;
    params.max_motif_size = argc == 7 ? (uint32_t)atoi(argv[4]) : 50;
    params.min_repeats = argc == 7 ? (uint32_t)atoi(argv[5]) : 3;
    params.min_span = argc == 7 ? (uint32_t)atoi(argv[6]) : 9;

    /* host only: read, drop line ends, pack to 2-bit planes + mask (page-locked planes when a device is there) */
    CHECK(crf_fasta_open(argv[1], 0, 2, &fa));
    uint64_t n_records = 0, total = 0, names_bytes = 0, n_exotic = 0;
    const uint8_t *bases = NULL;
    const uint64_t *offsets = NULL, *exotic = NULL;
    const char *names = NULL;
    const uint32_t *H = NULL, *L = NULL, *NM = NULL;
    CHECK(crf_fasta_info(fa, &n_records, &total, NULL));
    CHECK(crf_fasta_data(fa, &bases, &offsets, &names, &names_bytes));
    printf("%" PRIu64 " records, %" PRIu64 " bp\n", n_records, total);
    if (n_records > 0xFFFFFFFFu || total + n_records * params.max_motif_size > crf_load_limit(params.max_motif_size)) {
        fprintf(stderr, "more than one load can hold: split the records over several loads (crf_load_limit)\n");
        status = 1;
        goto done;
    }
    if (total) CHECK(crf_fasta_packed(fa, 0, &H, &L, &NM, &exotic, &n_exotic));

    uint64_t n_rows = 0, bytes = 0;
    if (total) {
        CHECK(crf_ctx_create(0, &ctx));
        CHECK(crf_seq_load_packed(ctx, H, L, NM, exotic, n_exotic, offsets, (uint32_t)n_records, params.max_motif_size, 0, &seq));
        CHECK(crf_scan(seq, &params, &n_rows));      /* rows stay in HBM, sorted by (record, start, end) */
        const size_t n = n_rows ? (size_t)n_rows : 1;
        record = malloc(n * 4); start = malloc(n * 4); end = malloc(n * 4); motif_size = malloc(n * 4);
        if (!record || !start || !end || !motif_size) { fprintf(stderr, "out of memory\n"); status = 1; goto done; }
        CHECK(crf_fetch(seq, record, start, end, motif_size, n_rows, 0));
        crf_scan_stats_t stats;
        CHECK(crf_scan_stats(seq, &stats));
        printf("scan: %.3f ms on the device (%u kernel launches)\n", stats.scan_ms, stats.launches);
    }
    /* host only: chrom \t start \t end \t motif, the motif read back from the text */
    CHECK(crf_write_rows(argv[2], 0, 0, names, bases, offsets, record, start, end, motif_size, n_rows, &bytes));
    printf("Found %" PRIu64 " repeats\nWrote results to %s (%" PRIu64 " bytes)\n", n_rows, argv[2], bytes);

done:
    free(record); free(start); free(end); free(motif_size);
    if (seq) crf_seq_destroy(seq);
    if (ctx) crf_ctx_destroy(ctx);
    if (fa) crf_fasta_close(fa);
    return status;
}
