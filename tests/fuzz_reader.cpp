// Sanitizer harness for the host-only half of the library: the FASTA reader (csrc/crf_fasta.h: plain / gzip / BGZF input,
// header scan, threaded compaction), the packer (csrc/crf_pack.h: planes, exotic list, mask runs) and the BED / TSV row writer
// (csrc/crf_rows.h, with FUZZ_READER_ROWS=n) compiled WITHOUT the CUDA
// runtime -- the three runtime calls they make (page-locked allocation) are stubbed to "no device" -- under AddressSanitizer +
// UBSan or ThreadSanitizer.  Each file given on the command line is opened with 1 and 7 threads, packed, and the planes are
// checked against the text base by base; the record table is printed so that the caller (tests/test_host_cpu.py) can compare
// it with its own reading of the file.
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -pthread -mavx2? (no: target attributes) -I csrc -I include fuzz_reader.cpp -lz
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <new>
#include <vector>

#include "crf.h"

// ---- what crf_api.cu provides to the two headers ----
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNoDevice = 100, cudaHostAllocDefault = 0 };
static cudaError_t cudaHostAlloc(void **, size_t, unsigned) { return cudaErrorNoDevice; }
static cudaError_t cudaFreeHost(void *) { return cudaSuccess; }
static cudaError_t cudaGetLastError() { return cudaSuccess; }
static thread_local char g_err[512];
static void set_err(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
extern "C" const char *crf_last_error(void) { return g_err; }

#include "crf_rows.h"
#include "crf_fasta.h"
#include "crf_pack.h"

static int check_file(const char *path, unsigned threads) {
    crf_fasta *fa = nullptr;
    int rc = crf_fasta_open(path, threads, 0, &fa);
    if (rc != CRF_OK) { printf("%s threads %u: status %d\n", path, threads, rc); return 0; }
    uint64_t n_rec = 0, total = 0, names_bytes = 0, n_exotic = 0;
    const uint8_t *bases; const uint64_t *offsets, *exotic; const char *names; const uint32_t *H, *L, *NM;
    crf_fasta_info(fa, &n_rec, &total, nullptr);
    crf_fasta_data(fa, &bases, &offsets, &names, &names_bytes);
    printf("%s threads %u: %llu records %llu bases;", path, threads, (unsigned long long)n_rec, (unsigned long long)total);
    const char *q = names;
    for (uint64_t r = 0; r < n_rec; ++r) {
        uint64_t h = 1469598103934665603ull;             // FNV-1a of the record's bases
        for (uint64_t p = offsets[r]; p < offsets[r + 1]; ++p) h = (h ^ bases[p]) * 1099511628211ull;
        printf(" %s:%llu:%016llx", q, (unsigned long long)(offsets[r + 1] - offsets[r]), (unsigned long long)h);
        q += strlen(q) + 1;
    }
    printf("\n");
    int bad = 0;
    if (total) {
        if (crf_fasta_packed(fa, threads, &H, &L, &NM, &exotic, &n_exotic) != CRF_OK) { printf("packed failed: %s\n", g_err); return 1; }
        uint64_t seen_exotic = 0;
        for (uint64_t p = 0; p < total && !bad; ++p) {
            uint8_t c = bases[p];
            if (c >= 'a' && c <= 'z') c -= 32;
            const uint32_t h = (H[p >> 5] >> (p & 31)) & 1, l = (L[p >> 5] >> (p & 31)) & 1, m = (NM[p >> 5] >> (p & 31)) & 1;
            const char *acgt = strchr("ACGT", c);
            if (acgt && c) { if (m || (uint32_t)(acgt - "ACGT") != (h << 1 | l)) bad = 1; }
            else {
                if (!m) bad = 1;
                if (c != 'N') { if (seen_exotic >= n_exotic || exotic[seen_exotic] != ((p << 8) | c)) bad = 1; ++seen_exotic; }
            }
        }
        if (seen_exotic != n_exotic) bad = 1;
        std::vector<uint64_t> runs(16);
        uint64_t n_runs = 0;
        int rr = crf_mask_runs(NM, total, threads, runs.data(), runs.size() / 2, &n_runs);
        if (rr == CRF_ERR_CAPACITY) { runs.resize(2 * n_runs); rr = crf_mask_runs(NM, total, threads, runs.data(), n_runs, &n_runs); }
        if (rr != CRF_OK) bad = 1;
        uint64_t masked = 0, in_runs = 0;
        for (uint64_t p = 0; p < total; ++p) masked += (NM[p >> 5] >> (p & 31)) & 1;
        for (uint64_t i = 0; i < n_runs; ++i) in_runs += runs[2 * i + 1] - runs[2 * i];
        if (masked != in_runs) bad = 1;
        if (bad) printf("PLANES DIFFER from the text (%s)\n", path);
    }
    // the row writer (csrc/crf_rows.h) on made-up rows of these records: `rows` of them, one thread below 131 072 rows and
    // several above; the caller compares the file with its own formatting of the same rows (row i: record i % n, start
    // (i * 7919) % (len - k + 1), k = 1 + i % 50 clipped to the record, end = start + 3k)
    const char *rows_env = getenv("FUZZ_READER_ROWS");
    if (rows_env && total && threads == 7) {
        std::vector<uint32_t> rec, st, en, kk;
        const uint64_t want = strtoull(rows_env, nullptr, 10);
        for (uint64_t i = 0; rec.size() < want && i < 4 * want + 64; ++i) {
            const uint64_t r = i % n_rec, len = offsets[r + 1] - offsets[r];
            if (!len) continue;
            const uint32_t k = (uint32_t)std::min<uint64_t>(1 + i % 50, len);
            rec.push_back((uint32_t)r); kk.push_back(k);
            st.push_back((uint32_t)((i * 7919) % (len - k + 1))); en.push_back(st.back() + 3 * k);
        }
        const std::string out = std::string(path) + ".bed";
        uint64_t bytes = 0;
        for (int tsv = 0; tsv < 2; ++tsv) {
            const std::string o = tsv ? std::string(path) + ".tsv" : out;
            if (crf_write_rows(o.c_str(), 0, tsv, names, bases, offsets, rec.data(), st.data(), en.data(), kk.data(), rec.size(), &bytes) != CRF_OK) {
                printf("crf_write_rows failed: %s\n", g_err);
                bad = 1;
            }
        }
        printf("%s: %zu rows written\n", path, rec.size());
    }
    crf_fasta_close(fa);
    return bad;
}

int main(int argc, char **argv) {
    int bad = 0;
    for (int i = 1; i < argc; ++i)
        for (unsigned threads : {1u, 7u}) bad |= check_file(argv[i], threads);
    return bad;
}
