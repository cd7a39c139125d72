// Sanitizer fuzz of the FASTA reader's gzip decoder (csrc/crf_inflate.h), several-threads path included: streams made with zlib at
// random levels / strategies / flush points from mixed data (DNA, soft-masked FASTA lines, noise, N runs, periodic text, long-range
// copies), a quarter of them with flipped bits, decoded with 2-64 KB chunks on 2-8 threads.  A valid stream must come back byte
// for byte; a corrupted one must be refused or (CRC-clean) identical; AddressSanitizer / UBSan watch every access meanwhile.
// Built and run by tests/test_host_cpu.py::test_gunzip_fuzz_under_sanitizers:  g++ -fsanitize=address,undefined -I csrc ... -lz
#include "crf_inflate.h"
#include <random>
// streams built with zlib at random settings from mixed data; decoded with small chunks on several threads; compared with the input
int main(int argc, char **argv) {
    unsigned seed = argc > 1 ? atoi(argv[1]) : 1, cases = argc > 2 ? atoi(argv[2]) : 50;
    std::mt19937_64 rng(seed);
    size_t n_par = 0, n_ok = 0;
    for (unsigned c = 0; c < cases; ++c) {
        // one to three gzip members back to back (`cat a.fa.gz b.fa.gz`): each may be decoded on several threads
        std::vector<uint8_t> data, gz;
        int level = 0, strat = 0;
        const int members = (rng() % 3 == 0) ? 2 + rng() % 2 : 1;
        for (int mem = 0; mem < members; ++mem) {
            const size_t data0 = data.size();
            int pieces = 1 + rng() % 6;
            for (int p = 0; p < pieces; ++p) {
                size_t n = 20000 + rng() % 600000;
                int kind = rng() % 6;
                size_t at = data.size();
                data.resize(at + n);
                if (kind == 0) for (size_t i = 0; i < n; ++i) data[at + i] = "ACGT"[rng() & 3];
                else if (kind == 1) for (size_t i = 0; i < n; ++i) data[at + i] = (i % 61 == 60) ? '\n' : "ACGTacgtNN"[rng() % 10];
                else if (kind == 2) for (size_t i = 0; i < n; ++i) data[at + i] = (uint8_t)rng();
                else if (kind == 3) memset(data.data() + at, 'N', n);
                else if (kind == 4) { size_t per = 1 + rng() % 50; for (size_t i = 0; i < n; ++i) data[at + i] = i < per ? "ACGT"[rng() & 3] : data[at + i - per]; }
                else { for (size_t i = 0; i < n; ++i) data[at + i] = (i > 40000 && (rng() % 100) < 95) ? data[at + i - 1 - rng() % 32000] : "ACGT\n"[rng() % 5]; }
            }
            z_stream zs; memset(&zs, 0, sizeof zs);
            level = rng() % 10; strat = (int[]){0, 0, 0, 0, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE, Z_FILTERED}[rng() % 8];
            deflateInit2(&zs, level, Z_DEFLATED, 31, 1 + rng() % 9, strat);
            const size_t n_mem = data.size() - data0, gz0 = gz.size();
            gz.resize(gz0 + deflateBound(&zs, n_mem) + 4096 + n_mem / 100);
            zs.next_in = data.data() + data0; zs.next_out = gz.data() + gz0; zs.avail_out = gz.size() - gz0;
            size_t fed = 0;
            while (fed < n_mem) {
                size_t step = std::min<size_t>(n_mem - fed, 1 + rng() % 300000);
                zs.avail_in = step;
                int fl = (int[]){Z_NO_FLUSH, Z_NO_FLUSH, Z_NO_FLUSH, Z_SYNC_FLUSH, Z_FULL_FLUSH, Z_BLOCK}[rng() % 6];
                deflate(&zs, fl);
                fed += step;
            }
            deflate(&zs, Z_FINISH);
            gz.resize(gz0 + zs.total_out);
            deflateEnd(&zs);
        }
        size_t glen = gz.size();
        bool corrupt = rng() % 4 == 0;
        if (corrupt) { for (int k = 0; k < 1 + (int)(rng() % 3); ++k) gz[10 + rng() % (glen - 18)] ^= 1u << (rng() % 8); }
        // exact-size copy so that ASan sees any read beyond the input
        std::vector<uint8_t> in(gz.begin(), gz.begin() + glen);
        crf_inflate::OutBuf out;
        size_t chunk = (size_t[]){2048, 4096, 16384, 65536}[rng() % 4];
        bool ok = crf_inflate::gunzip(in.data(), in.size(), out, 2 + rng() % 7, chunk);
        if (ok) { if (out.size != data.size() || memcmp(out.p, data.data(), data.size())) { if (!corrupt) { printf("MISMATCH case %u\n", c); return 1; } else if (out.size != data.size() || memcmp(out.p, data.data(), data.size())) { /* CRC collision would be astronomically unlikely */ printf("corrupt accepted with different bytes, case %u\n", c); return 1; } } ++n_ok; }
        else if (!corrupt) { printf("REJECTED a valid stream, case %u (level %d strat %d chunk %zu)\n", c, level, strat, chunk); return 1; }
    }
    printf("seed %u: %u cases, %zu accepted\n", seed, cases, n_ok);
}
