// Native FASTA ingest (host only): file -> one contiguous byte-per-base buffer + record table.
//
// Stands in for what the reference takes from pyfastx (perfect_repeat_finder.py:117-137: iterate the records,
// `.name` = first whitespace-delimited token of the header, `.seq` = the record's lines joined, case preserved).
// Plain or gzip (multi-member / bgzip too).  The buffer is laid out exactly as crf_seq_load_ascii wants it
// (records back to back + n+1 offsets), so a whole file goes to the GPU with one call and no Python-side copy.
//
// Included at the end of crf_api.cu (shares set_err()).
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "crf_inflate.h"

#include <atomic>
#include <exception>
#include <chrono>
#include <functional>
#include <thread>

struct crf_fasta {
    std::vector<uint64_t> offsets;      // n_records + 1
    std::vector<uint64_t> name_off;     // into names (NUL-separated, record order: what crf_write_rows takes)
    std::vector<char> names;
    uint8_t *bases = nullptr;
    uint64_t total = 0;
    bool pinned = false;
    // packed planes (crf_fasta_packed): H, L, NM of plane_words words each, one allocation
    uint32_t *planes = nullptr;
    uint64_t plane_words = 0;
    bool planes_pinned = false;
    bool want_pinned_planes = false;                     // crf_fasta_open(pinned & 3): page-lock the planes crf_fasta_packed makes
    std::vector<uint64_t> exotic;
};

namespace fasta_detail {

// The file's bytes: mapped when possible (pages are then faulted in by the worker threads, in parallel), else read.
struct FileBytes {
    const uint8_t *data = nullptr;
    size_t size = 0;
    void *map = nullptr;
    std::vector<uint8_t> owned;
    ~FileBytes() {
        if (map) munmap(map, size);
    }
};

static bool read_file(const char *path, std::vector<uint8_t> &out);

static bool open_bytes(const char *path, FileBytes &fb) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
        void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) {
            madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
            fb.map = m;
            fb.data = (const uint8_t *)m;
            fb.size = (size_t)st.st_size;
            close(fd);
            return true;
        }
    }
    close(fd);
    if (!read_file(path, fb.owned)) return false;
    fb.data = fb.owned.data();
    fb.size = fb.owned.size();
    return true;
}

static bool read_file(const char *path, std::vector<uint8_t> &out) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    bool ok = true;
    if (fseeko(f, 0, SEEK_END) == 0) {
        const off_t n = ftello(f);
        rewind(f);
        if (n > 0) {
            out.resize((size_t)n);
            ok = fread(out.data(), 1, (size_t)n, f) == (size_t)n;
        }
    } else {                                            // not seekable (a pipe)
        uint8_t buf[1 << 16];
        size_t got;
        while ((got = fread(buf, 1, sizeof buf, f)) > 0) out.insert(out.end(), buf, buf + got);
        ok = !ferror(f);
    }
    fclose(f);
    return ok;
}

// gzip -> text; concatenated members (bgzip) are decoded one after the other
static bool inflate_all(const FileBytes &in, std::vector<uint8_t> &out) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) return false;
    out.resize(std::max<size_t>(in.size * 4, 1 << 16));
    size_t in_pos = 0, out_pos = 0;
    bool ok = true;
    int rc = Z_OK;
    while (in_pos < in.size) {
        if (out_pos == out.size()) out.resize(out.size() + out.size() / 2);
        const size_t in_chunk = std::min<size_t>(in.size - in_pos, 1u << 30);
        const size_t out_chunk = std::min<size_t>(out.size() - out_pos, 1u << 30);
        zs.next_in = const_cast<Bytef *>(in.data + in_pos);
        zs.avail_in = (uInt)in_chunk;
        zs.next_out = out.data() + out_pos;
        zs.avail_out = (uInt)out_chunk;
        rc = inflate(&zs, Z_NO_FLUSH);
        in_pos += in_chunk - zs.avail_in;
        out_pos += out_chunk - zs.avail_out;
        if (rc == Z_STREAM_END) {
            if (in_pos < in.size && inflateReset(&zs) != Z_OK) { ok = false; break; }
        } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
            ok = false;
            break;
        } else if (rc == Z_BUF_ERROR && zs.avail_in == 0 && in_pos == in.size) {
            ok = false;                                 // truncated stream
            break;
        }
    }
    inflateEnd(&zs);
    out.resize(out_pos);
    return ok && rc == Z_STREAM_END;                    // anything else: the last member is truncated
}

struct Piece {                                          // a slice of one record's body text
    uint64_t begin, end;                                // text range
    uint64_t out;                                       // where its bases go
    uint64_t n_eol;                                     // '\n' + '\r' bytes inside
};

static void run_parallel(unsigned n_threads, size_t n_items, const std::function<void(size_t)> &fn) {
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (size_t i; (i = next.fetch_add(1)) < n_items;) fn(i);
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_threads && t < n_items; ++t) {
        try {
            pool.emplace_back(worker);
        } catch (const std::exception &) {              // no more threads to be had: go on with those we have
            break;
        }
    }
    worker();
    for (auto &th : pool) th.join();
}

// BGZF (bgzip, what `samtools faidx` wants and the reference's Hail pipeline ships: run_hail_batch_pipeline.py:112 takes
// the .gzi next to the FASTA): every gzip member is a self-contained block of <= 64 KiB that says its own compressed size
// in a 'BC' extra field, so the blocks are found by hopping over the headers and inflated on all threads.
// Returns false if the file is not BGZF from the first to the last byte (the caller then inflates it serially);
// *bad is set if it is BGZF but a block is corrupt.
// Each block is decoded by the reader's own DEFLATE decoder (crf_inflate.h, ~2 x zlib on FASTA text) into a scratch buffer --
// its copies may write a few bytes beyond a match, which in the shared output would be the next block's -- and copied to its
// place once length and CRC-32 agree; a block it does not take (or CRF_GUNZIP_ZLIB=1) goes through zlib, which then decides.
// The output is a mapping of its own (no zero-fill before the blocks arrive, 2 MB pages).
static bool inflate_bgzf(const FileBytes &in, unsigned n_threads, crf_inflate::OutBuf &out, bool *bad) {
    struct Block { uint64_t cdata, clen, out; uint32_t isize, crc; };
    std::vector<Block> blocks;
    const uint8_t *d = in.data;
    const uint64_t n = in.size;
    uint64_t pos = 0, total = 0;
    while (pos < n) {
        if (n - pos < 18 || d[pos] != 0x1f || d[pos + 1] != 0x8b || d[pos + 2] != 8 || !(d[pos + 3] & 4)) return false;
        const uint32_t xlen = d[pos + 10] | (d[pos + 11] << 8);
        if (n - pos < 12ull + xlen + 8) return false;
        uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) {           // extra subfields: SI1 SI2 SLEN(2) data
            const uint8_t *f = d + pos + 12 + x;
            const uint32_t slen = f[2] | (f[3] << 8);
            if (f[0] == 'B' && f[1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = (f[4] | (f[5] << 8)) + 1u;
            x += 4 + slen;
        }
        if (bsize < 12 + xlen + 8 || pos + bsize > n) return false;
        if (d[pos + 3] & ~4) return false;                // other header fields (name, comment, crc): not what bgzip writes
        Block b;
        b.cdata = pos + 12 + xlen;
        b.clen = bsize - 12 - xlen - 8;
        const uint8_t *tail = d + pos + bsize - 8;
        b.crc = tail[0] | (tail[1] << 8) | (tail[2] << 16) | ((uint32_t)tail[3] << 24);
        b.isize = tail[4] | (tail[5] << 8) | (tail[6] << 16) | ((uint32_t)tail[7] << 24);
        if (b.isize > 65536) return false;
        b.out = total;
        total += b.isize;
        blocks.push_back(b);
        pos += bsize;
    }
    if (!out.reserve(total + 64)) throw std::bad_alloc();
    out.size = total;
    uint8_t *const dst = out.p;
    const bool own_decoder = getenv("CRF_GUNZIP_ZLIB") == nullptr;
    std::atomic<bool> failed{false}, no_memory{false};
    const size_t GROUP = 128;                             // blocks per work item (<= 8 MiB of text)
    run_parallel(n_threads, (blocks.size() + GROUP - 1) / GROUP, [&](size_t g) {
      try {
        crf_inflate::OutBuf scratch;
        crf_inflate::Tables tables;
        z_stream zs;
        bool zs_ready = false;
        for (size_t i = g * GROUP; i < std::min(blocks.size(), (g + 1) * GROUP) && !failed; ++i) {
            const Block &b = blocks[i];
            if (b.isize == 0) continue;                   // the empty end-of-file block
            if (own_decoder) {
                crf_inflate::Decoder dec;
                dec.base = dec.in = d + b.cdata;
                dec.in_end = dec.in + b.clen;
                dec.out = &scratch;
                dec.op = 0;
                if (dec.run(tables) == crf_inflate::Decoder::RUN_FINAL && dec.op == b.isize &&
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), scratch.p, b.isize) == b.crc) {
                    memcpy(dst + b.out, scratch.p, b.isize);
                    continue;
                }
            }
            if (!zs_ready) {
                memset(&zs, 0, sizeof zs);
                if (inflateInit2(&zs, -15) != Z_OK) { failed = true; return; }
                zs_ready = true;
            }
            zs.next_in = const_cast<Bytef *>(d + b.cdata);
            zs.avail_in = (uInt)b.clen;
            zs.next_out = dst + b.out;
            zs.avail_out = b.isize;
            const int rc = inflate(&zs, Z_FINISH);
            if (rc != Z_STREAM_END || zs.avail_out != 0 ||
                (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst + b.out, b.isize) != b.crc)
                failed = true;
            inflateReset(&zs);
        }
        if (zs_ready) inflateEnd(&zs);
      } catch (const std::bad_alloc &) { no_memory = true; failed = true; }      // (nothing leaves a worker thread)
    });
    if (no_memory) throw std::bad_alloc();
    *bad = failed;
    return true;
}

}  // namespace fasta_detail

static int fasta_open_impl(const char *path, uint32_t n_threads, int pinned, crf_fasta **out);
extern "C" int crf_fasta_close(crf_fasta *fa);

extern "C" int crf_fasta_open(const char *path, uint32_t n_threads, int pinned, crf_fasta **out) {
    if (!path || !out) { set_err("crf_fasta_open: null argument"); return CRF_ERR_ARG; }
    *out = nullptr;
    try {                                               // no exception crosses the C boundary
        return fasta_open_impl(path, n_threads, pinned, out);
    } catch (const std::bad_alloc &) {
        set_err("crf_fasta_open: out of host memory reading %s", path);
        return CRF_ERR_NOMEM;
    } catch (const std::exception &ex) {
        set_err("crf_fasta_open: %s", ex.what());
        return CRF_ERR_ARG;
    }
}

static int fasta_open_impl(const char *path, uint32_t n_threads, int pinned, crf_fasta **out) {
    using namespace fasta_detail;
    const bool trace = getenv("CRF_FASTA_TRACE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[crf_fasta] %-10s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    FileBytes file;
    std::vector<uint8_t> plain;
    crf_inflate::OutBuf inflated;
    if (!open_bytes(path, file)) { set_err("crf_fasta_open: cannot read %s", path); return CRF_ERR_ARG; }
    const uint8_t *t = file.data;
    uint64_t n = file.size;
    if (n_threads == 0) n_threads = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (n >= 2 && t[0] == 0x1f && t[1] == 0x8b) {
        bool bad = false;
        const bool bgzf = inflate_bgzf(file, n_threads, inflated, &bad);
        if (bgzf && bad) { set_err("crf_fasta_open: %s has a corrupt BGZF block", path); return CRF_ERR_ARG; }
        // a plain gzip stream: the reader's own decoder (crf_inflate.h); whatever it does not take goes to zlib, which decides
        const bool own = !bgzf && getenv("CRF_GUNZIP_ZLIB") == nullptr && crf_inflate::gunzip(file.data, file.size, inflated, n_threads);
        if (!bgzf && !own && !inflate_all(file, plain)) { set_err("crf_fasta_open: %s is not a valid gzip stream", path); return CRF_ERR_ARG; }
        t = bgzf || own ? inflated.p : plain.data();
        n = bgzf || own ? inflated.size : plain.size();
        lap(bgzf ? "bgzf" : own ? "gunzip" : "gzip/zlib");
    }
    lap("read");

    struct Guard {                                      // frees the handle unless it is handed to the caller
        crf_fasta *p;
        ~Guard() { if (p) crf_fasta_close(p); }
    } guard{new crf_fasta};
    crf_fasta *fa = guard.p;

    // headers: a '>' at the start of a line
    const uint64_t PIECE = 8u << 20;
    std::vector<std::vector<uint64_t>> hdr_in((n + PIECE - 1) / PIECE);
    run_parallel(n_threads, hdr_in.size(), [&](size_t i) {
        const uint64_t hi = std::min<uint64_t>(n, (i + 1) * PIECE);
        for (uint64_t p = i * PIECE; p < hi;) {
            const uint8_t *q = (const uint8_t *)memchr(t + p, '>', hi - p);
            if (!q) break;
            const uint64_t at = q - t;
            if (at == 0 || t[at - 1] == '\n' || t[at - 1] == '\r') hdr_in[i].push_back(at);
            p = at + 1;
        }
    });
    std::vector<uint64_t> hdr;
    for (const auto &v : hdr_in) hdr.insert(hdr.end(), v.begin(), v.end());
    const uint64_t n_rec = hdr.size();
    std::vector<uint64_t> body_lo(n_rec), body_hi(n_rec);
    for (uint64_t r = 0; r < n_rec; ++r) {
        const uint64_t stop = r + 1 < n_rec ? hdr[r + 1] : n;
        const uint8_t *eol = (const uint8_t *)memchr(t + hdr[r], '\n', stop - hdr[r]);
        uint64_t h1 = eol ? (uint64_t)(eol - t) : stop;            // header text = (hdr, h1)
        body_lo[r] = eol ? h1 + 1 : stop;
        body_hi[r] = stop;
        uint64_t a = hdr[r] + 1;
        auto is_space = [](uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; };
        while (a < h1 && is_space(t[a])) ++a;
        uint64_t b = a;
        while (b < h1 && !is_space(t[b])) ++b;
        fa->name_off.push_back(fa->names.size());
        fa->names.insert(fa->names.end(), t + a, t + b);
        fa->names.push_back('\0');
    }

    lap("headers");
    // bodies in pieces of <= 8 MiB of text: count line ends, prefix-sum, compact
    std::vector<Piece> pieces;
    std::vector<uint64_t> first_piece(n_rec + 1);
    for (uint64_t r = 0; r < n_rec; ++r) {
        first_piece[r] = pieces.size();
        for (uint64_t p = body_lo[r]; p < body_hi[r]; p += PIECE) pieces.push_back({p, std::min(body_hi[r], p + PIECE), 0, 0});
    }
    first_piece[n_rec] = pieces.size();
    run_parallel(n_threads, pieces.size(), [&](size_t i) {
        Piece &pc = pieces[i];
        uint64_t c = 0;
        for (uint64_t p = pc.begin; p < pc.end; ++p) c += (t[p] == '\n') | (t[p] == '\r');
        pc.n_eol = c;
    });
    lap("count");
    fa->offsets.assign(n_rec + 1, 0);
    uint64_t total = 0;
    for (uint64_t r = 0; r < n_rec; ++r) {
        fa->offsets[r] = total;
        for (uint64_t i = first_piece[r]; i < first_piece[r + 1]; ++i) {
            pieces[i].out = total;
            total += pieces[i].end - pieces[i].begin - pieces[i].n_eol;
        }
    }
    fa->offsets[n_rec] = total;
    fa->total = total;

    const size_t alloc = std::max<uint64_t>(total, 1);
    fa->want_pinned_planes = (pinned & 3) != 0;
    if (pinned & 1) {
        void *p = nullptr;
        if (cudaHostAlloc(&p, alloc, cudaHostAllocDefault) == cudaSuccess) {
            fa->bases = (uint8_t *)p;
            fa->pinned = true;
        } else {
            cudaGetLastError();                         // no device / no pinned memory left: plain pages
        }
    }
    if (!fa->bases) fa->bases = (uint8_t *)malloc(alloc);
    if (!fa->bases) { set_err("crf_fasta_open: out of host memory (%llu bases)", (unsigned long long)total); return CRF_ERR_NOMEM; }
#if defined(__linux__) && defined(MADV_HUGEPAGE)
    if (!fa->pinned && alloc >= ((size_t)8 << 20)) {    // gigabytes touched for the first time by the compaction below: 2 MB pages
        const uintptr_t lo = ((uintptr_t)fa->bases + 4095) & ~(uintptr_t)4095, hi = ((uintptr_t)fa->bases + alloc) & ~(uintptr_t)4095;
        if (hi > lo) madvise((void *)lo, hi - lo, MADV_HUGEPAGE);              // (advice only; the block is never realloc'ed)
    }
#endif

    lap("alloc");
    uint8_t *dst = fa->bases;
    run_parallel(n_threads, pieces.size(), [&](size_t i) {
        const Piece &pc = pieces[i];
        uint8_t *o = dst + pc.out;
        uint64_t p = pc.begin;
        const uint64_t K1 = 0x0101010101010101ull, K80 = 0x8080808080808080ull;
        while (p + 8 <= pc.end) {                       // 8 bytes at a time; a word holding a line end goes byte by byte
            uint64_t w;
            memcpy(&w, t + p, 8);
            const uint64_t a = w ^ (K1 * '\n'), b = w ^ (K1 * '\r');
            if ((((a - K1) & ~a) | ((b - K1) & ~b)) & K80) {
                for (int j = 0; j < 8; ++j) {
                    const uint8_t c = t[p + j];
                    if (c != '\n' && c != '\r') *o++ = c;      // never touch a byte beyond this piece's output
                }
            } else {
                memcpy(o, &w, 8);
                o += 8;
            }
            p += 8;
        }
        for (; p < pc.end; ++p) {
            const uint8_t c = t[p];
            if (c != '\n' && c != '\r') *o++ = c;
        }
    });
    lap("compact");
    guard.p = nullptr;
    *out = fa;
    return CRF_OK;
}

// gzip -> bytes with the reader's decoder (use_zlib == 0: falls back to zlib like the reader does; 2: no fall-back,
// CRF_ERR_UNSUPPORTED instead) or with zlib only (1).
// CRF_ERR_CAPACITY with *out_n = the size needed when `cap` is too small.
extern "C" int crf_gunzip(const uint8_t *gz, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *out_n, int use_zlib) {
    if (!gz || !out_n || (cap && !out)) { set_err("crf_gunzip: null argument"); return CRF_ERR_ARG; }
    try {
        using namespace fasta_detail;
        crf_inflate::OutBuf fast;
        std::vector<uint8_t> slow;
        const uint8_t *p = nullptr;
        uint64_t got = 0;
        if (use_zlib != 1 && crf_inflate::gunzip(gz, (size_t)n, fast, 4)) {
            p = fast.p; got = fast.size;
        } else if (use_zlib == 2) {
            set_err("crf_gunzip: the reader's decoder does not take this stream");
            return CRF_ERR_UNSUPPORTED;
        } else {
            FileBytes in;
            in.data = gz; in.size = (size_t)n;
            if (!inflate_all(in, slow)) { set_err("crf_gunzip: not a valid gzip stream"); return CRF_ERR_ARG; }
            p = slow.data(); got = slow.size();
        }
        *out_n = got;
        if (got > cap) { set_err("crf_gunzip: capacity %llu too small for %llu bytes", (unsigned long long)cap, (unsigned long long)got); return CRF_ERR_CAPACITY; }
        if (got) memcpy(out, p, (size_t)got);
        return CRF_OK;
    } catch (const std::bad_alloc &) {
        set_err("crf_gunzip: out of host memory");
        return CRF_ERR_NOMEM;
    }
}

extern "C" int crf_fasta_info(const crf_fasta *fa, uint64_t *n_records, uint64_t *total_bases, int *pinned) {
    if (!fa) { set_err("null fasta handle"); return CRF_ERR_ARG; }
    if (n_records) *n_records = fa->offsets.size() - 1;
    if (total_bases) *total_bases = fa->total;
    if (pinned) *pinned = fa->pinned ? 1 : 0;
    return CRF_OK;
}

extern "C" int crf_fasta_data(const crf_fasta *fa, const uint8_t **bases, const uint64_t **offsets, const char **names,
                              uint64_t *names_bytes) {
    if (!fa) { set_err("null fasta handle"); return CRF_ERR_ARG; }
    if (bases) *bases = fa->bases;
    if (offsets) *offsets = fa->offsets.data();
    if (names) *names = fa->names.data();
    if (names_bytes) *names_bytes = fa->names.size();
    return CRF_OK;
}

extern "C" int crf_fasta_close(crf_fasta *fa) {
    if (!fa) return CRF_OK;
    if (fa->pinned) cudaFreeHost(fa->bases);
    else free(fa->bases);
    if (fa->planes_pinned) cudaFreeHost(fa->planes);
    else free(fa->planes);
    delete fa;
    return CRF_OK;
}
