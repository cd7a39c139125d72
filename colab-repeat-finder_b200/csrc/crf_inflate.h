// gzip -> bytes for the FASTA reader (host only; part of row f1 of DESIGN.md: the reference reads `ref.fa.gz` through
// pyfastx, perfect_repeat_finder.py:117).
//
// A plain gzip file is ONE DEFLATE stream (RFC 1951/1952) with no index of block starts, so how fast it is decoded
// decides how long a `.fa.gz` run takes: zlib's inflate does ~110 MB/s of FASTA text (literal-heavy, 2-3 bit codes, one
// symbol per loop), 25 s for a human genome that is scanned in milliseconds.  This decoder is written for that data:
//   * 64-bit bit buffer, refilled branch-free with one unaligned load (8 input bytes are always there in the main loop);
//   * one 11-bit table look-up per literal/length symbol (sub-tables behind it for the rare longer codes), literals
//     decoded back to back without refilling in between; 8-bit table for the distance symbols;
//   * the whole output is one contiguous buffer (realloc/mremap-grown), so a match is a plain copy from earlier output --
//     no 32 KB window to maintain -- done 16 bytes at a time when the distance allows;
//   * a big stream is decoded on several threads all the same (gunzip_parallel below): chunks find a block start by trial,
//     decode into 16-bit symbols that stand for "byte j of the 32 KB I did not see", and are resolved once the chunk in
//     front is known;
//   * the CRC-32 of every member is checked afterwards on several threads (crc32_combine), as is ISIZE.
// Anything unexpected (a code set zlib would reject, a bad CRC, trailing bytes that are no gzip member) makes gunzip()
// return false and the caller falls back to zlib, which then gives the verdict.  tests/test_host_cpu.py compares both with
// Python's zlib on streams of every block type.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#if defined(__linux__)
#include <sys/mman.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

namespace crf_inflate {

// Output buffer that grows.  On Linux its own anonymous mapping, grown with mremap (pages move, bytes are not copied) and
// advised MADV_HUGEPAGE as a whole: first touch of a big buffer otherwise costs as much as decoding into it (488 MB: 409 ms
// on one thread in 4 KB pages).  (malloc + madvise on a part of the block splits the mapping, after which realloc can no
// longer mremap and copies instead: 160 ms for one growth step of a 155 MB buffer.)
struct OutBuf {
    uint8_t *p = nullptr;
    size_t size = 0, cap = 0;
    OutBuf() {}
    OutBuf(const OutBuf &) = delete;
    OutBuf &operator=(const OutBuf &) = delete;
#if defined(__linux__)
    ~OutBuf() { if (p) munmap(p, cap); }
    bool reserve(size_t need) {
        if (need <= cap) return true;
        need = (need + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
        void *q = p ? mremap(p, cap, need, MREMAP_MAYMOVE) : mmap(nullptr, need, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (q == MAP_FAILED) return false;
        p = (uint8_t *)q;
        cap = need;
#ifdef MADV_HUGEPAGE
        madvise(p, cap, MADV_HUGEPAGE);                  // (advice only)
#endif
        return true;
    }
#else
    ~OutBuf() { free(p); }
    bool reserve(size_t need) {
        if (need <= cap) return true;
        uint8_t *q = (uint8_t *)realloc(p, need);
        if (!q) return false;
        p = q;
        cap = need;
        return true;
    }
#endif
};

enum : uint32_t { K_INVALID = 0, K_LITERAL = 1, K_LENGTH = 2, K_EOB = 3, K_SUB = 4 };
// table entry: value (literal byte / length or distance base / sub-table start) << 16 | kind << 12 | extra bits << 8 |
// bits this entry consumes (code bits, plus the extra bits of a length / distance)
static inline uint32_t entry(uint32_t value, uint32_t kind, uint32_t extra) { return (value << 16) | (kind << 12) | (extra << 8); }
static inline uint32_t e_bits(uint32_t e) { return e & 0xFFu; }
static inline uint32_t e_extra(uint32_t e) { return (e >> 8) & 0xFu; }
static inline uint32_t e_kind(uint32_t e) { return (e >> 12) & 0xFu; }
static inline uint32_t e_value(uint32_t e) { return e >> 16; }

constexpr int LITLEN_BITS = 11, DIST_BITS = 8, PRE_BITS = 7;
#ifndef CRF_INFLATE_MULTI_BITS
#define CRF_INFLATE_MULTI_BITS 10
#endif
constexpr int MULTI_BITS = CRF_INFLATE_MULTI_BITS;       // look-up width of the several-literals-at-once table
constexpr size_t LITLEN_CAP = (1u << LITLEN_BITS) + 288 * 16, DIST_CAP = (1u << DIST_BITS) + 32 * 128;

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                       4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

static inline uint32_t litlen_entry(int s) {
    if (s < 256) return entry((uint32_t)s, K_LITERAL, 0);
    if (s == 256) return entry(0, K_EOB, 0);
    if (s < 286) return entry(LEN_BASE[s - 257], K_LENGTH, LEN_EXTRA[s - 257]);
    return entry(0, K_INVALID, 0);                       // 286, 287: have codes in the fixed set, may not occur
}
static inline uint32_t dist_entry(int s) {
    if (s < 30) return entry(DIST_BASE[s], K_LENGTH, DIST_EXTRA[s]);
    return entry(0, K_INVALID, 0);
}
static inline uint32_t pre_entry(int s) { return entry((uint32_t)s, K_LITERAL, 0); }

static inline uint32_t bit_reverse(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

// Canonical Huffman code (RFC 1951 3.2.2) -> look-up table indexed by the next `tablebits` input bits (codes are packed
// starting at the least significant bit, hence bit-reversed), longer codes in sub-tables behind a K_SUB entry.  False for
// an over-subscribed set and for an incomplete one other than "a single code of one bit" (what zlib accepts).
template <class F>
static bool build_table(const uint8_t *lens, int nsyms, int tablebits, uint32_t *table, size_t cap, F sym_entry) {
    int count[16] = {0};
    for (int s = 0; s < nsyms; ++s) ++count[lens[s]];
    count[0] = 0;
    int left = 1, total = 0, max_len = 0;
    for (int len = 1; len <= 15; ++len) {
        left = (left << 1) - count[len];
        if (left < 0) return false;
        total += count[len];
        if (count[len]) max_len = len;
    }
    const uint32_t primary = 1u << tablebits;
    for (uint32_t i = 0; i < primary; ++i) table[i] = 0;
    if (total == 0) return true;                         // no code at all (a block without matches has no distance code)
    if (left > 0 && max_len != 1) return false;
    uint32_t next_code[16], code = 0;
    for (int len = 1; len <= 15; ++len) {
        code = (code + (uint32_t)count[len - 1]) << 1;
        next_code[len] = code;
    }
    uint16_t code_of[288];
    for (int s = 0; s < nsyms; ++s)
        if (lens[s]) code_of[s] = (uint16_t)next_code[lens[s]]++;
    size_t next = primary;
    if (max_len > tablebits) {                           // size the sub-tables: the longest code behind each prefix
        uint8_t sub_bits[1u << LITLEN_BITS] = {0};
        for (int s = 0; s < nsyms; ++s)
            if (lens[s] > tablebits) {
                const uint32_t prefix = bit_reverse(code_of[s], lens[s]) & (primary - 1);
                sub_bits[prefix] = std::max<uint8_t>(sub_bits[prefix], (uint8_t)(lens[s] - tablebits));
            }
        for (uint32_t prefix = 0; prefix < primary; ++prefix)
            if (sub_bits[prefix]) {
                const size_t n = (size_t)1 << sub_bits[prefix];
                if (next + n > cap || next > 0xFFFFu) return false;
                table[prefix] = entry((uint32_t)next, K_SUB, sub_bits[prefix]) | (uint32_t)tablebits;
                for (size_t i = 0; i < n; ++i) table[next + i] = 0;
                next += n;
            }
    }
    for (int s = 0; s < nsyms; ++s) {
        const int len = lens[s];
        if (!len) continue;
        const uint32_t e = sym_entry(s), rev = bit_reverse(code_of[s], len);
        const uint32_t xb = e_kind(e) == K_LENGTH ? e_extra(e) : 0;   // the entry says how far to shift: code + extra bits
        if (len <= tablebits) {
            for (uint32_t i = rev; i < primary; i += 1u << len) table[i] = e | ((uint32_t)len + xb);
        } else {
            const uint32_t t = table[rev & (primary - 1)], start = e_value(t), bits = e_extra(t), sl = (uint32_t)(len - tablebits);
            for (uint32_t i = rev >> tablebits; i < (1u << bits); i += 1u << sl) table[start + i] = e | (sl + xb);
        }
    }
    return true;
}

// FASTA text is nearly all literals with codes of 2-3 bits, and one look-up per literal is a chain of dependent
// load -> shift -> load.  This second table answers "which literals do the next MULTI_BITS bits hold": up to four bytes,
// how many, and how many bits they take -- derived from the literal/length table of the block.  n = 0: the next symbol is
// no literal (or a literal with a long code): take the one-symbol path.
static void build_multi(const uint32_t *lt, uint64_t *mt) {
    for (uint32_t i = 0; i < (1u << MULTI_BITS); ++i) {
        uint32_t left = MULTI_BITS, idx = i, n = 0, total = 0, bytes = 0;
        while (n < 4) {
            const uint32_t e = lt[idx];                  // (idx has `left` significant bits; a code that fits in them is decided)
            if (e_kind(e) != K_LITERAL || e_bits(e) > left) break;
            bytes |= e_value(e) << (8 * n);
            ++n; total += e_bits(e); left -= e_bits(e); idx >>= e_bits(e);
        }
        mt[i] = bytes | ((uint64_t)total << 32) | ((uint64_t)n << 40);
    }
}

struct Tables {                                          // per decoder: the tables of the current block, and the fixed code's
    std::vector<uint32_t> lt, dt, flt, fdt;
    std::vector<uint64_t> mt, fmt;
    Tables() : lt(LITLEN_CAP), dt(DIST_CAP), mt((size_t)1 << MULTI_BITS) {}
    bool fixed() {
        if (!flt.empty()) return true;
        uint8_t l[288], d[32];                           // the fixed code of RFC 1951 3.2.6
        for (int s = 0; s < 288; ++s) l[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
        for (int s = 0; s < 32; ++s) d[s] = 5;
        flt.resize(LITLEN_CAP); fdt.resize(DIST_CAP); fmt.resize((size_t)1 << MULTI_BITS);
        if (!build_table(l, 288, LITLEN_BITS, flt.data(), LITLEN_CAP, litlen_entry) ||
            !build_table(d, 32, DIST_BITS, fdt.data(), DIST_CAP, dist_entry)) return false;
        build_multi(flt.data(), fmt.data());
        return true;
    }
};

// T = uint8_t: bytes.  T = uint16_t: "marker" symbols for a decoder that starts in the middle of a stream (gunzip_parallel):
// 0..255 = a byte, 256 + j = whatever byte j of the unknown 32 KB before the starting point is; the buffer then begins with
// those 32 768 markers, so that matches reaching back before the start copy them like anything else.
template <class T>
struct DecoderT {
    const uint8_t *base = nullptr, *in = nullptr, *in_end = nullptr;
    uint64_t bb = 0;                                     // bit buffer: the next input bits, least significant first
    int bc = 0;                                          // valid bits in it
    OutBuf *out = nullptr;                               // (capacity in bytes; `op` counts elements of T)
    size_t op = 0;
    size_t op_limit = ~(size_t)0;                        // give up beyond this many elements (marker chunks: bounded memory)
    bool hit_limit = false;                              // ... and say that this, not bad data, was the reason

    inline uint64_t bitpos() const { return 8 * (uint64_t)(in - base) - (uint64_t)bc; }
    inline void seek(uint64_t bit) {
        in = base + (bit >> 3); bb = 0; bc = 0;
        uint32_t v;
        if (bit & 7) bits((int)(bit & 7), &v);
    }
    inline void refill() {
        if (in_end - in >= 8) {
            uint64_t w;
            memcpy(&w, in, 8);                           // (little-endian host: x86-64)
            bb |= w << bc;
            in += (63 - bc) >> 3;
            bc |= 56;
        } else {
            while (bc <= 56 && in < in_end) { bb |= (uint64_t)*in++ << bc; bc += 8; }
        }
    }
    inline bool bits(int n, uint32_t *v) {               // header fields: n <= 16
        if (bc < n) refill();
        if (bc < n) return false;
        *v = (uint32_t)(bb & ((1ull << n) - 1));
        bb >>= n;
        bc -= n;
        return true;
    }
    inline void to_byte_boundary() {                     // drop the rest of the current byte, hand whole bytes back
        const int drop = bc & 7;
        bb >>= drop;
        bc -= drop;
        in -= bc >> 3;
        bb = 0;
        bc = 0;
    }

    bool stored_block() {
        to_byte_boundary();
        if (in_end - in < 4) return false;
        const uint32_t len = in[0] | (in[1] << 8), nlen = in[2] | (in[3] << 8);
        if ((len ^ 0xFFFFu) != nlen) return false;
        in += 4;
        if ((size_t)(in_end - in) < len) return false;
        if (op > op_limit) { hit_limit = true; return false; }
        if (!out->reserve((op + len + 512) * sizeof(T))) return false;
        T *dst = (T *)out->p + op;
        for (uint32_t i = 0; i < len; ++i) dst[i] = in[i];
        in += len;
        op += len;
        return true;
    }

    // One block of Huffman-coded symbols.  Per match the dependent chain is: table load -> shift by (code + extra bits, one
    // number in the entry) -> table load -> shift; the extra-bit values and the copy hang off it.  One refill (>= 56 bits)
    // covers a length (<= 20 bits) and a distance (<= 28) -- a run of literals goes back to the top to refill.
    bool huffman_block(const uint32_t *lt, const uint32_t *dt, const uint64_t *mt) {
        const uint8_t *in_ = in, *const end_ = in_end;
        uint64_t b = bb;
        int c = bc;
        size_t o = op, cap = out->cap / sizeof(T);
        T *buf = (T *)out->p;
        bool ok = false;
        for (;;) {
            if (cap - o < 512) {                         // room for a run of literals + the longest match + copy overrun
                if (o > op_limit) { hit_limit = true; break; }
                if (!out->reserve((o + o / 2 + (1u << 20)) * sizeof(T))) break;
                buf = (T *)out->p;
                cap = out->cap / sizeof(T);
            }
            if (end_ - in_ >= 8) {
                uint64_t w;
                memcpy(&w, in_, 8);
                b |= w << c;
                in_ += (63 - c) >> 3;
                c |= 56;
            } else {
                while (c <= 56 && in_ < end_) { b |= (uint64_t)*in_++ << c; c += 8; }
            }
            uint64_t m = mt[b & ((1u << MULTI_BITS) - 1)];
            if (m >> 40) {                               // a run of literals, up to four per look-up
                do {
                    const uint32_t four = (uint32_t)m;
                    if (sizeof(T) == 1) {
                        memcpy(buf + o, &four, 4);
                    } else {
                        buf[o] = (T)(four & 0xFF); buf[o + 1] = (T)((four >> 8) & 0xFF);
                        buf[o + 2] = (T)((four >> 16) & 0xFF); buf[o + 3] = (T)(four >> 24);
                    }
                    o += (size_t)(m >> 40);
                    const int used = (int)((m >> 32) & 0xFF);
                    b >>= used; c -= used;
                    if (c < 32) break;
                    m = mt[b & ((1u << MULTI_BITS) - 1)];
                } while (m >> 40);
                if (c < 0) break;                        // the input ended inside a symbol
                continue;
            }
            uint32_t e = lt[b & ((1u << LITLEN_BITS) - 1)];
            if (e_kind(e) == K_SUB) {
                b >>= LITLEN_BITS; c -= LITLEN_BITS;
                e = lt[e_value(e) + (uint32_t)(b & ((1u << e_extra(e)) - 1))];
            }
            uint64_t s = b;
            uint32_t tot = e_bits(e);                    // code bits + extra bits
            b >>= tot; c -= (int)tot;
            if (c < 0) break;
            const uint32_t kind = e_kind(e);
            if (kind == K_LITERAL) { buf[o++] = (T)e_value(e); continue; }           // (a literal with a long code)
            if (kind != K_LENGTH) { ok = kind == K_EOB; break; }                      // end of block, or an unused code
            const uint32_t xl = e_extra(e);
            const uint32_t len = e_value(e) + (uint32_t)((s >> (tot - xl)) & ((1u << xl) - 1));
            uint32_t d = dt[b & ((1u << DIST_BITS) - 1)];
            if (e_kind(d) == K_SUB) {
                b >>= DIST_BITS; c -= DIST_BITS;
                d = dt[e_value(d) + (uint32_t)(b & ((1u << e_extra(d)) - 1))];
            }
            s = b;
            tot = e_bits(d);
            b >>= tot; c -= (int)tot;
            if (c < 0 || e_kind(d) != K_LENGTH) break;
            const uint32_t xd = e_extra(d);
            const size_t dist = e_value(d) + (size_t)((s >> (tot - xd)) & ((1u << xd) - 1));
            if (dist > o) break;                         // reaches back beyond the start of the output
            T *dst = buf + o;
            const T *src = dst - dist;
            if (dist >= 16) {
                for (uint32_t i = 0; i < len; i += 16) memcpy(dst + i, src + i, 16 * sizeof(T));
            } else if (dist == 1) {
                const T v = *src;
                for (uint32_t i = 0; i < len; ++i) dst[i] = v;
            } else {
                for (uint32_t i = 0; i < len; ++i) dst[i] = src[i];
            }
            o += len;
        }
        in = in_; bb = b; bc = c; op = o;
        return ok;
    }

    bool dynamic_tables(uint32_t *lt, uint32_t *dt) {
        static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint32_t hlit, hdist, hclen, v;
        if (!bits(5, &hlit) || !bits(5, &hdist) || !bits(4, &hclen)) return false;
        hlit += 257; hdist += 1; hclen += 4;
        if (hlit > 286 || hdist > 30) return false;
        uint8_t pre_lens[19] = {0};
        for (uint32_t i = 0; i < hclen; ++i) {
            if (!bits(3, &v)) return false;
            pre_lens[ORDER[i]] = (uint8_t)v;
        }
        uint32_t pt[1u << PRE_BITS];
        if (!build_table(pre_lens, 19, PRE_BITS, pt, 1u << PRE_BITS, pre_entry)) return false;
        uint8_t lens[286 + 30 + 138] = {0};
        uint32_t n = 0;
        while (n < hlit + hdist) {
            if (bc < 7) refill();
            const uint32_t e = pt[bb & ((1u << PRE_BITS) - 1)];
            if (e_kind(e) != K_LITERAL || (int)e_bits(e) > bc) return false;
            bb >>= e_bits(e); bc -= (int)e_bits(e);
            const uint32_t s = e_value(e);
            if (s < 16) { lens[n++] = (uint8_t)s; continue; }
            uint32_t rep, val = 0;
            if (s == 16) {
                if (n == 0 || !bits(2, &rep)) return false;
                rep += 3; val = lens[n - 1];
            } else if (s == 17) {
                if (!bits(3, &rep)) return false;
                rep += 3;
            } else {
                if (!bits(7, &rep)) return false;
                rep += 11;
            }
            if (n + rep > hlit + hdist) return false;
            for (uint32_t i = 0; i < rep; ++i) lens[n++] = (uint8_t)val;
        }
        if (lens[256] == 0) return false;                // no end-of-block code
        return build_table(lens, (int)hlit, LITLEN_BITS, lt, LITLEN_CAP, litlen_entry) &&
               build_table(lens + hlit, (int)hdist, DIST_BITS, dt, DIST_CAP, dist_entry);
    }

    // Blocks from the current position.  RUN_FINAL: the final block is done and `in` is the first byte after the stream.
    // RUN_STOPPED: the next block starts at bitpos() >= stop_bit (nothing of it has been read).  RUN_ERROR: invalid data.
    enum { RUN_ERROR = 0, RUN_STOPPED = 1, RUN_FINAL = 2 };
    int run(Tables &tb, uint64_t stop_bit = ~0ull) {
        for (;;) {
            if (bitpos() >= stop_bit) return RUN_STOPPED;
            uint32_t final_block, type;
            if (!bits(1, &final_block) || !bits(2, &type)) return RUN_ERROR;
            if (type == 0) {
                if (!stored_block()) return RUN_ERROR;
            } else if (type == 1) {
                if (!tb.fixed() || !huffman_block(tb.flt.data(), tb.fdt.data(), tb.fmt.data())) return RUN_ERROR;
            } else if (type == 2) {
                if (!dynamic_tables(tb.lt.data(), tb.dt.data())) return RUN_ERROR;
                build_multi(tb.lt.data(), tb.mt.data());
                if (!huffman_block(tb.lt.data(), tb.dt.data(), tb.mt.data())) return RUN_ERROR;
            } else {
                return RUN_ERROR;
            }
            if (final_block) break;
        }
        to_byte_boundary();
        return RUN_FINAL;
    }
};
typedef DecoderT<uint8_t> Decoder;

struct Member { size_t out_lo, out_hi; uint32_t crc, isize; };

// CRC-32 of [p, p + n) on up to n_threads threads (zlib's crc32_z per piece, crc32_combine)
static uint32_t crc32_parallel(const uint8_t *p, size_t n, unsigned n_threads) {
    const size_t PIECE = (size_t)32 << 20;
    const size_t n_pieces = (n + PIECE - 1) / PIECE;
    if (n_pieces <= 1 || n_threads <= 1) return (uint32_t)crc32_z(crc32_z(0, nullptr, 0), p, n);
    std::vector<uint32_t> part(n_pieces);
    std::vector<std::thread> th;
    const unsigned nt = (unsigned)std::min<size_t>(n_threads, n_pieces);
    auto work = [&](unsigned t) {
        for (size_t i = t; i < n_pieces; i += nt) {
            const size_t lo = i * PIECE, len = std::min(PIECE, n - lo);
            part[i] = (uint32_t)crc32_z(crc32_z(0, nullptr, 0), p + lo, len);
        }
    };
    for (unsigned t = 1; t < nt; ++t) {
        try { th.emplace_back(work, t); } catch (const std::exception &) { work(t); }
    }
    work(0);
    for (auto &t : th) t.join();
    uLong crc = part[0];
    for (size_t i = 1; i < n_pieces; ++i) crc = crc32_combine(crc, part[i], (z_off_t)std::min(PIECE, n - i * PIECE));
    return (uint32_t)crc;
}

// ---- one DEFLATE stream on several threads ----------------------------------------------------------------------------------
// The stream is cut into chunks of `chunk_bytes` of compressed data, P chunks per round.  The first chunk of a round starts
// at a known block boundary and is decoded straight into `out`.  Every other chunk has to FIND a block boundary first --
// bit offsets are tried one by one until a non-final dynamic-Huffman block header parses into two complete codes, the block
// decodes and another plausible header follows -- and, not knowing the 32 KB before it, decodes into 16-bit marker symbols
// (DecoderT<uint16_t>).  Each chunk stops at the first block boundary at or beyond its nominal end.  Then the chunks are
// stitched in order: the boundary the previous chunk stopped at must be exactly where this one started (if this one
// started later -- stored or fixed blocks are no starting points -- the previous decoder carries on up to it; if it started
// EARLIER, it started on a phantom: give up); the last 32 KB in front of each chunk are handed down the chain and all chunks
// are resolved into `out` in parallel.  Whatever goes wrong returns false, and so does a CRC mismatch afterwards (gunzip):
// the serial decoder then does the job.
struct MarkerChunk {
    OutBuf buf;                                          // uint16_t symbols: 32 768 window markers, then the output
    DecoderT<uint16_t> dec;
    Tables tb;
    uint64_t start_bit = ~0ull;                          // where it found its first block (~0: nowhere)
    int state = 0;                                       // RUN_* of its decoder
};
constexpr size_t WINDOW = 32768;
#ifndef CRF_INFLATE_MARKER_BUDGET
#define CRF_INFLATE_MARKER_BUDGET ((size_t)96 << 20)
#endif
constexpr size_t MARKER_BUDGET_MIN = CRF_INFLATE_MARKER_BUDGET;      // output symbols per chunk, at least

static bool marker_reset(MarkerChunk &c, size_t expect_out) {
    if (!c.buf.reserve((WINDOW + expect_out + (1u << 16)) * 2)) return false;
    uint16_t *q = (uint16_t *)c.buf.p;
    for (size_t j = 0; j < WINDOW; ++j) q[j] = (uint16_t)(256 + j);
    c.dec.out = &c.buf;
    c.dec.op = WINDOW;
    // Budget of a chunk: 3 x what FASTA text expands to, and room for the N runs of a genome (tens of Mbp that take a few KB of
    // input) -- buffers grow on demand, so only a chunk that does expand like that pays for it.  A chunk beyond it stops, the
    // decoder in front may spend the same again to cover it, and failing that the stream is left to one thread.
    c.dec.op_limit = WINDOW + std::max<size_t>(3 * expect_out, MARKER_BUDGET_MIN);
    c.dec.hit_limit = false;
    return true;
}

// first plausible block start in [from_bit, to_bit), decoding begun: on return the decoder stands behind the first block
static bool marker_find_start(MarkerChunk &c, const uint8_t *gz, size_t n, uint64_t from_bit, uint64_t to_bit, size_t expect_out) {
    for (uint64_t p = from_bit; p < to_bit; ++p) {
        const uint32_t three = (gz[p >> 3] | ((uint32_t)gz[(p >> 3) + 1] << 8)) >> (p & 7);
        if ((three & 7u) != 4u) continue;                // BFINAL = 0, BTYPE = 2 (dynamic), least significant bit first
        DecoderT<uint16_t> &d = c.dec;
        d.base = gz; d.in_end = gz + n;
        d.seek(p + 3);
        if (!d.dynamic_tables(c.tb.lt.data(), c.tb.dt.data())) continue;
        if (!marker_reset(c, expect_out)) return false;
        build_multi(c.tb.lt.data(), c.tb.mt.data());
        if (!d.huffman_block(c.tb.lt.data(), c.tb.dt.data(), c.tb.mt.data())) continue;
        {                                                // what follows must look like a block too
            DecoderT<uint16_t> peek = d;
            uint32_t fin, type;
            if (!peek.bits(1, &fin) || !peek.bits(2, &type) || type == 3) continue;
            if (type == 2) {
                Tables scratch;
                if (!peek.dynamic_tables(scratch.lt.data(), scratch.dt.data())) continue;
            } else if (type == 0) {
                peek.to_byte_boundary();
                if (peek.in_end - peek.in < 4) continue;
                const uint32_t len = peek.in[0] | (peek.in[1] << 8), nlen = peek.in[2] | (peek.in[3] << 8);
                if ((len ^ 0xFFFFu) != nlen) continue;
            }
        }
        c.start_bit = p;
        return true;
    }
    return false;
}

static bool gunzip_parallel(const uint8_t *gz, size_t n, size_t data_pos, OutBuf &out, size_t *op_io, size_t *end_pos,
                            unsigned n_threads, size_t chunk_bytes) {
    const unsigned P = std::min(16u, n_threads);
    if (P < 2 || n < data_pos + 2 * chunk_bytes + 64) return false;
    const size_t ratio = 5;                              // expected output per compressed byte (buffers grow if it is more)
    DecoderT<uint8_t> head;
    Tables head_tb;
    head.base = gz; head.in = gz + data_pos; head.in_end = gz + n; head.out = &out; head.op = *op_io;
    std::vector<MarkerChunk> chunks(P);
    const uint64_t end_bit = 8 * (uint64_t)(n - 8);      // (the trailer is no block)
    const bool trace = getenv("CRF_GUNZIP_TRACE") != nullptr;
    auto give_up = [&](const char *why) {                // (the caller then decodes the stream on one thread)
        if (trace) fprintf(stderr, "[crf_inflate] parallel: gave up: %s\n", why);
        return false;
    };
    unsigned n_rounds = 0, n_used = 0, n_carried = 0;
    std::vector<uint8_t> luts;
    double ms_decode = 0, ms_stitch = 0, ms_resolve = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
    for (;;) {
        ++n_rounds;
        auto t_round = now();
        // nominal chunk ends of this round (bytes); the round's first chunk starts where the stream stands
        const uint64_t here = head.bitpos();
        std::vector<uint64_t> cut(P + 1);
        for (unsigned k = 0; k <= P; ++k) cut[k] = std::min<uint64_t>(end_bit, here + 8 * (uint64_t)k * chunk_bytes);
        int head_state = 0;
        std::atomic<bool> threw{false};
        std::vector<std::thread> th;
        auto work = [&](unsigned k) {
          try {
            if (k == 0) {
                head_state = head.run(head_tb, cut[1]);
                return;
            }
            MarkerChunk &c = chunks[k];
            c.start_bit = ~0ull; c.state = 0;
            if (cut[k] >= end_bit) return;
            uint64_t from = cut[k];
            while (from < cut[k + 1] && marker_find_start(c, gz, n, from, cut[k + 1], ratio * chunk_bytes)) {
                c.state = c.dec.run(c.tb, cut[k + 1]);
                if (c.state != DecoderT<uint16_t>::RUN_ERROR) return;
                from = c.start_bit + 1;                  // it was no block after all: look further
                c.start_bit = ~0ull;
                if (c.dec.hit_limit) break;              // ... unless it was the output that outgrew the chunk's budget (a long
                                                         // run of N: 1000 bytes per byte): the decoder in front covers this range
            }
            c.state = 0;
          } catch (...) { threw = true; }                // (bad_alloc in a worker must not leave its thread)
        };
        for (unsigned k = 1; k < P; ++k) {
            try { th.emplace_back(work, k); } catch (const std::exception &) { work(k); }
        }
        work(0);
        for (auto &t : th) t.join();
        if (threw || head_state == Decoder::RUN_ERROR) return give_up("the first chunk of a round does not decode");
        ms_decode += ms_since(t_round);
        t_round = now();
        // stitch: `pos` = the true block boundary reached so far, by the decoder `last` (-1: the head)
        int last = -1;
        bool final_seen = head_state == Decoder::RUN_FINAL;
        uint64_t pos = final_seen ? 0 : head.bitpos();
        std::vector<unsigned> used;
        for (unsigned k = 1; k < P; ++k) {
            MarkerChunk &c = chunks[k];
            if (c.start_bit == ~0ull) continue;          // found nothing in its range: the previous decoder covers it
            if (final_seen) break;                       // (what later chunks found belongs to the next member of the file)
            if (pos > c.start_bit) return give_up("a chunk started on a phantom block");
            if (pos < c.start_bit) {                     // carry the previous decoder on up to this chunk's start
                int st;
                if (last < 0) { st = head.run(head_tb, c.start_bit); pos = head.bitpos(); }
                else {                                   // (what lies in between may be a chunk that stopped because its output
                    DecoderT<uint16_t> &pd = chunks[last].dec;                  //  outgrew the budget: the same budget again)
                    pd.op_limit = pd.op + std::max<size_t>(3 * ratio * chunk_bytes, MARKER_BUDGET_MIN);
                    st = pd.run(chunks[last].tb, c.start_bit); pos = pd.bitpos();
                }
                if (st == Decoder::RUN_FINAL) {          // the member ends before this chunk: it is the next member's
                    final_seen = true;
                    break;
                }
                if (st != Decoder::RUN_STOPPED || pos != c.start_bit) return give_up("the decoder in front does not arrive at a chunk's start");
                ++n_carried;
            }
            used.push_back(k);
            last = (int)k;
            final_seen = c.state == DecoderT<uint16_t>::RUN_FINAL;
            pos = final_seen ? 0 : c.dec.bitpos();
        }
        // resolve the marker chunks behind the head's output
        size_t total = head.op;
        std::vector<size_t> at(used.size());
        for (size_t i = 0; i < used.size(); ++i) { at[i] = total; total += chunks[used[i]].dec.op - WINDOW; }
        if (!out.reserve(total + total / 8 + (1u << 20))) return give_up("out of memory");
        std::vector<std::vector<uint8_t>> window(used.size(), std::vector<uint8_t>(WINDOW, 0));
        for (size_t i = 0; i < used.size(); ++i) {       // the 32 KB in front of each chunk, handed down the chain
            std::vector<uint8_t> &w = window[i];
            if (i == 0) {
                const size_t have = std::min(WINDOW, head.op);
                memcpy(w.data() + WINDOW - have, out.p + head.op - have, have);
            } else {
                const MarkerChunk &pc = chunks[used[i - 1]];
                const uint16_t *sym = (const uint16_t *)pc.buf.p;
                const size_t n_prev = pc.dec.op - WINDOW, have = std::min(WINDOW, n_prev);
                const std::vector<uint8_t> &pw = window[i - 1];
                memcpy(w.data(), pw.data() + have, WINDOW - have);
                for (size_t j = 0; j < have; ++j) {
                    const uint16_t v = sym[pc.dec.op - have + j];
                    w[WINDOW - have + j] = v < 256 ? (uint8_t)v : pw[v - 256];
                }
            }
        }
        th.clear();
        ms_stitch += ms_since(t_round);
        t_round = now();
        luts.resize(used.size() * (256 + WINDOW));       // (allocated here: nothing in the workers throws)
        auto resolve = [&](size_t i) {                   // (no allocation in here: nothing to throw)
            const MarkerChunk &c = chunks[used[i]];
            const uint16_t *sym = (const uint16_t *)c.buf.p + WINDOW;
            const size_t cnt = c.dec.op - WINDOW;
            const uint8_t *w = window[i].data();
            uint8_t *dst = out.p + at[i];
            // In DNA text nearly every byte is a copy of a copy (gzip finds a match everywhere), so markers do not die out
            // with distance from the chunk start: about half of ALL symbols are markers.  One table, no branch:
            uint8_t *lut = luts.data() + i * (256 + WINDOW);
            for (size_t v = 0; v < 256; ++v) lut[v] = (uint8_t)v;
            memcpy(lut + 256, w, WINDOW);
            size_t j = 0;
            for (; j + 4 <= cnt; j += 4) {
                uint64_t four;
                memcpy(&four, sym + j, 8);
                const uint32_t r = (uint32_t)lut[four & 0xFFFF] | ((uint32_t)lut[(four >> 16) & 0xFFFF] << 8) |
                                   ((uint32_t)lut[(four >> 32) & 0xFFFF] << 16) | ((uint32_t)lut[four >> 48] << 24);
                memcpy(dst + j, &r, 4);
            }
            for (; j < cnt; ++j) dst[j] = lut[sym[j]];
        };
        for (size_t i = 1; i < used.size(); ++i) {
            try { th.emplace_back(resolve, i); } catch (const std::exception &) { resolve(i); }
        }
        if (!used.empty()) resolve(0);
        for (auto &t : th) t.join();
        n_used += (unsigned)used.size();
        ms_resolve += ms_since(t_round);
        if (final_seen) {
            if (trace) fprintf(stderr, "[crf_inflate] parallel: %u round(s), %u marker chunk(s) stitched, %u carried on to a later start"
                               " (decode %.0f ms, stitch %.0f ms, resolve %.0f ms)\n",
                               n_rounds, n_used, n_carried, ms_decode, ms_stitch, ms_resolve);
            const uint8_t *after = last < 0 ? head.in : chunks[last].dec.in;
            *op_io = total;
            *end_pos = (size_t)(after - gz);
            return true;
        }
        // next round: the head decoder goes on from the boundary the last chunk reached
        if (pos <= here || pos >= end_bit) return give_up("no progress, or blocks that run into the trailer");
        head.op = total;
        if (last >= 0) head.seek(pos);
    }
}

#ifndef CRF_INFLATE_CHUNK_BYTES
#define CRF_INFLATE_CHUNK_BYTES (8u << 20)
#endif

// All members of a gzip file -> out.  False: not handled here (the caller lets zlib decide).
// parallel_chunk_bytes: chunk size of the several-threads decoder (0: the default; ~0: one thread only).
static bool gunzip(const uint8_t *gz, size_t n, OutBuf &out, unsigned n_threads, size_t parallel_chunk_bytes = 0) {
    if (n < 18) return false;
    if (!parallel_chunk_bytes) {
        const char *kb = getenv("CRF_GUNZIP_CHUNK_KB");  // (tests: small chunks send small inputs through the several-threads path)
        if (kb && atoi(kb) > 0) {
            parallel_chunk_bytes = (size_t)atoi(kb) << 10;
        } else {                                         // 8 MB, less when that would leave threads without a chunk (>= 1 MB:
            const size_t P = std::min(16u, std::max(1u, n_threads));             // finding a block start costs 1-15 ms)
            parallel_chunk_bytes = std::min<size_t>(CRF_INFLATE_CHUNK_BYTES, std::max<size_t>((size_t)1 << 20, n / P + 1));
        }
    }
    const uint32_t isize_hint = gz[n - 4] | (gz[n - 3] << 8) | (gz[n - 2] << 16) | ((uint32_t)gz[n - 1] << 24);
    if (!out.reserve(std::max<size_t>((size_t)isize_hint, n * 4) + (1u << 20))) return false;
    std::vector<Member> members;
    Tables tb;
    size_t pos = 0, op = 0;
    while (pos < n) {
        if (n - pos < 18 || gz[pos] != 0x1f || gz[pos + 1] != 0x8b || gz[pos + 2] != 8) return false;
        const uint8_t flg = gz[pos + 3];
        if (flg & 0xE0) return false;
        size_t p = pos + 10;
        if (flg & 4) {                                   // FEXTRA
            if (n - p < 2) return false;
            const size_t xlen = gz[p] | (gz[p + 1] << 8);
            p += 2;
            if (n - p < xlen) return false;
            p += xlen;
        }
        for (int f = 8; f <= 16; f <<= 1)                // FNAME, FCOMMENT: NUL-terminated
            if (flg & f) {
                const void *z = memchr(gz + p, 0, n - p);
                if (!z) return false;
                p = (size_t)((const uint8_t *)z - gz) + 1;
            }
        if (flg & 2) p += 2;                             // FHCRC
        if (p >= n) return false;
        Member m;
        m.out_lo = op;
        size_t after = 0, op_par = op;
        // several threads for a member with at least two chunks of input behind its header (for the last or only member that
        // is known; an earlier one may turn out shorter: the chunks beyond its end are then dropped)
        if (parallel_chunk_bytes != ~(size_t)0 && n_threads > 1 &&
            gunzip_parallel(gz, n, p, out, &op_par, &after, n_threads, parallel_chunk_bytes) && after + 8 <= n) {
            op = op_par;
            p = after;
        } else {
            Decoder d;
            d.base = gz; d.in = gz + p; d.in_end = gz + n; d.out = &out; d.op = op;
            if (d.run(tb) != Decoder::RUN_FINAL) return false;
            op = d.op;
            p = (size_t)(d.in - gz);
        }
        if (n - p < 8) return false;
        m.out_hi = op;
        m.crc = gz[p] | (gz[p + 1] << 8) | (gz[p + 2] << 16) | ((uint32_t)gz[p + 3] << 24);
        m.isize = gz[p + 4] | (gz[p + 5] << 8) | (gz[p + 6] << 16) | ((uint32_t)gz[p + 7] << 24);
        if ((uint32_t)(m.out_hi - m.out_lo) != m.isize) return false;
        members.push_back(m);
        pos = p + 8;
    }
    for (const Member &m : members)
        if (crc32_parallel(out.p + m.out_lo, m.out_hi - m.out_lo, n_threads) != m.crc) return false;
    out.size = op;
    return true;
}

}  // namespace crf_inflate
