// gzip -> bytes for the FASTA reader (host only; part of row f1 of DESIGN.md: the reference reads `ref.fa.gz` through
// pyfastx, perfect_repeat_finder.py:117).
//
// A plain gzip file is ONE DEFLATE stream (RFC 1951/1952) and cannot be split over threads, so how fast it is decoded
// decides how long a `.fa.gz` run takes: zlib's inflate does ~110 MB/s of FASTA text (literal-heavy, 2-3 bit codes, one
// symbol per loop), 25 s for a human genome that is scanned in milliseconds.  This decoder is written for that data:
//   * 64-bit bit buffer, refilled branch-free with one unaligned load (8 input bytes are always there in the main loop);
//   * one 11-bit table look-up per literal/length symbol (sub-tables behind it for the rare longer codes), literals
//     decoded back to back without refilling in between; 8-bit table for the distance symbols;
//   * the whole output is one contiguous buffer (realloc/mremap-grown), so a match is a plain copy from earlier output --
//     no 32 KB window to maintain -- done 16 bytes at a time when the distance allows;
//   * the CRC-32 of every member is checked afterwards on several threads (crc32_combine), as is ISIZE.
// Anything unexpected (a code set zlib would reject, a bad CRC, trailing bytes that are no gzip member) makes gunzip()
// return false and the caller falls back to zlib, which then gives the verdict.  tests/test_host_cpu.py compares both with
// Python's zlib on streams of every block type.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <thread>
#include <vector>

namespace crf_inflate {

struct OutBuf {                                          // malloc'ed so that growing it is a realloc (mremap for big blocks)
    uint8_t *p = nullptr;
    size_t size = 0, cap = 0;
    ~OutBuf() { free(p); }
    bool reserve(size_t need) {
        if (need <= cap) return true;
        uint8_t *q = (uint8_t *)realloc(p, need);
        if (!q) return false;
        p = q;
        cap = need;
        return true;
    }
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

struct Decoder {
    const uint8_t *in, *in_end;
    uint64_t bb = 0;                                     // bit buffer: the next input bits, least significant first
    int bc = 0;                                          // valid bits in it
    OutBuf *out;
    size_t op;                                           // bytes of output so far

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
        if (!out->reserve(op + len + 512)) return false;
        memcpy(out->p + op, in, len);
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
        size_t o = op, cap = out->cap;
        uint8_t *buf = out->p;
        bool ok = false;
        for (;;) {
            if (cap - o < 512) {                         // room for a run of literals + the longest match + copy overrun
                if (!out->reserve(o + o / 2 + (1u << 20))) break;
                buf = out->p;
                cap = out->cap;
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
                    memcpy(buf + o, &four, 4);
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
            if (kind == K_LITERAL) { buf[o++] = (uint8_t)e_value(e); continue; }     // (a literal with a long code)
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
            uint8_t *dst = buf + o;
            const uint8_t *src = dst - dist;
            if (dist >= 16) {
                for (uint32_t i = 0; i < len; i += 16) memcpy(dst + i, src + i, 16);
            } else if (dist == 1) {
                memset(dst, *src, len);
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

    // one raw DEFLATE stream from `in`; on success `in` is the first byte after it
    bool inflate_stream() {
        std::vector<uint32_t> lt(LITLEN_CAP), dt(DIST_CAP), flt, fdt;
        std::vector<uint64_t> mt(1u << MULTI_BITS), fmt;
        for (;;) {
            uint32_t final_block, type;
            if (!bits(1, &final_block) || !bits(2, &type)) return false;
            if (type == 0) {
                if (!stored_block()) return false;
            } else if (type == 1) {
                if (flt.empty()) {                       // the fixed code of RFC 1951 3.2.6
                    uint8_t l[288], d[32];
                    for (int s = 0; s < 288; ++s) l[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
                    for (int s = 0; s < 32; ++s) d[s] = 5;
                    flt.resize(LITLEN_CAP); fdt.resize(DIST_CAP);
                    if (!build_table(l, 288, LITLEN_BITS, flt.data(), LITLEN_CAP, litlen_entry) ||
                        !build_table(d, 32, DIST_BITS, fdt.data(), DIST_CAP, dist_entry)) return false;
                    fmt.resize(1u << MULTI_BITS);
                    build_multi(flt.data(), fmt.data());
                }
                if (!huffman_block(flt.data(), fdt.data(), fmt.data())) return false;
            } else if (type == 2) {
                if (!dynamic_tables(lt.data(), dt.data())) return false;
                build_multi(lt.data(), mt.data());
                if (!huffman_block(lt.data(), dt.data(), mt.data())) return false;
            } else {
                return false;
            }
            if (final_block) break;
        }
        to_byte_boundary();
        return true;
    }
};

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

// All members of a gzip file -> out.  False: not handled here (the caller lets zlib decide).
static bool gunzip(const uint8_t *gz, size_t n, OutBuf &out, unsigned n_threads) {
    if (n < 18) return false;
    const uint32_t isize_hint = gz[n - 4] | (gz[n - 3] << 8) | (gz[n - 2] << 16) | ((uint32_t)gz[n - 1] << 24);
    if (!out.reserve(std::max<size_t>((size_t)isize_hint, n * 4) + (1u << 20))) return false;
    std::vector<Member> members;
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
        Decoder d;
        d.in = gz + p; d.in_end = gz + n; d.out = &out; d.op = op;
        if (!d.inflate_stream()) return false;
        p = (size_t)(d.in - gz);
        if (n - p < 8) return false;
        Member m;
        m.out_lo = op; m.out_hi = d.op;
        m.crc = gz[p] | (gz[p + 1] << 8) | (gz[p + 2] << 16) | ((uint32_t)gz[p + 3] << 24);
        m.isize = gz[p + 4] | (gz[p + 5] << 8) | (gz[p + 6] << 16) | ((uint32_t)gz[p + 7] << 24);
        if ((uint32_t)(m.out_hi - m.out_lo) != m.isize) return false;
        members.push_back(m);
        op = d.op;
        pos = p + 8;
    }
    for (const Member &m : members)
        if (crc32_parallel(out.p + m.out_lo, m.out_hi - m.out_lo, n_threads) != m.crc) return false;
    out.size = op;
    return true;
}

}  // namespace crf_inflate
