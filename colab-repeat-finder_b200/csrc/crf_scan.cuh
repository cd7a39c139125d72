// crf_scan.cuh -- the scan kernel: every motif size k in [kmin, kmax] over one shared-memory tile.
//
// Replaces the reference's lock-step loop (perfect_repeat_finder.py:66-74) and the tracker state
// machine (utils/perfect_repeat_tracker.py:43-101).  Closed form being computed (SURVEY App. A.2,
// checked against the CPU restatement by tests/): for each k, every maximal run [st, i0) of ones of
//   M_k[j] = (S[j] == S[j+k]) and S[j] != 'N'
// with i0 - st >= r_min(k) = max(min_span - k, (min_repeats-1)*k) and a primitive motif S[st:st+k]
// is reported as (st, i0 + k, k).  (min_repeats == 1: r_min = max(min_span - k, k-1); the host table carries it, and
// single_copy_filter_kernel settles the one base such a run leaves untested -- DESIGN.md section 6.)
//
// One CTA owns one tile of 256*T words (T*8192 bases) and emits the runs that START in it:
//   1. stage the tile (one word of left context, a halo of kmax bases) of the H, L and NM planes in
//      shared memory (bank-conflict-free padded layout);
//   2. fast phase -- each thread keeps a strip of T(+1) consecutive words of H and L in registers
//      and, for every k, forms the shifted-compare words with funnel shifts and tests a
//      *necessary* condition for "a qualifying run starts in my strip" (FilterMode); the motif
//      sizes come as segments of constant filter, so the inner loop has no table look-ups and no
//      branches; for 2 <= k <= 16 positions that start a stretch of more than 8 / 16 equal bases count
//      as mismatches (no primitive motif of that size can contain them), which removes the
//      candidates that homopolymers would otherwise raise for every multiple of 1; the result is
//      one bit per (strip, k) in a register;
//   3. exact phase -- the block compacts the (strip, k) hits (prefix sum, no atomics); one
//      thread per hit re-evaluates the exact mask from shared memory (N mask; exotic symbols via
//      global memory), erodes it by min(r_min, 32) across word boundaries -- the rising edges of the
//      eroded mask are the starts of runs of at least that many matches --, walks right to the run
//      end, applies the thresholds and the primitivity rule and appends (start, end, k) to the
//      tile's shared-memory result list;
//   4. runs longer than walk_limit words are finished by the whole block (256 words per step);
//   5. the tile's results are ordered by (start, end) with a counting sort over strips and
//      written as one segment; a later pass concatenates the segments in tile order, which is
//      global (start, end) order because a run is owned by the tile of its start.
#pragma once
#include "crf_device.cuh"

#ifndef CRF_MIN_CTAS
#define CRF_MIN_CTAS 4   // resident CTAs per SM the block-tiled kernel is compiled for (4 x 256 threads x 64 registers = the register file)
#endif
namespace crf {

__host__ __device__ __forceinline__ uint32_t pad_idx(uint32_t v) { return v + (v >> 5); }

struct TileCtx {
    const uint32_t *sH, *sL, *sN;  // padded; index j <-> absolute word (wbase + j)
    uint32_t wbase;                // absolute word of smem index 0, plus 1 (so tile 0 needs no negative)
    uint32_t nsm;                  // words staged
    uint64_t *key;
    uint16_t *kk;
    uint32_t *nout, *nlong, *longq;
    uint32_t cap;
    uint2 *startq;                 // run starts found by stage A of the exact phase
    uint32_t *nstart;
    uint32_t longcap, startcap;    // capacities of longq / startq
};

// Exact M_k word w: from the staged tile when everything needed is there, else from global memory.
__device__ __forceinline__ uint32_t tile_mask(const ScanParams &p, const TileCtx &t, uint32_t k, uint32_t w) {
    const uint32_t q = k >> 5, s = k & 31;
    const uint32_t j = w + 1 - t.wbase;  // wbase = w0, smem index 0 = word w0-1
    if (p.n_exotic == 0 && w + 1 >= t.wbase && j + q + 1 < t.nsm) {
        const uint32_t a = pad_idx(j), b = pad_idx(j + q), c = pad_idx(j + q + 1);
        const uint32_t dh = t.sH[a] ^ __funnelshift_r(t.sH[b], t.sH[c], s);
        const uint32_t dl = t.sL[a] ^ __funnelshift_r(t.sL[b], t.sL[c], s);
        const uint32_t nm = t.sN[a] | __funnelshift_r(t.sN[b], t.sN[c], s);
        return ~(dh | dl | nm);
    }
    return exact_mask(p, k, w);
}

__device__ inline bool tile_walk(const ScanParams &p, const TileCtx &t, uint32_t k, uint32_t from, uint32_t limit,
                                 uint32_t *i0) {
    uint32_t w = from >> 5;
    uint32_t m = tile_mask(p, t, k, w) | ((1u << (from & 31)) - 1u);
    for (uint32_t n = 0;; ++n) {
        if (m != 0xFFFFFFFFu) { *i0 = (w << 5) + (__ffs(~m) - 1); return true; }
        ++w;
        if (n >= limit || w >= p.n_words) {  // beyond n_words everything is masked
            *i0 = w << 5;
            return w >= p.n_words;
        }
        m = tile_mask(p, t, k, w);
    }
}

__device__ inline bool tile_all_match(const ScanParams &p, const TileCtx &t, uint32_t d, uint32_t a, uint32_t len) {
    uint32_t pos = a;
    const uint32_t end = a + len;
    while (pos < end) {
        const uint32_t b = pos & 31;
        const uint32_t n = min(32u - b, end - pos);
        const uint32_t mask = (n == 32 ? 0xFFFFFFFFu : ((1u << n) - 1u)) << b;
        if ((tile_mask(p, t, d, pos >> 5) & mask) != mask) return false;
        pos += n;
    }
    return true;
}

__device__ inline void emit_result(const ScanParams &p, const TileCtx &t, uint32_t st, uint32_t end, uint32_t k) {
    const uint64_t key = ((uint64_t)st << 32) | end;
    const uint32_t idx = atomicAdd(t.nout, 1u);
    if (idx < t.cap) {
        t.key[idx] = key;
        t.kk[idx] = (uint16_t)k;
    } else {
        const unsigned long long g = atomicAdd(p.counters + C_SPILL, 1ull);
        if (g < p.spill_cap) {
            p.spill_key[g] = key;
            p.spill_k[g] = (uint16_t)k;
        }
    }
}

// A run start: M_k[st .. st+known) are ones and M_k[st-1] is zero.
__device__ inline void handle_start(const ScanParams &p, const TileCtx &t, const KEntry &ke, uint32_t k, uint32_t st,
                                    uint32_t known) {
    if (p.own_lo) {  // partitioned load: a run belongs to the unit that owns its start
        uint32_t lo = 0, hi = p.n_records - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (__ldg(p.rec_dev_off + mid) <= st) lo = mid; else hi = mid - 1;
        }
        if (st < __ldg(p.own_lo + lo) || st >= __ldg(p.own_hi + lo)) return;
    }
    uint32_t i0;
    bool found = tile_walk(p, t, k, st + known, p.walk_limit, &i0);
    while (!found && i0 - st < ke.rmin) found = tile_walk(p, t, k, i0, p.walk_limit, &i0);
    if (found && i0 - st < ke.rmin) return;                       // trk:86/91 thresholds
#pragma unroll 1
    for (int j = 0; j < 6; ++j) {                                 // primitivity, trk:98,108-142
        const uint32_t d = ke.div[j];
        if (!d) break;
        if (tile_all_match(p, t, d, st, k - d)) return;
    }
    if (found) {
        emit_result(p, t, st, i0 + k, k);                         // end = i0 + k (trk:87-92)
        return;
    }
    const uint32_t slot = atomicAdd(t.nlong, 1u);
    if (slot < t.longcap) {
        t.longq[3 * slot] = st;
        t.longq[3 * slot + 1] = k;
        t.longq[3 * slot + 2] = i0 >> 5;
    } else {  // queue full: finish it alone (correct, slow, practically never)
        while (!tile_walk(p, t, k, i0, 0x7FFFFFFFu, &i0)) {}
        emit_result(p, t, st, i0 + k, k);
    }
}

constexpr uint32_t STARTQ_CAP = 512;  // overflow queue: run starts beyond two per (strip, k) hit
constexpr uint32_t NBATCH = 4;        // groups of 32 motif sizes whose hits share one exact phase

// Exact phase for one (strip, k) hit.  The strip's words of the exact mask M_k are rebuilt from the
// staged planes (HASN = false: the tile has no masked position, the N plane is not read), each
// word is eroded by re = min(r_min, 32) across the word boundary, and the rising edges of the
// eroded mask are exactly the starts of runs of >= re matches.  Each start is then followed to the
// end of its run (handle_start).  All threads of a warp run the same fixed-length loop.
template <bool HASN>
__device__ inline void exact_item(const ScanParams &p, const TileCtx &t, uint32_t wfirst, uint32_t nw, uint32_t k) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(p.ktab + k));  // rmin, mode/sh, esh[0..3], esh[4]/re/div0
    const uint32_t e0 = raw.z & 0xFF, e1 = (raw.z >> 8) & 0xFF, e2 = (raw.z >> 16) & 0xFF, e3 = raw.z >> 24;
    const uint32_t e4 = raw.w & 0xFF, re = (raw.w >> 8) & 0xFF;
    const uint32_t q = k >> 5, s = k & 31;
    const bool in_smem = (p.n_exotic == 0);  // strip + look-ahead + halo are always staged
    uint32_t j = wfirst - t.wbase;           // smem index of word wfirst - 1
    uint32_t hb = 0, lb = 0, nb = 0;
    if (in_smem) {
        const uint32_t b0 = pad_idx(j + q);
        hb = t.sH[b0]; lb = t.sL[b0];
        if (HASN) nb = t.sN[b0];
    }
    uint32_t wnext = wfirst - 1;             // absolute word the next call returns (wraps for wfirst == 0)
    auto next_mask = [&]() -> uint32_t {
        uint32_t m;
        if (in_smem) {
            const uint32_t a = pad_idx(j), c = pad_idx(j + q + 1);
            const uint32_t hc = t.sH[c], lc = t.sL[c];
            uint32_t x = (t.sH[a] ^ __funnelshift_r(hb, hc, s)) | (t.sL[a] ^ __funnelshift_r(lb, lc, s));
            if (HASN) {
                const uint32_t nc = t.sN[c];
                x |= t.sN[a] | __funnelshift_r(nb, nc, s);
                nb = nc;
            }
            m = ~x;
            hb = hc; lb = lc;
            ++j;
        } else {
            m = (wnext == 0xFFFFFFFFu) ? 0u : exact_mask(p, k, wnext);
        }
        ++wnext;
        return m;
    };
    uint32_t prevtop = next_mask() >> 31;  // word wfirst-1 (tile 0 stages an all-masked word there)
    uint32_t cur = next_mask();
    uint32_t st0 = NOPOS, st1 = NOPOS;
#pragma unroll 1
    for (uint32_t i = 0; i < nw; ++i) {
        const uint32_t nxt = next_mask();
        uint64_t v = ((uint64_t)nxt << 32) | cur;  // bit j <- bits j .. j+re-1 all set
        v &= v >> e0; v &= v >> e1; v &= v >> e2; v &= v >> e3; v &= v >> e4;
        uint32_t starts = (uint32_t)v & ~((cur << 1) | prevtop);
        while (starts) {  // rare: genuine starts of runs of >= re matches
            const uint32_t st = ((wfirst + i) << 5) + (__ffs(starts) - 1);
            starts &= starts - 1;
            if (st0 == NOPOS) st0 = st;
            else if (st1 == NOPOS) st1 = st;
            else {
                const uint32_t slot = atomicAdd(t.nstart, 1u);
                if (slot < t.startcap) t.startq[slot] = make_uint2(st, k | (re << 16));
                else handle_start(p, t, p.ktab[k], k, st, re);
            }
        }
        prevtop = cur >> 31;
        cur = nxt;
    }
    if (st0 != NOPOS) {
        const KEntry ke = p.ktab[k];
#pragma unroll 1
        for (uint32_t st = st0; st != NOPOS; st = st1, st1 = NOPOS) handle_start(p, t, ke, k, st, re);
    }
}

// mask |= bit, predicated on the filter's verdict: one ISETP + one predicated LOP3 instead of SEL + SHF.L + LOP3 behind the
// compare (the bit itself lives in the uniform datapath: it only depends on the motif size)
__device__ __forceinline__ void set_bit_if(uint32_t &mask, bool hit, uint32_t bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p or.b32 %0, %0, %2;\n\t}" : "+r"(mask) : "r"((uint32_t)hit), "r"(bit));
}

// ---- fast-phase filters (T+1 words: the strip plus one look-ahead word) ------------------------
template <int T>
__device__ __forceinline__ bool filter_word(const uint32_t (&NH)[T + 1], const uint32_t (&FH)[T + 2], uint32_t s) {
    bool hit = false;
#pragma unroll
    for (int i = 0; i <= T; ++i) hit |= (NH[i] == __funnelshift_r(FH[i], FH[i + 1], s));
    return hit;
}
// "some aligned half-word of the compare is zero" = the minimum over all half-words is zero: one packed three-input
// minimum (VIMNMX3.U16x2, ALU rate on sm_100a -- profiles/microbench/dpx_min.cu) per TWO words instead of a
// subtract + LOP3 per word
template <int T>
__device__ __forceinline__ bool filter_half(const uint32_t (&NH)[T + 1], const uint32_t (&FH)[T + 2], uint32_t s) {
    uint32_t acc = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i + 1 <= T; i += 2) {
        const uint32_t x0 = NH[i] ^ __funnelshift_r(FH[i], FH[i + 1], s);
        const uint32_t x1 = NH[i + 1] ^ __funnelshift_r(FH[i + 1], FH[i + 2], s);
        acc = __vimin3_u16x2(acc, x0, x1);
    }
    if (((T + 1) & 1) != 0) acc = __vminu2(acc, NH[T] ^ __funnelshift_r(FH[T], FH[T + 1], s));
    return ((acc - 0x00010001u) & ~acc & 0x80008000u) != 0;
}
// SUP: positions that start a stretch of > k equal bases (HD = dilated mismatch word of M'_1, ~HD = such
// positions) are treated as mismatches.  A primitive motif of size k cannot contain k+1 equal consecutive
// bases, so no run that will be reported loses a position; runs of k that only exist because of a
// homopolymer (they are dropped by the primitivity rule anyway) stop producing candidates.
template <int T, bool SUP>
__device__ __forceinline__ bool filter_byte(const uint32_t (&NH)[T + 1], const uint32_t (&NL)[T + 1],
                                            const uint32_t (&FH)[T + 2], const uint32_t (&FL)[T + 2], uint32_t s,
                                            const uint32_t (&HD)[T + 2]) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i <= T; ++i) {
        uint32_t x = NH[i] ^ __funnelshift_r(FH[i], FH[i + 1], s);
        if (SUP) x |= ~HD[i];
        x |= NL[i] ^ __funnelshift_r(FL[i], FL[i + 1], s);
        acc |= (x - 0x01010101u) & ~x;
    }
    return (acc & 0x80808080u) != 0;
}
// NSH = number of dilation steps (compile time), shifts at run time
template <int T, int NSH, bool SUP>
__device__ __forceinline__ bool filter_erode(const uint32_t (&NH)[T + 1], const uint32_t (&NL)[T + 1],
                                             const uint32_t (&FH)[T + 2], const uint32_t (&FL)[T + 2], uint32_t s,
                                             uint32_t sh0, uint32_t sh1, uint32_t sh2, const uint32_t (&HD)[T + 2]) {
    uint32_t x[T + 1];  // mismatch words; dilating mismatches == eroding matches
#pragma unroll
    for (int i = 0; i <= T; ++i) {
        uint32_t v = NH[i] ^ __funnelshift_r(FH[i], FH[i + 1], s);
        if (SUP) v |= ~HD[i];
        x[i] = v | (NL[i] ^ __funnelshift_r(FL[i], FL[i + 1], s));
    }
    if (NSH >= 1) {
#pragma unroll
        for (int i = 0; i < T; ++i) x[i] |= __funnelshift_r(x[i], x[i + 1], sh0);
        x[T] |= x[T] >> sh0;
    }
    if (NSH >= 2) {
#pragma unroll
        for (int i = 0; i < T; ++i) x[i] |= __funnelshift_r(x[i], x[i + 1], sh1);
        x[T] |= x[T] >> sh1;
    }
    if (NSH >= 3) {
#pragma unroll
        for (int i = 0; i < T; ++i) x[i] |= __funnelshift_r(x[i], x[i + 1], sh2);
    }
    uint32_t all = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < T; ++i) all &= x[i];
    return all != 0xFFFFFFFFu;
}

// dynamic shared memory a scan block needs (NS = strips per thread: the tile is THREADS * T * NS words)
__host__ __device__ inline uint32_t scan_tile_words(int T, uint32_t kmax, int NS = 1) {
    return (uint32_t)THREADS * T * NS + (kmax >> 5) + 5;  // 1 word of left context + tile + halo (q+4)
}
__host__ __device__ inline size_t scan_smem_bytes(int T, uint32_t kmax, uint32_t outcap, int NS = 1) {
    const size_t plane = pad_idx(scan_tile_words(T, kmax, NS)) + 1;
    size_t words = 3 * plane + (NBATCH + 1) * THREADS * NS + 32 + 3 * LONGCAP;
    words = (words + 1) & ~(size_t)1;
    return words * 4 + (size_t)STARTQ_CAP * 8 + (size_t)outcap * 8 + (((size_t)outcap * 2 + 7) & ~(size_t)7);
}

// NS strips per thread: with NS = 2 a tile holds twice the (strip, k) hits, so the exact phase runs in ~3 rounds of which
// only the last is partly filled -- the warps that have no item in the last round wait 1/3 as long at the barrier that
// closes the phase as with NS = 1 (2 rounds, the second ~40 % filled); registers are unchanged (the strips of a thread are
// processed one after the other), shared memory grows by the second set of plane words.
template <int T, int NS = 1>
__global__ void __launch_bounds__(THREADS, (T <= 8 && NS == 1) ? CRF_MIN_CTAS : (T <= 8 ? 3 : 1)) scan_kernel(const ScanParams p) {
    constexpr int TW = THREADS * T * NS;
    constexpr uint32_t STRIPS = THREADS * NS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nsm = scan_tile_words(T, p.kmax, NS);
    const uint32_t plane = pad_idx(nsm) + 1;

    uint32_t *sH = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *sL = sH + plane;
    uint32_t *sN = sL + plane;
    uint32_t *s_hit = sN + plane;               // NBATCH x STRIPS hit masks (later: strip histogram)
    uint32_t *s_pre = s_hit + NBATCH * STRIPS;
    uint32_t *s_misc = s_pre + STRIPS;  // [0..7] warp sums [8] nout [9] nlong [10] nstart/minpos [11] base [12..15] q [16..23] warp sums (2nd strip set)
    uint32_t *s_long = s_misc + 32;
    size_t off_words = 3 * (size_t)plane + (NBATCH + 1) * STRIPS + 32 + 3 * LONGCAP;
    off_words = (off_words + 1) & ~(size_t)1;
    uint2 *s_startq = reinterpret_cast<uint2 *>(smem_raw + off_words * 4);
    uint64_t *s_key = reinterpret_cast<uint64_t *>(s_startq + STARTQ_CAP);
    uint16_t *s_k = reinterpret_cast<uint16_t *>(s_key + p.outcap);

    const uint32_t tile = blockIdx.x;
    const uint32_t w0 = tile * TW;
    // smem index j <-> absolute word w0 - 1 + j
    uint32_t any_mask = 0;
    for (uint32_t j = tid; j < nsm; j += THREADS) {
        const bool real = (w0 + j) != 0;
        const uint32_t w = w0 + j - 1;
        const uint32_t a = pad_idx(j);
        sH[a] = real ? __ldg(p.H + w) : 0u;
        sL[a] = real ? __ldg(p.L + w) : 0u;
        const uint32_t nmw = real ? __ldg(p.NM + w) : 0xFFFFFFFFu;
        sN[a] = nmw;
        any_mask |= nmw;
    }
    if (tid < 32) s_misc[tid] = 0;
    const bool tile_has_n = __syncthreads_or(any_mask != 0);  // masked positions anywhere in the staged words

    TileCtx tc;
    tc.sH = sH; tc.sL = sL; tc.sN = sN; tc.wbase = w0; tc.nsm = nsm;
    tc.key = s_key; tc.kk = s_k; tc.nout = &s_misc[8]; tc.nlong = &s_misc[9]; tc.longq = s_long; tc.cap = p.outcap;
    tc.startq = s_startq; tc.nstart = &s_misc[10];
    tc.longcap = LONGCAP; tc.startcap = STARTQ_CAP;

    unsigned long long ncand = 0;
    uint32_t si = 0;
    while (si < p.n_segs) {
        // ---- fast phase: up to NBATCH groups of <= 32 motif sizes (one group per k >> 5), every strip of this thread
        const uint32_t si0 = si;
        uint32_t nb = 0;
#pragma unroll 1
        for (int sub = 0; sub < NS; ++sub) {
        const uint32_t sbase = 1 + (sub * THREADS + tid) * T;   // smem index of the strip's first word
        uint32_t NH[T + 1], NL[T + 1];
#pragma unroll
        for (int i = 0; i <= T; ++i) {
            NH[i] = sH[pad_idx(sbase + i)];
            NL[i] = sL[pad_idx(sbase + i)];
        }
        si = si0;
        for (nb = 0; nb < NBATCH && si < p.n_segs; ++nb) {
            const uint32_t qb = (uint32_t)p.segs[si].k_lo >> 5;
            uint32_t FH[T + 2], FL[T + 2];
#pragma unroll
            for (int i = 0; i <= T + 1; ++i) {
                FH[i] = sH[pad_idx(sbase + qb + i)];
                FL[i] = sL[pad_idx(sbase + qb + i)];
            }
            // homopolymer mask for this strip (only k <= 16, i.e. the first group): HD bit j = some mismatch of
            // M'_1 in [j, j+8): ~HD = nine equal bases from j on; the word after the window is filled with mismatches
            // HD: all ones = no suppression.  It becomes the homopolymer mask when the first segment that wants one comes up
            // (2 <= k <= 16 of the ERODE / BYTE filters) and goes back to all ones for a later segment without suppression.
            uint32_t HD[T + 2];
            uint32_t hd_level = 0;
#pragma unroll
            for (int i = 0; i <= T + 1; ++i) HD[i] = 0xFFFFFFFFu;
            uint32_t hitmask = 0;
            for (; si < p.n_segs; ++si) {
                const Seg sg = p.segs[si];
                if (((uint32_t)sg.k_lo >> 5) != qb) break;
                const uint32_t s_lo = sg.k_lo & 31, s_hi = sg.k_hi & 31;
                const uint32_t sh0 = sg.sh0, sh1 = sg.sh1, sh2 = sg.sh2;
                if ((sg.mode & 15u) == MODE_WORD) {
#pragma unroll 1
                    for (uint32_t s = s_lo, bit = 1u << s_lo; s <= s_hi; ++s, bit <<= 1) set_bit_if(hitmask, filter_word<T>(NH, FH, s), bit);
                } else if ((sg.mode & 15u) == MODE_HALF) {
#pragma unroll 1
                    for (uint32_t s = s_lo, bit = 1u << s_lo; s <= s_hi; ++s, bit <<= 1) set_bit_if(hitmask, filter_half<T>(NH, FH, s), bit);
                } else {
                    const uint32_t sup = (sg.mode >> 4) & 3u;  // 0: none, 1: stretches of > 8 equal bases, 2: > 16
                    if (sup && hd_level == 0) {          // HD bit j = some mismatch of M'_1 in [j, j+8): ~HD = nine equal bases from j on;
                        const uint32_t hx = sH[pad_idx(sbase + T + 2)], lx = sL[pad_idx(sbase + T + 2)];   // (only k <= 16: group 0, FH = the strip itself)
#pragma unroll
                        for (int i = 0; i <= T; ++i)
                            HD[i] = (FH[i] ^ __funnelshift_r(FH[i], FH[i + 1], 1)) | (FL[i] ^ __funnelshift_r(FL[i], FL[i + 1], 1));
                        HD[T + 1] = (FH[T + 1] ^ __funnelshift_r(FH[T + 1], hx, 1)) | (FL[T + 1] ^ __funnelshift_r(FL[T + 1], lx, 1));
#pragma unroll
                        for (int sh = 1; sh <= 4; sh <<= 1) {
#pragma unroll
                            for (int i = 0; i <= T; ++i) HD[i] |= __funnelshift_r(HD[i], HD[i + 1], sh);
                            HD[T + 1] |= __funnelshift_r(HD[T + 1], 0xFFFFFFFFu, sh);
                        }
                        hd_level = 1;
                    }
                    if (sup == 2 && hd_level == 1) {    // widen the homopolymer mask from 8 to 16 matches
#pragma unroll
                        for (int i = 0; i <= T; ++i) HD[i] |= __funnelshift_r(HD[i], HD[i + 1], 8);
                        HD[T + 1] |= __funnelshift_r(HD[T + 1], 0xFFFFFFFFu, 8);
                        hd_level = 2;
                    }
                    if (sup == 0 && hd_level != 0) {    // (only with unusual filter settings: a BYTE / ERODE segment beyond k = 16)
#pragma unroll
                        for (int i = 0; i <= T + 1; ++i) HD[i] = 0xFFFFFFFFu;
                        hd_level = 0;
                    }
                    const uint32_t mode = sg.mode & 15u;
                    // (one loop per filter: without suppression HD is all ones, which the 3-input LOP3 of the compare absorbs for free, and
                    // an unused dilation step has shift 0 -- every extra specialisation of these loops costs more in instruction fetch
                    // than it saves in arithmetic, profiles/r02_kernel_iterations.md)
                    if (mode == MODE_BYTE) {
#pragma unroll 1
                        for (uint32_t s = s_lo, bit = 1u << s_lo; s <= s_hi; ++s, bit <<= 1)
                            set_bit_if(hitmask, filter_byte<T, true>(NH, NL, FH, FL, s, HD), bit);
                    } else {
#pragma unroll 1
                        for (uint32_t s = s_lo, bit = 1u << s_lo; s <= s_hi; ++s, bit <<= 1)
                            set_bit_if(hitmask, filter_erode<T, 3, true>(NH, NL, FH, FL, s, sh0, sh1, sh2, HD), bit);
                    }
                }
            }
            s_hit[nb * STRIPS + sub * THREADS + tid] = hitmask;
            if (tid == 0) s_misc[12 + nb] = qb;
        }
        }

        // ---- exact phase over the hits of these groups: compact (prefix sum, no atomics) ...
        uint32_t cnt[NS], incl[NS];
#pragma unroll
        for (int sub = 0; sub < NS; ++sub) {
            uint32_t c = 0;                       // hits of my strip `sub` (my own stores, read back)
            for (uint32_t b = 0; b < nb; ++b) c += __popc(s_hit[b * STRIPS + sub * THREADS + tid]);
            cnt[sub] = c;
            uint32_t in = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, in, o);
                if (lane >= (uint32_t)o) in += v;
            }
            incl[sub] = in;
            if (lane == 31) s_misc[16 * sub + warp] = in;
        }
        if (tid == 0) s_misc[10] = 0;  // start-queue fill
        __syncthreads();
        uint32_t total = 0;
#pragma unroll
        for (int sub = 0; sub < NS; ++sub) {
            uint32_t woff = 0, tsub = 0;
#pragma unroll
            for (uint32_t i = 0; i < THREADS / 32; ++i) {
                const uint32_t v = s_misc[16 * sub + i];
                if (i < warp) woff += v;
                tsub += v;
            }
            s_pre[sub * THREADS + tid] = total + woff + incl[sub] - cnt[sub];
            total += tsub;
        }
        __syncthreads();
        if (tid == 0) ncand += total;
        if (p.debug_flags & 1u) total = 0;  // profiling only: fast phase alone
        // ... one thread per (strip, k) hit: find the run starts, follow each run, emit
        for (uint32_t item = tid; item < total; item += THREADS) {
            uint32_t lo = 0, hi = STRIPS - 1;  // last strip whose exclusive prefix is <= item
            while (lo < hi) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (s_pre[mid] <= item) lo = mid; else hi = mid - 1;
            }
            uint32_t n = item - s_pre[lo], k = 0;
            for (uint32_t b = 0; b < nb; ++b) {
                const uint32_t mask = s_hit[b * STRIPS + lo];
                const uint32_t c = __popc(mask);
                if (n < c) { k = s_misc[12 + b] * 32 + __fns(mask, 0, n + 1); break; }
                n -= c;
            }
            if (tile_has_n) exact_item<true>(p, tc, w0 + lo * T, T, k);
            else exact_item<false>(p, tc, w0 + lo * T, T, k);
        }
        __syncthreads();
        // ... overflow queue (more than two run starts in one strip for one k)
        const uint32_t nst = min(s_misc[10], STARTQ_CAP);
        for (uint32_t e = tid; e < nst; e += THREADS) {
            const uint2 q = s_startq[e];
            const uint32_t k = q.y & 0xFFFFu;
            const KEntry ke = p.ktab[k];
            handle_start(p, tc, ke, k, q.x, q.y >> 16);
        }
        __syncthreads();
    }

    // ---- long runs: the whole block walks 256 words per step
    const uint32_t nlong = min(s_misc[9], (uint32_t)LONGCAP);
    for (uint32_t e = 0; e < nlong; ++e) {
        const uint32_t st = s_long[3 * e], k = s_long[3 * e + 1];
        uint32_t wcur = s_long[3 * e + 2], i0;
        for (;;) {
            if (tid == 0) s_misc[10] = NOPOS;
            __syncthreads();
            const uint32_t w = wcur + tid;
            const uint32_t m = (w < p.n_words) ? exact_mask(p, k, w) : 0u;
            if (m != 0xFFFFFFFFu) atomicMin(&s_misc[10], (w << 5) + (__ffs(~m) - 1));
            __syncthreads();
            i0 = s_misc[10];
            __syncthreads();
            if (i0 != NOPOS) break;
            wcur += THREADS;
        }
        if (tid == 0) emit_result(p, tc, st, i0 + k, k);
    }
    __syncthreads();

    // ---- order the tile's results by (start, end): counting sort over strips, then each strip's
    //      few results are put in order by one thread; written as one segment of the staging list
    const uint32_t n = min(s_misc[8], p.outcap);
#pragma unroll
    for (int sub = 0; sub < NS; ++sub) s_hit[sub * THREADS + tid] = 0;
    if (tid == 0) {
        const unsigned long long base = atomicAdd(p.counters + C_STAGE, (unsigned long long)n);
        s_misc[11] = (base + n <= p.stage_cap) ? (uint32_t)base : NOPOS;
        p.tile_cnt[tile] = n;
        p.tile_base[tile] = (uint32_t)base;
        if (nlong) atomicAdd(p.counters + C_LONG, (unsigned long long)nlong);
        if (ncand) atomicAdd(p.counters + C_CAND, ncand);
    }
    __syncthreads();
    const uint32_t base = s_misc[11];
    if (base == NOPOS || n == 0) return;  // (too small a result buffer: the host grows it and re-runs)
    for (uint32_t i = tid; i < n; i += THREADS) {
        const uint32_t bin = ((uint32_t)(s_key[i] >> 37) - w0) / T;  // strip of the start position
        atomicAdd(&s_hit[bin], 1u);
    }
    __syncthreads();
    {
        uint32_t c[NS], in[NS];
#pragma unroll
        for (int sub = 0; sub < NS; ++sub) {
            c[sub] = s_hit[sub * THREADS + tid];
            uint32_t v = c[sub];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
                if (lane >= (uint32_t)o) v += t;
            }
            in[sub] = v;
            if (lane == 31) s_misc[16 * sub + warp] = v;
        }
        __syncthreads();
        uint32_t run = 0;
#pragma unroll
        for (int sub = 0; sub < NS; ++sub) {
            uint32_t woff = 0, tsub = 0;
#pragma unroll
            for (uint32_t i = 0; i < THREADS / 32; ++i) {
                const uint32_t v = s_misc[16 * sub + i];
                if (i < warp) woff += v;
                tsub += v;
            }
            s_pre[sub * THREADS + tid] = run + woff + in[sub] - c[sub];  // cursor: start of this strip's slots
            run += tsub;
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += THREADS) {
        const uint64_t key = s_key[i];
        const uint32_t bin = ((uint32_t)(key >> 37) - w0) / T;
        const uint32_t pos = atomicAdd(&s_pre[bin], 1u);
        p.stage_key[base + pos] = key;
        p.stage_k[base + pos] = s_k[i];
    }
    __syncthreads();  // block-wide visibility of the staged rows
#pragma unroll 1
    for (int sub = 0; sub < NS; ++sub) {
        const uint32_t c = s_hit[sub * THREADS + tid];
        if (c > 1) {  // insertion sort of this strip's rows (in place, by (key, k))
            const uint32_t first = base + s_pre[sub * THREADS + tid] - c;
            for (uint32_t a = 1; a < c; ++a) {
                const uint64_t key = p.stage_key[first + a];
                const uint16_t kk = p.stage_k[first + a];
                uint32_t b = a;
                while (b > 0) {
                    const uint64_t kb = p.stage_key[first + b - 1];
                    const uint16_t vb = p.stage_k[first + b - 1];
                    if (kb < key || (kb == key && vb <= kk)) break;
                    p.stage_key[first + b] = kb;
                    p.stage_k[first + b] = vb;
                    --b;
                }
                p.stage_key[first + b] = key;
                p.stage_k[first + b] = kk;
            }
        }
    }
}

}  // namespace crf
