// crf_device.cuh -- device-side data model and exact match-mask primitives (sm_100a).
//
// Data layout in HBM (see DESIGN.md "Data layout"): every record of a load is laid out in one
// position space ("layout"), records separated by >= max_motif_cap masked positions so that no
// compare S[j] vs S[j+k] can pair two records.  Position p lives in word p>>5, bit p&31 of
//   H, L : the two bits of the base code (A=00 C=01 G=10 T=11)            -- 0.25 B/bp
//   NM   : 1 = not one of A,C,G,T (N, gap, any other symbol)              -- 0.125 B/bp
//   X    : 1 = "exotic" symbol: neither ACGT nor N (IUPAC codes ...), read only by the exact path
// Masked positions carry a filler code in H/L: a position hash for N/gap (aperiodic, so the
// unmasked H/L compare has no long runs inside N blocks), a letter hash for exotic symbols
// (equal letters get equal codes, so the unmasked compare is a superset of the true mask).
//
// The match mask the reference's tracker evaluates one base at a time
// (utils/perfect_repeat_tracker.py:53:  seq[i] == seq[i+k] and seq[i] != "N") is, for 32
// positions at once,  M_k[w] = ~((H^H>>k) | (L^L>>k)) & ~(NM | NM>>k)  (+ exotic equal letters).
#pragma once
#include <stdint.h>

namespace crf {

constexpr int THREADS = 256;     // threads per scan block
constexpr int LONGCAP = 32;      // per-tile queue of runs handed to the block-cooperative walker
constexpr uint32_t NOPOS = 0xFFFFFFFFu;

enum FilterMode : uint8_t {
    MODE_ERODE = 0,  // r_min < 15 : dilate the mismatch word by min(r_min, 8)
    MODE_BYTE = 1,   // r_min >= 15: a run contains an aligned all-match byte
    MODE_HALF = 2,   // r_min >= 31: ... an aligned all-match half-word (H plane only)
    MODE_WORD = 3,   // r_min >= 63: ... an aligned all-match word (H plane only)
};

// Per motif size k (index = k).  r_min = max(min_span - k, (min_repeats-1)*k) is the number of
// consecutive ones of M_k a run needs (trk:86,91 in closed form).
// Exact phase: a word of M_k is eroded by re = min(r_min, 32) with the shift schedule esh (unused
// steps are 0), so only positions followed by >= re matches survive.
struct __align__(16) KEntry {
    uint32_t rmin;
    uint8_t mode;      // FilterMode
    uint8_t sh[3];     // dilation shifts of the ERODE filter (0 = unused)
    uint8_t esh[5];    // erosion shifts of the exact phase
    uint8_t re;        // min(rmin, 32)
    uint16_t div[6];   // k/p for the distinct primes p | k (0-terminated): primitivity, trk:108-142
    uint8_t pad_[6];
};
static_assert(sizeof(KEntry) == 32, "KEntry layout");

// A maximal range of motif sizes that share (k >> 5) and the fast-phase filter.
struct Seg {
    uint16_t k_lo, k_hi;
    uint8_t mode, sh0, sh1, sh2;  // mode: FilterMode | (suppression level << 4)
};
static_assert(sizeof(Seg) == 8, "Seg layout");

struct ScanParams {
    const uint32_t *H, *L, *NM, *X;
    const KEntry *ktab;
    const Seg *segs;
    uint32_t n_segs;
    const uint64_t *ex_key;  // sorted (pos << 8 | letter) of exotic symbols
    uint32_t n_exotic;
    const uint32_t *rec_dev_off;  // layout position of each record's first base
    const uint32_t *own_lo, *own_hi;  // per record: layout range of run starts this load owns
    uint32_t n_records;
    uint32_t n_words;        // words holding layout positions (reads beyond are allocated pads)
    uint32_t kmin, kmax;
    uint32_t outcap;         // per-tile sorted-output slots
    uint32_t walk_limit;     // words one thread walks before the block takes over
    uint32_t sup_enabled;    // some segment carries a homopolymer-suppression level
    uint32_t debug_flags;    // bit 0: skip the exact phase (profiling only, results are then empty)
    uint64_t *stage_key;     // (start << 32 | end), tile-sorted segments
    uint16_t *stage_k;
    uint32_t stage_cap;
    uint32_t *tile_cnt, *tile_base;
    uint64_t *spill_key;     // results that overflowed a tile's slots (unsorted)
    uint16_t *spill_k;
    uint32_t spill_cap;
    unsigned long long *counters;  // see CounterIdx
};

enum CounterIdx { C_STAGE = 0, C_SPILL = 1, C_LONG = 2, C_CAND = 3, C_TOTAL = 4, C_TILE = 5, C_OPEN = 6, C_COUNT = 8 };

__device__ __forceinline__ uint32_t plane_at(const uint32_t *__restrict__ P, uint32_t w, uint32_t q, uint32_t s) {
    return __funnelshift_r(__ldg(P + w + q), __ldg(P + w + q + 1), s);
}

// letter of the exotic symbol at layout position pos (0 if none)
__device__ inline uint32_t exotic_letter(const uint64_t *__restrict__ ex_key, uint32_t n_exotic, uint32_t pos) {
    uint32_t lo = 0, hi = n_exotic;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if ((uint32_t)(ex_key[mid] >> 8) < pos) lo = mid + 1; else hi = mid;
    }
    if (lo < n_exotic && (uint32_t)(ex_key[lo] >> 8) == pos) return (uint32_t)(ex_key[lo] & 0xFF);
    return 0;
}

// Exact M_k for the 32 positions of word w, from global memory.  Out of line on purpose (plain arguments, so that the
// kernel's parameter block need not be copied to local memory): the scan kernel reaches it only where a run leaves the
// staged tile or the load has exotic symbols, but it is referenced from every walk / primitivity site of the exact phase --
// inlined, its copies (exotic look-up included) were 40 % of the kernel's code, and the kernel is sensitive to its
// instruction footprint (profiles/r02_kernel_iterations.md).
__device__ __noinline__ uint32_t exact_mask_global(const uint32_t *__restrict__ H, const uint32_t *__restrict__ L,
                                                   const uint32_t *__restrict__ NM, const uint32_t *__restrict__ X,
                                                   const uint64_t *__restrict__ ex_key, uint32_t n_exotic, uint32_t k, uint32_t w) {
    const uint32_t q = k >> 5, s = k & 31;
    const uint32_t dh = __ldg(H + w) ^ plane_at(H, w, q, s);
    const uint32_t dl = __ldg(L + w) ^ plane_at(L, w, q, s);
    const uint32_t nm = __ldg(NM + w) | plane_at(NM, w, q, s);
    const uint32_t eq = ~(dh | dl);
    uint32_t m = eq & ~nm;
    if (n_exotic) {  // equal exotic letters match too (only "N" is special, trk:53)
        uint32_t x = __ldg(X + w) & plane_at(X, w, q, s) & eq;
        while (x) {
            const uint32_t b = __ffs(x) - 1;
            x &= x - 1;
            const uint32_t pos = (w << 5) + b;
            if (exotic_letter(ex_key, n_exotic, pos) == exotic_letter(ex_key, n_exotic, pos + k)) m |= 1u << b;
        }
    }
    return m;
}
__device__ __forceinline__ uint32_t exact_mask(const ScanParams &p, uint32_t k, uint32_t w) {
    return exact_mask_global(p.H, p.L, p.NM, p.X, p.ex_key, p.n_exotic, k, w);
}

// First position >= from with M_k == 0, looking at no more than `limit` further words.
// Returns true and the position; or false with *i0 = first position not yet examined.
__device__ inline bool walk_run(const ScanParams &p, uint32_t k, uint32_t from, uint32_t limit, uint32_t *i0) {
    uint32_t w = from >> 5;
    uint32_t m = exact_mask(p, k, w) | ((1u << (from & 31)) - 1u);
    for (uint32_t n = 0;; ++n) {
        if (m != 0xFFFFFFFFu) { *i0 = (w << 5) + (__ffs(~m) - 1); return true; }
        ++w;
        if (n >= limit || w >= p.n_words) {  // beyond n_words everything is masked
            *i0 = w << 5;
            return w >= p.n_words;
        }
        m = exact_mask(p, k, w);
    }
}

// M_d[a .. a+len) all ones?
__device__ inline bool all_match(const ScanParams &p, uint32_t d, uint32_t a, uint32_t len) {
    uint32_t pos = a;
    const uint32_t end = a + len;
    while (pos < end) {
        const uint32_t b = pos & 31;
        const uint32_t n = min(32u - b, end - pos);
        const uint32_t mask = (n == 32 ? 0xFFFFFFFFu : ((1u << n) - 1u)) << b;
        if ((exact_mask(p, d, pos >> 5) & mask) != mask) return false;
        pos += n;
    }
    return true;
}

// The motif S[st:st+k] is primitive iff it is not u^n, n >= 2 (consists_of_perfect_repeats,
// trk:108-142).  u^n with |u| = d | k  <=>  S[j] == S[j+d] on [st, st+k-d); it suffices to try
// d = k/p for the primes p | k.
__device__ inline bool motif_is_primitive(const ScanParams &p, const KEntry &ke, uint32_t k, uint32_t st) {
#pragma unroll 1
    for (int j = 0; j < 6; ++j) {
        const uint32_t d = ke.div[j];
        if (!d) break;
        if (all_match(p, d, st, k - d)) return false;
    }
    return true;
}

__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

}  // namespace crf
