// crf_scan.cuh -- the scan kernel: every motif size k in [kmin, kmax] over one shared-memory tile.
//
// Replaces the reference's lock-step loop (perfect_repeat_finder.py:66-74) and the tracker state
// machine (utils/perfect_repeat_tracker.py:43-101).  Closed form being computed (SURVEY App. A.2,
// checked against the oracle by tests/): for each k, every maximal run [st, i0) of ones of
//   M_k[j] = (S[j] == S[j+k]) and S[j] != 'N'
// with i0 - st >= r_min(k) = max(min_span - k, (min_repeats-1)*k) and a primitive motif S[st:st+k]
// is reported as (st, i0 + k, k).
//
// One CTA owns one tile of 256*T words (T*8192 bases) and emits the runs that START in it:
//   1. stage the tile (+ a halo of kmax bases) of the H and L planes in shared memory;
//   2. fast phase -- each thread keeps a strip of T(+1) consecutive words in registers and, for
//      every k, forms the shifted-compare words with funnel shifts and tests a *necessary*
//      condition for "a qualifying run starts in my strip" (FilterMode); no branches, the
//      result is one bit per (strip, k) in a register;
//   3. exact phase -- the block compacts the (strip, k) hits (prefix sum, no atomics), one
//      thread per hit re-evaluates the exact mask (N mask, exotic symbols), finds the run
//      starts by erosion, walks right to the run end, applies the thresholds and the
//      primitivity rule and appends (start, end, k) to the tile's shared-memory result list;
//   4. runs longer than walk_limit words are finished by the whole block (256 words per step);
//   5. the tile's results are rank-sorted by (start, end) and written as one segment; a later
//      pass concatenates the segments in tile order, which is global (start, end) order because
//      a run is owned by the tile of its start.
#pragma once
#include "crf_device.cuh"

namespace crf {

struct TileOut {
    uint64_t *key;
    uint16_t *kk;
    uint32_t *nout;
    uint32_t *nlong;
    uint32_t *longq;
    uint32_t cap;
};

__device__ inline void emit_result(const ScanParams &p, const TileOut &to, uint32_t st, uint32_t end, uint32_t k) {
    const uint64_t key = ((uint64_t)st << 32) | end;
    const uint32_t idx = atomicAdd(to.nout, 1u);
    if (idx < to.cap) {
        to.key[idx] = key;
        to.kk[idx] = (uint16_t)k;
    } else {
        const unsigned long long g = atomicAdd(p.counters + C_SPILL, 1ull);
        if (g < p.spill_cap) {
            p.spill_key[g] = key;
            p.spill_k[g] = (uint16_t)k;
        }
    }
}

// A qualifying-run start candidate: M_k[st .. st+known) are ones and M_k[st-1] is zero.
__device__ inline void handle_start(const ScanParams &p, const TileOut &to, const KEntry &ke, uint32_t k,
                                    uint32_t st, uint32_t known) {
    if (p.own_lo) {  // partitioned load: a run belongs to the unit that owns its start
        uint32_t lo = 0, hi = p.n_records - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (__ldg(p.rec_dev_off + mid) <= st) lo = mid; else hi = mid - 1;
        }
        if (st < __ldg(p.own_lo + lo) || st >= __ldg(p.own_hi + lo)) return;
    }
    uint32_t i0;
    bool found = walk_run(p, k, st + known, p.walk_limit, &i0);
    while (!found && i0 - st < ke.rmin) found = walk_run(p, k, i0, p.walk_limit, &i0);
    if (found && i0 - st < ke.rmin) return;                    // trk:86/91 thresholds
    if (!motif_is_primitive(p, ke, k, st)) return;              // trk:98
    if (found) {
        emit_result(p, to, st, i0 + k, k);                      // end = i0 + k (trk:87-92)
        return;
    }
    const uint32_t slot = atomicAdd(to.nlong, 1u);
    if (slot < LONGCAP) {
        to.longq[3 * slot] = st;
        to.longq[3 * slot + 1] = k;
        to.longq[3 * slot + 2] = i0 >> 5;
    } else {  // queue full: finish it alone (correct, slow, practically never)
        while (!walk_run(p, k, i0, 0x7FFFFFFFu, &i0)) {}
        emit_result(p, to, st, i0 + k, k);
    }
}

// Exact phase for one (strip, k) hit: all qualifying run starts inside the strip's nw words.
__device__ inline void exact_item(const ScanParams &p, const TileOut &to, uint32_t wfirst, int nw, uint32_t k) {
    const KEntry ke = p.ktab[k];
    const uint32_t rex = ke.rexact;
    uint32_t prevtop = wfirst ? (exact_mask(p, k, wfirst - 1) >> 31) : 0u;
    uint32_t cur = exact_mask(p, k, wfirst);
    for (int i = 0; i < nw; ++i) {
        const uint32_t w = wfirst + i;
        const uint32_t nxt = exact_mask(p, k, w + 1);
        if (cur) {
            uint64_t v = ((uint64_t)nxt << 32) | cur;  // erode by rex: bit j <- bits j..j+rex-1 all set
            for (uint32_t covered = 1; covered < rex;) {
                const uint32_t sh = min(covered, rex - covered);
                v &= v >> sh;
                covered += sh;
            }
            uint32_t starts = (uint32_t)v & ~((cur << 1) | prevtop);
            while (starts) {
                const uint32_t b = __ffs(starts) - 1;
                starts &= starts - 1;
                handle_start(p, to, ke, k, (w << 5) + b, rex);
            }
        }
        prevtop = cur >> 31;
        cur = nxt;
    }
}

// ---- fast-phase filters (T+1 words: the strip plus one look-ahead word) ------------------------
template <int T>
__device__ __forceinline__ bool filter_word(const uint32_t (&NH)[T + 1], const uint32_t (&FH)[T + 2], uint32_t s) {
    bool hit = false;
#pragma unroll
    for (int i = 0; i <= T; ++i) hit |= (NH[i] == __funnelshift_r(FH[i], FH[i + 1], s));
    return hit;
}
template <int T>
__device__ __forceinline__ bool filter_half(const uint32_t (&NH)[T + 1], const uint32_t (&FH)[T + 2], uint32_t s) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i <= T; ++i) {
        const uint32_t x = NH[i] ^ __funnelshift_r(FH[i], FH[i + 1], s);
        acc |= (x - 0x00010001u) & ~x;
    }
    return (acc & 0x80008000u) != 0;
}
template <int T>
__device__ __forceinline__ bool filter_byte(const uint32_t (&NH)[T + 1], const uint32_t (&NL)[T + 1],
                                            const uint32_t (&FH)[T + 2], const uint32_t (&FL)[T + 2], uint32_t s) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i <= T; ++i) {
        const uint32_t x = (NH[i] ^ __funnelshift_r(FH[i], FH[i + 1], s)) |
                           (NL[i] ^ __funnelshift_r(FL[i], FL[i + 1], s));
        acc |= (x - 0x01010101u) & ~x;
    }
    return (acc & 0x80808080u) != 0;
}
template <int T>
__device__ __forceinline__ bool filter_erode(const uint32_t (&NH)[T + 1], const uint32_t (&NL)[T + 1],
                                             const uint32_t (&FH)[T + 2], const uint32_t (&FL)[T + 2], uint32_t s,
                                             uint32_t sh0, uint32_t sh1, uint32_t sh2) {
    uint32_t x[T + 1];  // mismatch words; dilating mismatches == eroding matches
#pragma unroll
    for (int i = 0; i <= T; ++i)
        x[i] = (NH[i] ^ __funnelshift_r(FH[i], FH[i + 1], s)) | (NL[i] ^ __funnelshift_r(FL[i], FL[i + 1], s));
    if (sh0) {
#pragma unroll
        for (int i = 0; i < T; ++i) x[i] |= __funnelshift_r(x[i], x[i + 1], sh0);
        x[T] |= x[T] >> sh0;
    }
    if (sh1) {
#pragma unroll
        for (int i = 0; i < T; ++i) x[i] |= __funnelshift_r(x[i], x[i + 1], sh1);
        x[T] |= x[T] >> sh1;
    }
    if (sh2) {
#pragma unroll
        for (int i = 0; i < T; ++i) x[i] |= __funnelshift_r(x[i], x[i + 1], sh2);
    }
    uint32_t all = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < T; ++i) all &= x[i];
    return all != 0xFFFFFFFFu;
}

// dynamic shared memory a scan block needs
__host__ __device__ inline size_t scan_smem_bytes(int T, uint32_t kmax, uint32_t outcap) {
    const size_t tile_words = (size_t)THREADS * T + (kmax >> 5) + 3;
    size_t words = 2 * tile_words + 2 * THREADS + 16 + 3 * LONGCAP;
    words = (words + 1) & ~(size_t)1;
    return words * 4 + (size_t)outcap * 8 + (((size_t)outcap * 2 + 7) & ~(size_t)7);
}

template <int T>
__global__ void __launch_bounds__(THREADS) scan_kernel(const ScanParams p) {
    constexpr int TW = THREADS * T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t qmin = p.kmin >> 5, qmax = p.kmax >> 5;
    const uint32_t tile_words = TW + qmax + 3;

    uint32_t *sH = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *sL = sH + tile_words;
    uint32_t *s_hit = sL + tile_words;
    uint32_t *s_pre = s_hit + THREADS;
    uint32_t *s_misc = s_pre + THREADS;  // [0..7] warp sums [8] nout [9] nlong [10] minpos [11] base
    uint32_t *s_long = s_misc + 16;
    size_t off_words = 2 * (size_t)tile_words + 2 * THREADS + 16 + 3 * LONGCAP;
    off_words = (off_words + 1) & ~(size_t)1;
    uint64_t *s_key = reinterpret_cast<uint64_t *>(smem_raw + off_words * 4);
    uint16_t *s_k = reinterpret_cast<uint16_t *>(s_key + p.outcap);

    const uint32_t tile = blockIdx.x;
    const uint32_t w0 = tile * TW;
    for (uint32_t i = tid; i < tile_words; i += THREADS) {
        sH[i] = __ldg(p.H + w0 + i);
        sL[i] = __ldg(p.L + w0 + i);
    }
    if (tid < 16) s_misc[tid] = 0;
    __syncthreads();

    const TileOut to{s_key, s_k, &s_misc[8], &s_misc[9], s_long, p.outcap};

    uint32_t NH[T + 1], NL[T + 1];
#pragma unroll
    for (int i = 0; i <= T; ++i) {
        NH[i] = sH[tid * T + i];
        NL[i] = sL[tid * T + i];
    }

    unsigned long long ncand = 0;
    for (uint32_t qb = qmin; qb <= qmax; ++qb) {
        uint32_t FH[T + 2], FL[T + 2];
#pragma unroll
        for (int i = 0; i <= T + 1; ++i) {
            FH[i] = sH[tid * T + qb + i];
            FL[i] = sL[tid * T + qb + i];
        }
        const uint32_t s_lo = (qb == qmin) ? (p.kmin & 31) : 0u;
        const uint32_t s_hi = (qb == qmax) ? (p.kmax & 31) : 31u;
        uint32_t hitmask = 0;
#pragma unroll 1
        for (uint32_t s = s_lo; s <= s_hi; ++s) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(p.ktab + (qb * 32 + s)));
            const uint32_t mode = raw.z & 0xFF;
            bool hit;
            if (mode == MODE_WORD) hit = filter_word<T>(NH, FH, s);
            else if (mode == MODE_HALF) hit = filter_half<T>(NH, FH, s);
            else if (mode == MODE_BYTE) hit = filter_byte<T>(NH, NL, FH, FL, s);
            else hit = filter_erode<T>(NH, NL, FH, FL, s, (raw.z >> 8) & 0xFF, (raw.z >> 16) & 0xFF, raw.z >> 24);
            hitmask |= (hit ? 1u : 0u) << s;
        }

        // ---- compact the hits of this group of <= 32 motif sizes and run the exact phase
        const uint32_t cnt = __popc(hitmask);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        if (lane == 31) s_misc[warp] = incl;
        s_hit[tid] = hitmask;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (uint32_t i = 0; i < THREADS / 32; ++i) {
            const uint32_t v = s_misc[i];
            if (i < warp) woff += v;
            total += v;
        }
        s_pre[tid] = woff + incl - cnt;
        __syncthreads();
        if (tid == 0) ncand += total;
        for (uint32_t item = tid; item < total; item += THREADS) {
            uint32_t lo = 0, hi = THREADS - 1;  // last strip whose exclusive prefix is <= item
            while (lo < hi) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (s_pre[mid] <= item) lo = mid; else hi = mid - 1;
            }
            uint32_t mask = s_hit[lo];
            for (uint32_t n = item - s_pre[lo]; n; --n) mask &= mask - 1;
            const uint32_t k = qb * 32 + (__ffs(mask) - 1);
            exact_item(p, to, w0 + lo * T, T, k);
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
        if (tid == 0) emit_result(p, to, st, i0 + k, k);
    }
    __syncthreads();

    // ---- rank-sort the tile's results by (start, end) and write them as one segment
    const uint32_t n = min(s_misc[8], p.outcap);
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
    if (base == NOPOS) return;  // result buffer too small: the host grows it and re-runs
    for (uint32_t i = tid; i < n; i += THREADS) {
        const uint64_t key = s_key[i];
        const uint32_t kk = s_k[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; ++j) {
            const uint64_t kj = s_key[j];
            rank += (kj < key) || (kj == key && s_k[j] < kk);
        }
        p.stage_key[base + rank] = key;
        p.stage_k[base + rank] = (uint16_t)kk;
    }
}

}  // namespace crf
