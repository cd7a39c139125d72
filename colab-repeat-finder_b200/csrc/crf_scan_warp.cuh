// crf_scan_warp.cuh -- the scan kernel with WARP-owned tiles: no block-wide barrier anywhere.
//
// Same algorithm, same filters and the same exact phase as scan_kernel (crf_scan.cuh; reference:
// perfect_repeat_finder.py:66-74 and utils/perfect_repeat_tracker.py:43-101), re-scheduled:
//
//   * the grid is persistent (as many CTAs as fit the GPU); every WARP fetches its next tile -- NS sub-tiles of 32*T words,
//     one strip of T words per lane and sub-tile -- from a global counter, so the SMs stay busy to the last tile;
//   * a warp stages its tile (+ 1 word of left context, + the kmax halo) in its own slice of shared memory, runs the fast
//     phase on every sub-tile, compacts the (strip, k) hits with shuffles and works them off 32 at a time (exact_item /
//     handle_start, unchanged), finishes long runs 32 words per step, orders its rows and writes them as one segment;
//   * warps never wait for each other: a warp whose tile has many candidates does not hold up seven others at a barrier
//     (the block-tiled kernel lost ~20 % of its issue slots there), and the load latency of one warp's staging is covered
//     by the other warps' arithmetic.
// Downstream (tile_offsets / gather / translate) is unchanged: a "tile" is now 32*T*NS words.
#pragma once
#include "crf_scan.cuh"

namespace crf {

constexpr uint32_t WARP_STARTQ = 32;   // overflow queue: run starts beyond two per (strip, k) hit
constexpr uint32_t WARP_LONGQ = 8;     // runs handed to the warp-cooperative walker
constexpr uint32_t WARP_NBATCH = 2;    // groups of 32 motif sizes whose hits share one exact phase
constexpr int WARPS_PER_CTA = 4;

__host__ __device__ inline uint32_t warp_tile_words(int T, int NS, uint32_t kmax) {
    return 32u * T * NS + (kmax >> 5) + 5;  // 1 word of left context + tile + halo (q + 4)
}
// shared memory one warp needs (bytes, multiple of 16)
__host__ __device__ inline size_t warp_smem_bytes(int T, int NS, uint32_t kmax, uint32_t outcap) {
    const size_t plane = pad_idx(warp_tile_words(T, NS, kmax)) + 1;
    size_t words = 3 * plane + (WARP_NBATCH + 1) * 32 * NS + 16 + 3 * WARP_LONGQ;
    words = (words + 1) & ~(size_t)1;
    size_t bytes = words * 4 + (size_t)WARP_STARTQ * 8 + (size_t)outcap * 8 + (((size_t)outcap * 2 + 7) & ~(size_t)7);
    return (bytes + 15) & ~(size_t)15;
}

template <int T, int NS>
__global__ void __launch_bounds__(32 * WARPS_PER_CTA, NS == 1 ? 8 : 6) scan_warp_kernel(const ScanParams p) {
    constexpr uint32_t STRIPS = 32 * NS;           // strips (= lanes x sub-tiles) of one tile
    constexpr uint32_t TW = 32 * T * NS;           // words of one tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nsm = warp_tile_words(T, NS, p.kmax);
    const uint32_t plane = pad_idx(nsm) + 1;
    unsigned char *my = smem_raw + (size_t)warp * warp_smem_bytes(T, NS, p.kmax, p.outcap);

    uint32_t *sH = reinterpret_cast<uint32_t *>(my);
    uint32_t *sL = sH + plane;
    uint32_t *sN = sL + plane;
    uint32_t *s_hit = sN + plane;                  // WARP_NBATCH x STRIPS hit masks (later: strip histogram)
    uint32_t *s_pre = s_hit + WARP_NBATCH * STRIPS;
    uint32_t *s_misc = s_pre + STRIPS;             // [8] nout [9] nlong [10] nstart [12..] q of each batch group
    uint32_t *s_long = s_misc + 16;
    size_t off_words = 3 * (size_t)plane + (WARP_NBATCH + 1) * STRIPS + 16 + 3 * WARP_LONGQ;
    off_words = (off_words + 1) & ~(size_t)1;
    uint2 *s_startq = reinterpret_cast<uint2 *>(my + off_words * 4);
    uint64_t *s_key = reinterpret_cast<uint64_t *>(s_startq + WARP_STARTQ);
    uint16_t *s_k = reinterpret_cast<uint16_t *>(s_key + p.outcap);

    const uint32_t n_tiles = (p.n_words + TW - 1) / TW;
    unsigned long long ncand = 0, nlong_total = 0;

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = (uint32_t)atomicAdd(p.counters + C_TILE, 1ull);
        tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
        if (tile >= n_tiles) break;
        const uint32_t w0 = tile * TW;

        // ---- stage: smem index j <-> absolute word w0 - 1 + j
        uint32_t any_mask = 0;
        for (uint32_t j = lane; j < nsm; j += 32) {
            const bool real = (w0 + j) != 0;
            const uint32_t w = w0 + j - 1;
            const uint32_t a = pad_idx(j);
            sH[a] = real ? __ldg(p.H + w) : 0u;
            sL[a] = real ? __ldg(p.L + w) : 0u;
            const uint32_t nmw = real ? __ldg(p.NM + w) : 0xFFFFFFFFu;
            sN[a] = nmw;
            any_mask |= nmw;
        }
        if (lane < 16) s_misc[lane] = 0;
        const bool tile_has_n = __any_sync(0xFFFFFFFFu, any_mask != 0);   // also orders the staging stores
        __syncwarp();

        TileCtx tc;
        tc.sH = sH; tc.sL = sL; tc.sN = sN; tc.wbase = w0; tc.nsm = nsm;
        tc.key = s_key; tc.kk = s_k; tc.nout = &s_misc[8]; tc.nlong = &s_misc[9]; tc.longq = s_long; tc.cap = p.outcap;
        tc.startq = s_startq; tc.nstart = &s_misc[10];
        tc.longcap = WARP_LONGQ; tc.startcap = WARP_STARTQ;

        uint32_t si = 0;
        while (si < p.n_segs) {
            // ---- fast phase: up to WARP_NBATCH groups of <= 32 motif sizes, every sub-tile
            const uint32_t si0 = si;
            uint32_t nb = 0;
#pragma unroll 1
            for (int sub = 0; sub < NS; ++sub) {
                const uint32_t sbase = 1 + (sub * 32 + lane) * T;      // smem index of my strip's first word
                uint32_t NH[T + 1], NL[T + 1];
#pragma unroll
                for (int i = 0; i <= T; ++i) {
                    NH[i] = sH[pad_idx(sbase + i)];
                    NL[i] = sL[pad_idx(sbase + i)];
                }
                si = si0;
                for (nb = 0; nb < WARP_NBATCH && si < p.n_segs; ++nb) {
                    const uint32_t qb = (uint32_t)p.segs[si].k_lo >> 5;
                    uint32_t FH[T + 2], FL[T + 2];
#pragma unroll
                    for (int i = 0; i <= T + 1; ++i) {
                        FH[i] = sH[pad_idx(sbase + qb + i)];
                        FL[i] = sL[pad_idx(sbase + qb + i)];
                    }
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
                    s_hit[nb * STRIPS + sub * 32 + lane] = hitmask;
                    if (lane == 0) s_misc[12 + nb] = qb;
                }
            }

            // ---- exact phase over the hits of these groups: compact with shuffles ...
            uint32_t total = 0;
#pragma unroll
            for (int sub = 0; sub < NS; ++sub) {
                uint32_t c = 0;                                        // my strip's hits (read back: my own stores)
                for (uint32_t b = 0; b < nb; ++b) c += __popc(s_hit[b * STRIPS + sub * 32 + lane]);
                uint32_t incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= (uint32_t)o) incl += v;
                }
                s_pre[sub * 32 + lane] = total + incl - c;
                total += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            if (lane == 0) s_misc[10] = 0;  // start-queue fill
            __syncwarp();
            if (lane == 0) ncand += total;
            if (p.debug_flags & 1u) total = 0;  // profiling only: fast phase alone
            // ... one lane per (strip, k) hit: find the run starts, follow each run, emit
            for (uint32_t base = 0; base < total; base += 32) {
                const uint32_t item = base + lane;
                if (item < total) {
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
                __syncwarp();
            }
            // ... overflow queue (more than two run starts in one strip for one k)
            const uint32_t nst = min(s_misc[10], WARP_STARTQ);
            for (uint32_t e = lane; e < nst; e += 32) {
                const uint2 q = s_startq[e];
                const uint32_t k = q.y & 0xFFFFu;
                const KEntry ke = p.ktab[k];
                handle_start(p, tc, ke, k, q.x, q.y >> 16);
            }
            __syncwarp();
        }

        // ---- long runs: the warp walks 32 words per step
        const uint32_t nlong = min(s_misc[9], WARP_LONGQ);
        for (uint32_t e = 0; e < nlong; ++e) {
            const uint32_t st = s_long[3 * e], k = s_long[3 * e + 1];
            uint32_t wcur = s_long[3 * e + 2], i0;
            for (;;) {
                const uint32_t w = wcur + lane;
                const uint32_t m = (w < p.n_words) ? exact_mask(p, k, w) : 0u;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, m != 0xFFFFFFFFu);
                if (bal) {
                    const uint32_t first = __ffs(bal) - 1;
                    const uint32_t mm = __shfl_sync(0xFFFFFFFFu, m, first);
                    i0 = ((wcur + first) << 5) + (__ffs(~mm) - 1);
                    break;
                }
                wcur += 32;
            }
            if (lane == 0) emit_result(p, tc, st, i0 + k, k);
        }
        nlong_total += nlong;
        __syncwarp();

        // ---- order the tile's rows by (start, end): counting sort over strips, then each strip's few rows are put in
        //      order by one lane; written as one segment of the staging list
        const uint32_t n = min(s_misc[8], p.outcap);
        uint32_t base = 0;
        if (lane == 0) {
            const unsigned long long b64 = atomicAdd(p.counters + C_STAGE, (unsigned long long)n);
            base = (b64 + n <= p.stage_cap) ? (uint32_t)b64 : NOPOS;
            p.tile_cnt[tile] = n;
            p.tile_base[tile] = (uint32_t)b64;
        }
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base == NOPOS || n == 0) continue;  // (too small a result buffer: the host grows it and re-runs)
#pragma unroll
        for (int sub = 0; sub < NS; ++sub) s_hit[sub * 32 + lane] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t bin = ((uint32_t)(s_key[i] >> 37) - w0) / T;  // strip of the start position
            atomicAdd(&s_hit[bin], 1u);
        }
        __syncwarp();
        {
            uint32_t run = 0;
#pragma unroll
            for (int sub = 0; sub < NS; ++sub) {
                const uint32_t c = s_hit[sub * 32 + lane];
                uint32_t incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= (uint32_t)o) incl += v;
                }
                s_pre[sub * 32 + lane] = run + incl - c;  // cursor: start of this strip's slots
                run += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
        }
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) {
            const uint64_t key = s_key[i];
            const uint32_t bin = ((uint32_t)(key >> 37) - w0) / T;
            const uint32_t pos = atomicAdd(&s_pre[bin], 1u);
            p.stage_key[base + pos] = key;
            p.stage_k[base + pos] = s_k[i];
        }
        __syncwarp();  // warp-wide visibility of the staged rows
#pragma unroll 1
        for (int sub = 0; sub < NS; ++sub) {
            const uint32_t c = s_hit[sub * 32 + lane];
            if (c > 1) {  // insertion sort of this strip's rows (in place, by (key, k))
                const uint32_t first = base + s_pre[sub * 32 + lane] - c;
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
        __syncwarp();  // the next tile overwrites this warp's shared memory
    }
    if (lane == 0) {
        if (nlong_total) atomicAdd(p.counters + C_LONG, nlong_total);
        if (ncand) atomicAdd(p.counters + C_CAND, ncand);
    }
}

}  // namespace crf
