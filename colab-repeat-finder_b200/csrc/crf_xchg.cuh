// crf_xchg.cuh -- gathering the compacted rows of N GPUs on rank 0 ("root") over NVLink peer memory.
//
// The reference scales out by fanning intervals over CPU jobs and concatenating their BED files
// (hail_batch_pipeline/run_hail_batch_pipeline.py:76-77, 115-123, 153).  Here every GPU scans a contiguous
// range of (record, chunk) units, so rank order is genome order and the final list is the concatenation of
// the ranks' sorted rows.  That concatenation is the only exchange of the path and it is done by the GPUs
// themselves, without a host round trip and without a collective library:
//
//   * every rank owns one exchange block in its HBM (XchgBlock + on the root the row buffer), mapped into
//     every peer's address space (CUDA IPC between processes, plain peer access inside one process);
//   * publish_kernel (one warp per rank and step): publishes this rank's row count to every peer with one
//     8-byte store that carries the step number (so the store is its own flag) and waits for the counts of the
//     ranks before it -- they arrive in its OWN memory; push_kernel then stores the rank's rows straight into the
//     root's buffer at the prefix offset (coalesced 4-byte peer stores) and its last block posts a done word there;
//   * settle_kernel (one warp): waits until the counts of ALL ranks (on the root also the done words) of this
//     step have arrived and writes status / totals for the host.
// Two slots per rank (step parity) are enough: no rank can be two steps ahead of another one, because a step
// cannot complete on the root before every rank has pushed, and no rank > 0 can push before the root has
// published its count of that step.
#pragma once
#include "crf_aux.cuh"

namespace crf {

constexpr uint32_t XCHG_MAX_WORLD = 16;
constexpr uint32_t XCHG_RESULT_WORDS = 8 + XCHG_MAX_WORLD;

// slot word: [63:40] step (24 bits)  [39] void  [38] the rank has open-ended rows  [31:0] value
constexpr unsigned long long XCHG_VOID = 1ull << 39;
constexpr unsigned long long XCHG_HAS_OPEN = 1ull << 38;
__host__ __device__ __forceinline__ unsigned long long xchg_enc(uint32_t step, bool is_void, uint32_t value, bool has_open = false) {
    return ((unsigned long long)(step & 0xFFFFFFu) << 40) | (is_void ? XCHG_VOID : 0ull) | (has_open ? XCHG_HAS_OPEN : 0ull) | value;
}
__host__ __device__ __forceinline__ bool xchg_is_step(unsigned long long w, uint32_t step) {
    return (uint32_t)(w >> 40) == (step & 0xFFFFFFu);
}

struct XchgBlock {
    unsigned long long count_slot[2][XCHG_MAX_WORLD];  // [step parity][rank]: written by that rank's push_kernel
    unsigned long long done_slot[2][XCHG_MAX_WORLD];   // root only: rows of [rank] have landed (value = its open-ended rows)
    unsigned long long my_offset;                      // where this rank's rows start in the root buffer (last push)
    unsigned long long base_rows;                      // rows gathered by the earlier phases of the job in progress
    unsigned int blocks_done;                          // last-block counter of push_kernel
    unsigned int push_ok;                              // publish_kernel: this rank may push (nothing void, everything fits)
    unsigned long long snap_total, snap_open;          // publish_kernel: this step's row / open-row counts, so that the scan
                                                       // counters are free for the next step's scan while the rows travel
    unsigned long long result[XCHG_RESULT_WORDS];      // settle_kernel: [0] status [1] total rows [2] total open
                                                       // [3] my offset [4] some rank has open-ended rows
                                                       // [5] rows of the job's earlier phases (before this step)
                                                       // [8 + r] rows of rank r
};

enum XchgStatus : uint32_t {
    XCHG_OK = 0,
    XCHG_VOID_STEP = 1,   // some rank's scan outgrew its buffers (or has a long spill list): repeat the step the slow way
    XCHG_ROOT_FULL = 2,   // the rows of all ranks do not fit the root buffer
    XCHG_TIMEOUT = 3,     // a peer never showed up
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// spin until the slot carries `step`; false on timeout
__device__ inline bool wait_slot(const unsigned long long *slot, uint32_t step, unsigned long long timeout_ns,
                                 unsigned long long *out) {
    unsigned long long v = ld_acquire_sys(slot);
    if (!xchg_is_step(v, step)) {
        const unsigned long long t0 = global_timer_ns();
        for (;;) {
            __nanosleep(200);
            v = ld_acquire_sys(slot);
            if (xchg_is_step(v, step)) break;
            if (global_timer_ns() - t0 > timeout_ns) { *out = v; return false; }
        }
    }
    *out = v;
    return true;
}

struct PushParams {
    XchgBlock *self;
    XchgBlock *peer[XCHG_MAX_WORLD];     // rank r's block in this rank's address space (peer[rank] == self)
    uint32_t *root_rows;                 // root buffer: 4 arrays of row_cap (record, start, end, k)
    uint64_t row_cap;
    uint32_t rank, world, step;
    const uint32_t *o_rec, *o_start, *o_end, *o_k;
    const unsigned long long *counters;
    uint32_t res_cap, open_cap;
    uint32_t trusted;                    // the host has already checked (and fixed up) the scan these rows come from
    uint32_t append;                     // this step's rows go after those of the job's earlier phases (base_rows)
    uint32_t compact;                    // 12 bytes per row over NVLink: record and motif size share one word (both < 65 536)
    unsigned long long timeout_ns;
};

// One warp per rank and step: publish my row count, wait for the counts of the ranks before me (lane q waits for rank
// q; they land in my OWN memory) and leave my prefix offset for push_kernel.  Only this warp ever spins -- the copy
// kernel behind it starts when its offset is known.
__global__ void __launch_bounds__(32) publish_kernel(const PushParams p) {
    const uint32_t lane = threadIdx.x, par = p.step & 1u;
    const unsigned long long n_stage = p.counters[C_STAGE], n_spill = p.counters[C_SPILL];
    const unsigned long long n_total = p.counters[C_TOTAL], n_open = p.counters[C_OPEN];
    // the conditions under which crf_scan would have grown a buffer, sorted a long spill list or re-run
    const bool valid = p.trusted || (n_stage <= p.res_cap && n_spill <= p.res_cap && n_spill <= SPILL_SMALL &&
                                     n_total <= p.res_cap && n_open <= p.open_cap);
    if (lane < p.world) {
        const unsigned long long w = xchg_enc(p.step, !valid, valid ? (uint32_t)n_total : 0u, n_open != 0);
        st_release_sys(&p.peer[lane]->count_slot[par][p.rank], w);
    }
    unsigned long long mine = 0;
    bool bad = false;
    if (lane < p.rank) {
        unsigned long long w;
        if (!wait_slot(&p.self->count_slot[par][lane], p.step, p.timeout_ns, &w) || (w & XCHG_VOID)) bad = true;
        mine = (uint32_t)w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
    const bool any_bad = __any_sync(0xFFFFFFFFu, bad);
    if (lane == 0) {
        const unsigned long long base = p.append ? p.self->base_rows : 0ull;   // (left by this rank's previous settle_kernel)
        const bool ok = valid && !any_bad && base + mine + n_total <= p.row_cap;
        p.self->my_offset = base + mine;
        p.self->push_ok = ok ? 1u : 0u;
        p.self->snap_total = n_total;
        p.self->snap_open = n_open;
    }
}

// My rows -> the root's buffer at my prefix offset (coalesced 4-byte peer stores over NVLink); the last block posts the
// done word on the root.
__global__ void __launch_bounds__(256) push_kernel(const PushParams p) {
    const uint32_t par = p.step & 1u;
    const unsigned long long off = p.self->my_offset;
    const bool ok = p.self->push_ok != 0;
    const unsigned long long n_total = p.self->snap_total, n_open = p.self->snap_open;   // (not the scan counters: see XchgBlock)
    if (ok) {
        // four columns, each copied with 16-byte peer stores: the destination starts at an arbitrary row, so every column
        // has a scalar head up to the first 16-byte boundary of the DESTINATION, a vector body (four scalar loads from my
        // own HBM feed one st.v4 over NVLink) and a scalar tail
        // compact rows: the record column is not sent; the motif-size column carries (record << 16 | k) and the root
        // splits it again (unpack_kernel) -- rank 0's NVLink ingress is what bounds the gather, 12 bytes beat 16
        const uint32_t n = (uint32_t)n_total;
        const uint32_t *src[4] = {p.o_rec, p.o_start, p.o_end, p.o_k};
        const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c == 0 && p.compact) continue;
            const bool fold = c == 3 && p.compact;
            uint32_t *dst = p.root_rows + (size_t)c * p.row_cap + off;
            const uint32_t *sp = src[c];
            const uint32_t *rp = p.o_rec;
            const uint32_t head = min(n, (uint32_t)((4u - (uint32_t)(((uintptr_t)dst >> 2) & 3u)) & 3u));
            const uint32_t nvec = (n - head) >> 2;
            if (gtid < head) dst[gtid] = fold ? (sp[gtid] | (rp[gtid] << 16)) : sp[gtid];
            uint4 *dv = reinterpret_cast<uint4 *>(dst + head);
            for (uint32_t v = gtid; v < nvec; v += gsize) {
                const uint32_t i = head + 4 * v;
                uint4 q = make_uint4(sp[i], sp[i + 1], sp[i + 2], sp[i + 3]);
                if (fold) { q.x |= rp[i] << 16; q.y |= rp[i + 1] << 16; q.z |= rp[i + 2] << 16; q.w |= rp[i + 3] << 16; }
                dv[v] = q;
            }
            const uint32_t tail0 = head + 4 * nvec;
            if (gtid < n - tail0) dst[tail0 + gtid] = fold ? (sp[tail0 + gtid] | (rp[tail0 + gtid] << 16)) : sp[tail0 + gtid];
        }
    }
    __threadfence_system();                               // my rows are visible on the root before the done word can be
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&p.self->blocks_done, 1u);
        if (prev == gridDim.x - 1) {                      // last block of this launch
            p.self->blocks_done = 0;
            __threadfence_system();
            st_release_sys(&p.peer[0]->done_slot[par][p.rank], xchg_enc(p.step, !ok, (uint32_t)min(n_open, 0xFFFFFFFFull)));
        }
    }
}

struct SettleParams {
    XchgBlock *self;
    uint64_t row_cap;
    uint32_t rank, world, step, append;
    unsigned long long timeout_ns;
};

// one warp: lane q waits for rank q
__global__ void __launch_bounds__(32) settle_kernel(const SettleParams p) {
    const uint32_t lane = threadIdx.x, par = p.step & 1u;
    unsigned long long cnt = 0, done = 0;
    bool arrived = true, is_void = false, has_open = false;
    if (lane < p.world) {
        arrived = wait_slot(&p.self->count_slot[par][lane], p.step, p.timeout_ns, &cnt);
        is_void = (cnt & XCHG_VOID) != 0;
        has_open = (cnt & XCHG_HAS_OPEN) != 0;
        if (p.rank == 0 && arrived) {
            arrived = wait_slot(&p.self->done_slot[par][lane], p.step, p.timeout_ns, &done);
            is_void = is_void || (done & XCHG_VOID) != 0;
        }
    }
    const uint32_t n = (lane < p.world && arrived) ? (uint32_t)cnt : 0u;
    const uint32_t nopen = (lane < p.world && arrived) ? (uint32_t)done : 0u;
    unsigned long long total = n, open_total = nopen;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
        open_total += __shfl_xor_sync(0xFFFFFFFFu, open_total, o);
    }
    const bool any_timeout = __any_sync(0xFFFFFFFFu, !arrived);
    const bool any_void = __any_sync(0xFFFFFFFFu, is_void);
    const bool any_open = __any_sync(0xFFFFFFFFu, has_open);
    if (lane < p.world) p.self->result[8 + lane] = n;
    if (lane == 0) {
        const unsigned long long base = p.append ? p.self->base_rows : 0ull;
        total += base;                                     // cumulative over the phases of this job
        uint32_t status = XCHG_OK;
        if (any_timeout) status = XCHG_TIMEOUT;
        else if (total > p.row_cap) status = XCHG_ROOT_FULL;
        else if (any_void) status = XCHG_VOID_STEP;
        p.self->base_rows = total;                         // every rank computes the same value from the same counts
        p.self->result[5] = base;
        p.self->result[0] = status;
        p.self->result[1] = total;
        p.self->result[2] = open_total;
        p.self->result[3] = p.self->my_offset;
        p.self->result[4] = any_open ? 1ull : 0ull;
    }
}

// root, compact rows: (record << 16 | k) of this step's rows -> the record and motif-size columns
__global__ void __launch_bounds__(256) unpack_kernel(XchgBlock *self, uint32_t *root_rows, uint64_t row_cap) {
    if (self->result[0] != XCHG_OK) return;
    const unsigned long long lo = self->result[5], hi = self->result[1];      // this step's rows: [base, total)
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
        const uint32_t v = root_rows[3 * row_cap + i];
        root_rows[i] = v >> 16;
        root_rows[3 * row_cap + i] = v & 0xFFFFu;
    }
}

// root: overwrite the end of gathered rows (stitched open-ended runs), idx = global row numbers
__global__ void __launch_bounds__(256) patch_rows_kernel(uint32_t *root_rows, uint64_t row_cap, const uint64_t *idx,
                                                         const uint32_t *new_end, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && idx[i] < row_cap) root_rows[2 * row_cap + idx[i]] = new_end[i];
}

}  // namespace crf
