"""A CPU stand-in for crf_b200._cabi.Context used by the partition tests: same load_ranges / scan /
fetch / run_end surface, answers computed with the CPU oracle.  Test infrastructure only."""
import argparse
import ctypes

import numpy as np

from oracle import oracle


class FakeStats:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def as_dict(self):
        return dict(self.__dict__)


class FakeSeq:
    def __init__(self, bases, starts, lens, own_lo, own_hi):
        buf = np.frombuffer(bytes(bases), dtype=np.uint8) if not isinstance(bases, np.ndarray) else bases
        self.units = [buf[int(s):int(s) + int(n)] for s, n in zip(starts, lens)]
        self.own_lo = [0] * len(self.units) if own_lo is None else [int(x) for x in own_lo]
        self.own_hi = [len(u) for u in self.units] if own_hi is None else [int(x) for x in own_hi]
        self._rows = None
        self._map = None
        self._open = np.zeros((0, 5), np.uint32)

    # ---- the library's output map / open-ended rows / exchange calls (crf_seq_set_output_map, crf_fetch_open, ...) ----
    def set_output_map(self, out_record=None, out_shift=None, open_ended=None):
        n = len(self.units)
        self._map = (list(out_record) if out_record is not None else list(range(n)),
                     [int(x) for x in out_shift] if out_shift is not None else [0] * n,
                     [int(x) for x in open_ended] if open_ended is not None else [0] * n)

    def _apply_map(self):
        if self._map is None:
            return
        rec, shift, opened = self._map
        rows, open_rows = self._rows.copy(), []
        for i, (u, a, b, k) in enumerate(self._rows.tolist()):
            rows[i] = (rec[u], a + shift[u], b + shift[u], k)
            if opened[u] and b == len(self.units[u]):
                open_rows.append((i, rec[u], a + shift[u], b + shift[u], k))
        self._rows = rows
        self._open = np.array(open_rows, dtype=np.uint32).reshape(-1, 5)

    def stats(self):
        n = 0 if self._rows is None else len(self._rows)
        return FakeStats(scan_ms=1.0, kernel_ms=1.0, n_results=n, n_tiles=0, n_spilled=0, n_long=0, n_candidates=0,
                         word_k_pairs=0, reruns=0, launches=0, n_open=len(self._open))

    def fetch_open(self, cap=256):
        return self._open[:cap]

    def patch_end(self, row, new_end):
        self._rows[int(row), 2] = int(new_end)

    def scan_gather(self, xchg, kmin, kmax, min_repeats, min_span, append=False, **knobs):
        self.scan(kmin, kmax, min_repeats, min_span, **knobs)
        xchg.enqueue(self._rows.copy(), len(self._open), append)

    def push(self, xchg, append=False):
        xchg.enqueue(self._rows.copy(), len(self._open), append)

    def scan(self, kmin, kmax, min_repeats, min_span, **knobs):
        fs = argparse.Namespace(min_motif_size=kmin, max_motif_size=kmax, min_repeats=min_repeats, min_span=min_span)
        rows = []
        for i, u in enumerate(self.units):
            s, e, m, _ = oracle.detect_repeats_by_k(np.ascontiguousarray(u), fs, arrays=True)
            keep = (s >= self.own_lo[i]) & (s < self.own_hi[i])
            for a, b, c in zip(s[keep], e[keep], m[keep]):
                rows.append((i, int(a), int(b), int(c)))
        self._rows = np.array(rows, dtype=np.int64).reshape(-1, 4)
        self._open = np.zeros((0, 5), np.uint32)
        self._apply_map()
        return len(rows)

    def fetch(self, n):
        r = self._rows
        return tuple(r[:, j].astype(np.uint32) for j in range(4))

    def run_end(self, record, pos, k):
        u = bytes(self.units[record]).upper()
        j = pos
        while j + k < len(u) and u[j] == u[j + k] and u[j:j + 1] != b"N":
            j += 1
        return j

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def unpack_planes(planes):
    """PackedPlanes -> the upper-cased text they stand for (uint8 array)."""
    n = planes.n_bases
    nw = (n + 31) // 32

    def bits(a):
        a = np.ctypeslib.as_array((ctypes.c_uint32 * nw).from_address(a.value)) if isinstance(a, ctypes.c_void_p) else \
            (np.asarray(a)[:nw] if not isinstance(a, int) else np.ctypeslib.as_array((ctypes.c_uint32 * nw).from_address(a)))
        return ((a[:, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(np.uint8).reshape(-1)[:n]
    h, l, nm = bits(planes.H_ptr), bits(planes.L_ptr), bits(planes.NM_ptr)
    if planes.runs is not None:                       # the runs must say what the plane says
        from_runs = np.zeros(n, np.uint8)
        for a, b in planes.runs.tolist():
            from_runs[a:b] = 1
        assert np.array_equal(from_runs, nm)
    out = np.frombuffer(b"ACGT", dtype=np.uint8)[(h << 1) | l].copy()
    out[nm == 1] = ord("N")
    for key in planes.exotic.tolist():
        out[key >> 8] = key & 0xFF
    return out


class FakeContext:
    def load_packed(self, planes, offsets=None, max_motif_cap=50, ranges=None):
        text = unpack_planes(planes)
        if ranges is not None:
            return FakeSeq(text, *ranges)
        return self.load(text, offsets, max_motif_cap)

    def load(self, bases, offsets=None, max_motif_cap=50, on_device=False):
        n = len(bases)
        offsets = np.array([0, n], dtype=np.uint64) if offsets is None else np.asarray(offsets, dtype=np.uint64)
        return FakeSeq(bases, offsets[:-1], np.diff(offsets.astype(np.int64)), None, None)

    def load_ranges(self, bases, starts, lengths, own_lo=None, own_hi=None, max_motif_cap=50, on_device=False):
        return FakeSeq(bases, starts, lengths, own_lo, own_hi)


class FakeXchgResult:
    pass


class FakeXchg:
    """CPU stand-in for _cabi.Xchg (csrc/crf_xchg.cuh): same calls, the rows travel through torch.distributed (ranks are
    processes) or through the peer objects (ranks are threads of one process) instead of NVLink."""
    dist = None                                     # set by the test when the ranks are processes of a gloo group

    def __init__(self, ctx, rank, world, row_cap):
        self.rank, self.world, self.row_cap = rank, world, row_cap
        self.peers = {rank: self}
        self.queue, self.inbox = [], {}
        self.step = 0
        self.results = {}
        self.rows = np.zeros((0, 4), np.int64)

    def export(self):
        return b"%d" % self.rank

    def connect_ipc(self, peer, handle):
        assert int(handle) == peer

    def connect_local(self, peer, other):
        self.peers[peer] = other

    def set_timeout(self, seconds):
        pass

    def set_compact(self, on=True):
        pass

    def enqueue(self, rows, n_open, append):
        self.step += 1
        self.queue.append((self.step, rows, n_open, append))

    def _exchange(self, step, payload):
        if FakeXchg.dist is not None:
            out = [None] * self.world
            FakeXchg.dist.all_gather_object(out, payload)
            return out
        import time
        for r in range(self.world):
            self.peers[r].inbox[(step, self.rank)] = payload
        while any((step, r) not in self.inbox for r in range(self.world)):
            time.sleep(0.001)
        return [self.inbox.pop((step, r)) for r in range(self.world)]

    def wait(self):
        last, checked = None, 0
        for step, rows, n_open, append in self.queue:
            parts = self._exchange(step, (rows, n_open))
            base = len(self.rows) if append else 0
            res = FakeXchgResult()
            res.status = res.worst_status = 0
            res.step = step
            res.any_open = int(any(p[1] for p in parts))
            res.base_rows = base
            res.rows_of_rank = [len(p[0]) for p in parts] + [0] * (16 - self.world)
            res.total_rows = base + sum(res.rows_of_rank)
            res.total_open = sum(p[1] for p in parts)
            res.my_offset = base + sum(res.rows_of_rank[:self.rank])
            if self.rank == 0:
                self.rows = np.concatenate([self.rows[:base]] + [p[0].reshape(-1, 4) for p in parts])
            self.results[step] = res
            last, checked = res, checked + 1
        self.queue = []
        last.steps_checked = checked
        return last

    def step_result(self, step):
        return self.results[step]

    def fetch(self, n, first=0):
        r = self.rows[first:first + n]
        return tuple(r[:, j].astype(np.uint32) for j in range(4))

    def patch_end(self, rows, new_end):
        for i, e in zip(rows, new_end):
            self.rows[int(i), 2] = int(e)

    def close(self):
        pass
