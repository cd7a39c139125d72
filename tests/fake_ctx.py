"""A CPU stand-in for crf_b200._cabi.Context used by the partition tests: same load_ranges / scan /
fetch / run_end surface, answers computed with the CPU oracle.  Test infrastructure only."""
import argparse
import ctypes

import numpy as np

from oracle import oracle


class FakeSeq:
    def __init__(self, bases, starts, lens, own_lo, own_hi):
        buf = np.frombuffer(bytes(bases), dtype=np.uint8) if not isinstance(bases, np.ndarray) else bases
        self.units = [buf[int(s):int(s) + int(n)] for s, n in zip(starts, lens)]
        self.own_lo = [0] * len(self.units) if own_lo is None else [int(x) for x in own_lo]
        self.own_hi = [len(u) for u in self.units] if own_hi is None else [int(x) for x in own_hi]
        self._rows = None

    def scan(self, kmin, kmax, min_repeats, min_span, **knobs):
        fs = argparse.Namespace(min_motif_size=kmin, max_motif_size=kmax, min_repeats=min_repeats, min_span=min_span)
        rows = []
        for i, u in enumerate(self.units):
            s, e, m, _ = oracle.detect_repeats_by_k(np.ascontiguousarray(u), fs, arrays=True)
            keep = (s >= self.own_lo[i]) & (s < self.own_hi[i])
            for a, b, c in zip(s[keep], e[keep], m[keep]):
                rows.append((i, int(a), int(b), int(c)))
        self._rows = np.array(rows, dtype=np.int64).reshape(-1, 4)
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
