"""Compatibility module for the reference's utils/perfect_repeat_tracker.py.

`consists_of_perfect_repeats` (trk:108-142) is a pure string helper.  `PerfectRepeatTracker` (trk:3-105) keeps the
reference's interface -- advance() / is_in_middle_of_repeat() / done() / current_position, results written into the
caller's `output_intervals` dict -- but there is no per-base state machine behind it: the first call scans the whole
sequence for this motif size on the GPU (crf_b200: two crf_scan calls) and the methods then replay what the
reference's tracker would have done at each position.  Code that drives trackers in lock-step, as the reference's
detect_repeats does (perfect_repeat_finder.py:51-79), gets the same dict; detect_repeats() itself is the fast path.
"""
import bisect
import os
import sys

_PKG_PARENT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "colab-repeat-finder_b200")
if _PKG_PARENT not in sys.path:
    sys.path.insert(0, _PKG_PARENT)


def consists_of_perfect_repeats(sequence):
    """The repeat unit if `sequence` is two or more copies of a shorter unit (the shortest such unit), else None --
    same contract as trk:108-142 (e.g. "CAGCAGCAG" -> "CAG", "CAGCA" -> None)."""
    n = len(sequence)
    for unit_length in range(1, n // 2 + 1):
        if n % unit_length == 0 and sequence == sequence[:unit_length] * (n // unit_length):
            return sequence[:unit_length]
    return None


class PerfectRepeatTracker:
    """Tracks repeats of ONE motif size in `input_sequence` (same constructor and methods as trk:3-105).

    The reference's tracker reads `input_sequence` as given (detect_repeats upper-cases it first, prf:33); so does this
    one: lower-case letters are ordinary symbols here too.  min_repeats == 1 (the reference's negative-index corner,
    trk:86-91) is served by detect_repeats() only."""

    def __init__(self, motif_size, min_repeats, min_span, input_sequence, output_intervals, verbose=False):
        if min_repeats < 2:
            raise NotImplementedError("PerfectRepeatTracker with min_repeats == 1: use perfect_repeat_finder.detect_repeats(), "
                                      "which reproduces the reference's single-copy corner cases")
        self.motif_size = motif_size
        self.min_repeats = min_repeats
        self.min_span = min_span
        self.input_sequence = input_sequence
        self.output_intervals = output_intervals
        self.verbose = verbose
        self._current_position = 0
        self._scanned = False

    # ---- the GPU scan behind the facade ---------------------------------------------------------------------
    def _scan(self):
        import numpy as np
        from crf_b200 import _cabi, api
        k, seq = self.motif_size, self.input_sequence
        self._emit_at, self._mid_lo, self._mid_hi, self._rows = {}, [], [], []
        if len(seq) > k:
            raw = np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)
            # the library upper-cases ASCII, the tracker must not: lower-case letters are moved to byte values the text does
            # not use, which the library treats as distinct "exotic" symbols (equal bytes match, only the literal N never does)
            lower = (raw >= 97) & (raw <= 122)
            if lower.any():
                used = np.bincount(raw, minlength=256) > 0
                free = [v for v in range(128, 256) if not used[v]]
                table = np.arange(256, dtype=np.uint8)
                for c in np.flatnonzero(used[97:123]) + 97:
                    table[c] = free.pop()
                raw = table[raw]
            ctx = api.get_context()
            with ctx.load(raw, max_motif_cap=k) as s:
                n = s.scan(k, k, self.min_repeats, self.min_span)
                _, st, en, _ = s.fetch(n)                       # what the tracker reports (trk:86-101)
                m = s.scan(k, k, 2, 1, flags=_cabi.SCAN_NO_PRIMITIVITY)
                _, st2, en2, _ = s.fetch(m)                     # every run of >= k matches (trk:63-65)
            for a, b in zip(st.tolist(), en.tolist()):
                self._emit_at[b - k] = (a, b)                   # emitted when the mismatch position i0 = end - k is processed
            self._rows = sorted(self._emit_at.values())
            self._mid_lo = [a + k for a in st2.tolist()]        # run_length >= k + 1 for positions in [st + k, i0]
            self._mid_hi = [b - k for b in en2.tolist()]
        self._scanned = True

    def _emit(self, i0):
        row = self._emit_at.get(i0)
        if row is None:
            return
        motif = self.input_sequence[row[0]:row[0] + self.motif_size]
        previous = self.output_intervals.get(row)
        if previous is not None and len(motif) > len(previous):      # trk:94-96
            return
        self.output_intervals[row] = motif

    # ---- the reference's interface ----------------------------------------------------------------------------
    def advance(self):
        """One base forward; False once the tracker has reached len(seq) - motif_size (trk:43-61)."""
        if not self._scanned:
            self._scan()
        i = self._current_position
        if i >= len(self.input_sequence) - self.motif_size:
            return False
        self._emit(i)                                            # i is a mismatch position that closes a reportable run
        self._current_position += 1
        return True

    def is_in_middle_of_repeat(self):
        """run_length >= motif_size + 1 (trk:63-65): at least motif_size matches end just before the current position."""
        if not self._scanned:
            self._scan()
        p = self._current_position
        j = bisect.bisect_right(self._mid_lo, p) - 1
        return j >= 0 and p <= self._mid_hi[j]

    def done(self):
        """trk:67-69: the run the tracker is in (usually the one that touches the end of the sequence) is reported if what
        has been seen of it already passes the thresholds -- with its real end, which trk:87 finds by reading ahead."""
        if not self._scanned:
            self._scan()
        p, k = self._current_position, self.motif_size
        j = bisect.bisect_right(self._rows, (p, float("inf"))) - 1
        if j >= 0:
            st, end = self._rows[j]
            if p <= end - k and p - st + k >= max(self.min_span, self.min_repeats * k):
                self._emit(end - k)

    @property
    def current_position(self):
        return self._current_position
