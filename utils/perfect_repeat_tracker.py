"""Compatibility module for the reference's utils/perfect_repeat_tracker.py.

`consists_of_perfect_repeats` (trk:108-142) is a pure string helper and is provided.  The per-base state machine
`PerfectRepeatTracker` (trk:3-105) is what this build replaces with CUDA kernels: there is deliberately no CPU tracker
here, and constructing one says so.
"""


def consists_of_perfect_repeats(sequence):
    """The repeat unit if `sequence` is two or more copies of a shorter unit (the shortest such unit), else None --
    same contract as trk:108-142 (e.g. "CAGCAGCAG" -> "CAG", "CAGCA" -> None)."""
    n = len(sequence)
    for unit_length in range(1, n // 2 + 1):
        if n % unit_length == 0 and sequence == sequence[:unit_length] * (n // unit_length):
            return sequence[:unit_length]
    return None


class PerfectRepeatTracker:
    """Not available: the tracker loop runs on the GPU (perfect_repeat_finder.detect_repeats)."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "PerfectRepeatTracker is replaced by the CUDA scan in this build; call perfect_repeat_finder.detect_repeats() "
            "(there is no CPU implementation of the tracker loop)")
