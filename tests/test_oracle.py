"""The CPU oracle (oracle/crf_oracle.c) against the golden vectors generated from the reference
(tests/golden/make_golden.py).  This is what pins the oracle; the -m gpu tests then compare the
CUDA path with the oracle."""
import random

import numpy as np
import pytest

from oracle import oracle
from tests.helpers import exc_of, expected_of, load_golden, ns, random_seq


def _run(case, fn=oracle.detect_repeats):
    return fn(case["seq"], ns(**case["settings"]))


@pytest.mark.parametrize("name", ["kat.json", "fuzz_full.json", "fuzz_interval.json", "fuzz_minrep1.json", "fuzz_interval_long.json",
                                  "fuzz_interval_tail.json"])
def test_oracle_matches_reference_golden(name):
    cases = load_golden(name)
    assert cases
    for i, case in enumerate(cases):
        if "raises" in case:
            with pytest.raises(exc_of(case)):
                _run(case)
        else:
            assert _run(case) == expected_of(case), f"{name}[{i}] {case['settings']} {case['seq']!r}"


def test_by_k_matches_lockstep_on_golden():
    for name in ["kat.json", "fuzz_full.json"]:
        for i, case in enumerate(load_golden(name)):
            if "raises" in case or "interval_end" in case["settings"]:
                continue
            got = oracle.detect_repeats_by_k(case["seq"], ns(**case["settings"]), threads=3)
            assert got == expected_of(case), f"{name}[{i}]"


def test_by_k_matches_lockstep_medium():
    rng = random.Random(7)
    for _ in range(6):
        seq = random_seq(rng, 20000, exotic=True)
        fs = ns(min_motif_size=1, max_motif_size=50, min_repeats=rng.choice([2, 3]), min_span=rng.choice([6, 9]))
        assert oracle.detect_repeats_by_k(seq, fs) == oracle.detect_repeats(seq, fs)


def test_filter_validation_messages():
    with pytest.raises(ValueError, match="min_motif_size is set to 0"):
        oracle.detect_repeats("ACGT", ns(min_motif_size=0, max_motif_size=3, min_repeats=3, min_span=9))
    with pytest.raises(ValueError, match="max_motif_size is set to 1"):
        oracle.detect_repeats("ACGT", ns(min_motif_size=2, max_motif_size=1, min_repeats=3, min_span=9))
    with pytest.raises(ValueError, match="min_repeats"):
        oracle.detect_repeats("ACGT", ns(min_motif_size=1, max_motif_size=3, min_repeats=0, min_span=9))
    with pytest.raises(ValueError, match="min_span"):
        oracle.detect_repeats("ACGT", ns(min_motif_size=1, max_motif_size=3, min_repeats=3, min_span=None))
    with pytest.raises(AttributeError):
        oracle.detect_repeats("ACGT", ns(min_motif_size=1, max_motif_size=3, min_repeats=3))


def test_arrays_interface_and_steps():
    seq = np.frombuffer(b"ACGT" * 10 + b"N" * 5 + b"ca" * 9, dtype=np.uint8)
    fs = ns(min_motif_size=1, max_motif_size=6, min_repeats=3, min_span=9)
    start, end, mlen, steps = oracle.detect_repeats_by_k(seq, fs, arrays=True)
    assert list(zip(start.tolist(), end.tolist(), mlen.tolist())) == [(0, 40, 4), (45, 63, 2)]
    assert steps > 0
