#!/usr/bin/env python3
"""Generate golden vectors by running the REFERENCE itself (build container only).

Imports /root/reference/perfect_repeat_finder.py with `pyfastx` and `matplotlib` stubbed in
sys.modules (neither is installed; the hot path uses neither), runs the reference's own unit
tests as a sanity check, then records detect_repeats() outputs for

  * kat.json   -- the 32 known-answer calls of perfect_repeat_finder_tests.py:31-143, restated
                  as data (inputs + the reference's outputs),
  * fuzz_full.json, fuzz_interval.json, fuzz_minrep1.json, fuzz_interval_long.json, fuzz_interval_tail.json -- seeded random cases,
  * primitivity.json -- consists_of_perfect_repeats() on seeded strings.

The reference cannot travel to the GPU box, so these files are what pins the oracle
(oracle/crf_oracle.c) and, through it, the CUDA path.  Re-run:  python tests/golden/make_golden.py
"""
import argparse
import json
import os
import random
import sys
import types
import unittest

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ("pyfastx", "matplotlib", "matplotlib.pyplot", "matplotlib.colors"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.colors"].ListedColormap = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    sys.path.insert(0, REF)
    import perfect_repeat_finder as ref  # noqa
    return ref


def settings(**kw):
    return argparse.Namespace(**kw)


def run_case(ref, seq, fs_kwargs):
    fs = settings(**fs_kwargs)
    rec = {"seq": seq, "settings": fs_kwargs}
    try:
        out = ref.detect_repeats(seq, fs)
        rec["result"] = [[s, e, m] for s, e, m in out]
    except (ValueError, IndexError, AssertionError, AttributeError) as exc:
        rec["raises"] = type(exc).__name__
    return rec


def kat_cases():
    """The inputs of perfect_repeat_finder_tests.py:21-143, in order."""
    base = dict(min_motif_size=1, max_motif_size=7, min_repeats=3, min_span=6)
    cases = []
    for motif in "A", "CA", "CAG", "CAGA", "CAGAT", "CAGATT", "CAGATTA", "CAGATTAG":
        cases.append((6 * motif, dict(base)))
    for motif in "A", "CA", "CAG", "CAGA", "CAGAT", "CAGATT", "CAGATAT", "CAGATTAG":
        cases.append((6 * motif + 10 * "TA", dict(base)))
    cases.append(("A" * 9 + "C" * 11 + "G" * 10 + "T" * 9, dict(base)))
    cases.append(("CA" * 9 + "GT" * 11 + "CG" * 10 + "TA" * 9, dict(base)))
    cases.append((7 * "A", dict(base, min_motif_size=2)))
    seq = 7 * "A" + 7 * "AGAC" + 10 * "AAC" + "N" * 10
    cases.append((seq, dict(base, min_motif_size=2, max_motif_size=10)))
    cases.append((seq, dict(base, min_motif_size=1, max_motif_size=10)))
    for overlap_by in range(1, 10):
        left = "T" + "A" * overlap_by
        right = (overlap_by + 1) * "A" + "T"
        cases.append((9 * left + "T" + 11 * right, dict(base, max_motif_size=20, min_span=12)))
    seq = "N" * 10 + 7 * "A" + "N" + 7 * "AGAC" + "NNNN" + 10 * "AAC" + "N" * 10
    cases.append((seq, dict(base, max_motif_size=10, min_span=7)))
    seq = "GATGG" + "GGG" + "TGACATGACA" + "CAG" * 5 + "ACAGTTTTTTTTTT"
    cases.append((seq, dict(min_motif_size=1, max_motif_size=100, interval_start_0based=5, interval_end=20,
                            min_repeats=3, min_span=3)))
    assert len(cases) == 32
    return cases


ALPHABETS = ["ACGT", "ACGT", "ACGTN", "AC", "ACGTacgtNn", "ACGTRYN", "ACGTRYKMSWacgtryn", "AT"]


def random_seq(rng, max_len):
    n = rng.randint(0, max_len)
    alpha = rng.choice(ALPHABETS)
    parts = []
    total = 0
    while total < n:
        kind = rng.random()
        if kind < 0.45:
            piece = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 30)))
        elif kind < 0.85:
            unit = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 12)))
            piece = unit * rng.randint(1, 12) + unit[:rng.randint(0, len(unit))]
        elif kind < 0.95:
            piece = rng.choice("Nn") * rng.randint(1, 15)
        else:
            piece = rng.choice("RYKMSWBDHV") * rng.randint(1, 12)
        parts.append(piece)
        total += len(piece)
    return "".join(parts)[:n]


def fuzz(ref, seed, count, mode):
    rng = random.Random(seed)
    out = []
    for _ in range(count):
        seq = random_seq(rng, 260)
        kmin = rng.randint(1, 6)
        kmax = kmin + rng.choice([0, 1, 3, 8, 20, 45, 70])
        kw = dict(min_motif_size=kmin, max_motif_size=kmax,
                  min_repeats=rng.choice([2, 2, 3, 3, 3, 4, 5]), min_span=rng.choice([1, 2, 3, 6, 9, 9, 12, 20]))
        if mode == "minrep1":
            kw["min_repeats"] = 1
        if mode == "interval" or (mode == "minrep1" and rng.random() < 0.3):
            a = rng.randint(0, max(0, len(seq)))
            b = rng.randint(a, max(a, len(seq)))
            kw["interval_start_0based"] = a
            kw["interval_end"] = b
        out.append(run_case(ref, seq, kw))
    return out


def fuzz_interval_long(ref, seed, count):
    """Interval mode with repeats that run far past interval_end (the lock-step loop keeps going while a
    tracker is mid-repeat, prf:70-74) and with motif size 1 not always tracked."""
    rng = random.Random(seed)
    out = []
    for _ in range(count):
        seq = random_seq(rng, 1500)
        unit = "".join(rng.choice("ACGT") for _ in range(rng.choice([1, 1, 2, 3, 7])))
        pos = rng.randint(0, len(seq))
        seq = seq[:pos] + unit * rng.randint(150, 900) + seq[pos:]
        a = rng.randint(max(0, pos - 200), pos + 100)
        b = rng.randint(a, min(len(seq), pos + 300))
        kmin = rng.choice([1, 2, 3])
        kw = dict(min_motif_size=kmin, max_motif_size=kmin + rng.choice([3, 10, 30]), min_repeats=rng.choice([2, 3]),
                  min_span=rng.choice([9, 40]), interval_start_0based=a, interval_end=b)
        out.append(run_case(ref, seq, kw))
    return out


def fuzz_interval_tail(ref, seed, count):
    """Interval mode where interval_end lies within 2 * max_motif_size of the END of the sequence and the tail is periodic:
    trackers of large motif sizes have already stopped (trk:50) while still "in the middle of a repeat", so the lock-step
    loop runs on to the end (prf:70-74).  min_repeats 1 / 2 / 3."""
    rng = random.Random(seed)
    out = []
    for _ in range(count):
        seq = random_seq(rng, 400)
        kmin = rng.choice([1, 2, 3, 7])
        kmax = kmin + rng.choice([1, 5, 15, 30, 49])
        unit = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, kmax)))
        seq = seq + unit * rng.randint(2, 6) + unit[:rng.randint(0, len(unit))] + (random_seq(rng, 12) if rng.random() < 0.4 else "")
        b = max(0, len(seq) - rng.randint(0, 2 * kmax + 2))
        a = rng.randint(0, b)
        min_repeats = rng.choice([1, 2, 2, 3])
        min_span = rng.choice([2, 9, 9, 20])
        kw = dict(min_motif_size=kmin, max_motif_size=kmax, min_repeats=min_repeats, min_span=min_span,
                  interval_start_0based=a, interval_end=b)
        out.append(run_case(ref, seq, kw))
    return out


def primitivity_cases(seed, count):
    """Inputs / outputs of the reference's consists_of_perfect_repeats (utils/perfect_repeat_tracker.py:108-142)."""
    sys.path.insert(0, REF)
    from utils.perfect_repeat_tracker import consists_of_perfect_repeats
    rng = random.Random(seed)
    out = []
    for _ in range(count):
        alpha = rng.choice(["A", "AC", "ACG", "ACGT", "ACGTN"])
        if rng.random() < 0.6:
            unit = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 12)))
            seq = unit * rng.randint(1, 8)
        else:
            seq = "".join(rng.choice(alpha) for _ in range(rng.randint(0, 40)))
        out.append({"seq": seq, "unit": consists_of_perfect_repeats(seq)})
    return out


def main():
    ref = import_reference()
    # sanity: the reference's own test-suite passes under the stubs
    sys.path.insert(0, REF)
    suite = unittest.defaultTestLoader.loadTestsFromName("perfect_repeat_finder_tests")
    result = unittest.TextTestRunner(verbosity=0).run(suite)
    assert result.wasSuccessful() and result.testsRun == 3

    def dump(name, obj):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(obj, f, separators=(",", ":"))
        print(name, len(obj), "cases")

    dump("kat.json", [run_case(ref, s, kw) for s, kw in kat_cases()])
    dump("fuzz_full.json", fuzz(ref, 1001, 700, "full"))
    dump("fuzz_interval.json", fuzz(ref, 2002, 400, "interval"))
    dump("fuzz_minrep1.json", fuzz(ref, 3003, 300, "minrep1"))
    dump("fuzz_interval_long.json", fuzz_interval_long(ref, 4004, 40))
    dump("fuzz_interval_tail.json", fuzz_interval_tail(ref, 6006, 240))
    dump("primitivity.json", primitivity_cases(5005, 600))


if __name__ == "__main__":
    main()
