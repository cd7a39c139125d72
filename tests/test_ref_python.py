"""Tier L0: the UNMODIFIED Python reference, copied to oracle/_ref/ by oracle/make_ref.sh
(git-ignored, shipped to the GPU box by gpurun), against the C oracle (CPU) and against the CUDA path (-m gpu).
The committed golden JSON pins the oracle where oracle/_ref is absent; where it is present these tests pin it
again, live, on inputs that are in no fixture."""
import importlib.util
import os
import random
import sys
import unittest

import pytest

from oracle import oracle, ref
from tests.helpers import ns, random_seq

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref missing (run oracle/make_ref.sh where /root/reference exists)")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cases(seed, n_cases, max_len):
    rng = random.Random(seed)
    for _ in range(n_cases):
        n = rng.choice([rng.randint(0, 30), rng.randint(30, 300), rng.randint(300, max_len)])
        seq = random_seq(rng, n, exotic=rng.random() < 0.2)
        kmin = rng.randint(1, 6)
        fs = dict(min_motif_size=kmin, max_motif_size=rng.randint(kmin, rng.choice([6, 12, 50])),
                  min_repeats=rng.choice([1, 2, 2, 3, 3, 3, 4]), min_span=rng.choice([1, 5, 9, 9, 12, 30]))
        if rng.random() < 0.35 and n > 2:
            a = rng.randint(0, n - 1)
            fs["interval_start_0based"] = a
            fs["interval_end"] = rng.randint(a, n)
        yield seq, fs


def _outcome(fn, seq, fs):
    try:
        return fn(seq, ns(**fs))
    except (ValueError, IndexError, AssertionError, NotImplementedError) as exc:
        return type(exc).__name__


@needs_ref
def test_reference_own_tests_pass_from_the_copy():
    _, ref_tests = ref.load()
    res = unittest.TextTestRunner(verbosity=0).run(unittest.defaultTestLoader.loadTestsFromModule(ref_tests))
    assert res.wasSuccessful() and res.testsRun == 3


@needs_ref
def test_oracle_equals_python_reference_live_fuzz():
    n_rows = n_exc = 0
    for seq, fs in _cases(2026, 400, 2500):
        want = _outcome(ref.detect_repeats, seq, fs)
        got = _outcome(oracle.detect_repeats, seq, fs)
        assert got == want, f"{fs} {seq!r}"
        if isinstance(want, str):
            n_exc += 1
        else:
            n_rows += len(want)
    assert n_rows > 1000 and n_exc >= 2


def _load_reference_tests_against_dropin():
    """The reference's test module, source unmodified, with `perfect_repeat_finder` / `utils.plot_utils` resolving to this
    repo's drop-in modules (perfect_repeat_finder_tests.py:5-6 imports exactly those two names)."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import perfect_repeat_finder as dropin
    import utils.plot_utils as dropin_utils
    assert os.path.dirname(os.path.abspath(dropin.__file__)) == ROOT
    assert os.path.dirname(os.path.dirname(os.path.abspath(dropin_utils.__file__))) == ROOT
    path = os.path.join(ref.REF_DIR, "perfect_repeat_finder_tests.py")
    spec = importlib.util.spec_from_file_location("reference_tests_vs_dropin", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.detect_repeats is dropin.detect_repeats
    return mod


@needs_ref
@pytest.mark.gpu
def test_reference_test_file_unmodified_against_the_dropin():
    mod = _load_reference_tests_against_dropin()
    res = unittest.TextTestRunner(verbosity=0).run(unittest.defaultTestLoader.loadTestsFromModule(mod))
    assert res.testsRun == 3
    assert res.wasSuccessful(), [str(f[1]) for f in res.failures + res.errors]


@needs_ref
@pytest.mark.gpu
def test_gpu_equals_python_reference_live_fuzz():
    import perfect_repeat_finder as dropin
    n_rows = n_exc = 0
    for seq, fs in _cases(77, 300, 6000):
        if fs["min_repeats"] == 1 and fs["min_motif_size"] == 1 and fs["min_span"] <= 1:
            continue                                        # documented refusal
        want = _outcome(ref.detect_repeats, seq, fs)
        got = _outcome(dropin.detect_repeats, seq, fs)
        assert got == want, f"{fs} {seq!r}"
        if isinstance(want, str):
            n_exc += 1
        else:
            n_rows += len(want)
    assert n_rows > 1000


@needs_ref
@pytest.mark.gpu
def test_tracker_compat_class_replays_the_reference_tracker():
    """utils.perfect_repeat_tracker.PerfectRepeatTracker (GPU-backed facade) driven in lock-step exactly as the reference's
    detect_repeats drives its trackers (prf:51-79), against the reference's own class doing the same: same dict, same
    is_in_middle_of_repeat() at every position, same current_position; done() called mid-sequence too."""
    import importlib
    ref_prf, _ = ref.load()
    RefTracker = ref_prf.PerfectRepeatTracker
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    mine = importlib.import_module("utils.perfect_repeat_tracker")
    assert os.path.dirname(os.path.dirname(os.path.abspath(mine.__file__))) == ROOT
    rng = random.Random(5)
    checked = 0
    for it in range(40):
        seq = random_seq(rng, rng.choice([50, 400, 3000]), exotic=rng.random() < 0.3)
        if rng.random() < 0.5:
            seq = seq.upper()
        mr, ms = rng.choice([2, 3, 3, 4]), rng.choice([1, 6, 9, 20])
        ks = sorted(rng.sample(range(1, 30), 4))
        out_ref, out_mine = {}, {}
        a = [RefTracker(k, mr, ms, seq, out_ref) for k in ks]
        b = [mine.PerfectRepeatTracker(k, mr, ms, seq, out_mine) for k in ks]
        stop = rng.randint(0, len(seq)) if rng.random() < 0.4 else len(seq)
        for _pos in range(stop):
            for x, y in zip(a, b):
                assert x.advance() == y.advance()
                assert x.is_in_middle_of_repeat() == y.is_in_middle_of_repeat()
                assert x.current_position == y.current_position
        for x, y in zip(a, b):
            x.done()
            y.done()
        assert out_mine == out_ref, (seq, ks, mr, ms, stop)
        checked += len(out_ref)
    assert checked > 100
