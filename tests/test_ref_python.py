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
