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


# ---- file-level parity of the command line: the reference's main() (prf:83-179) against crf_b200.cli.main -----------------
def _cli_fasta(tmp_path, gz):
    import gzip
    rng = random.Random(31)
    records = []
    for name in ("chr1", "chrUn_gl000220 some description", "scaffold-3"):
        seq = random_seq(rng, rng.randint(400, 3000), exotic=True)
        unit = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 30)))
        pos = rng.randint(0, len(seq))
        seq = seq[:pos] + unit * rng.randint(3, 40) + seq[pos:]
        seq = "".join(c.lower() if rng.random() < 0.2 else c for c in seq)
        records.append((name, seq))
    text = "".join(f">{n}\n" + "".join(s[i:i + 70] + "\n" for i in range(0, len(s), 70)) for n, s in records)
    path = tmp_path / ("in.fa.gz" if gz else "in.fasta")
    (gzip.open(path, "wt") if gz else open(path, "wt")).write(text)
    return path, [(n.split()[0], s) for n, s in records]


def _cli_argument_lists(path, records):
    rng = random.Random(32)
    out = []
    for i in range(36):
        name, seq = rng.choice(records)
        a = rng.choice([0, rng.randint(0, len(seq)), max(0, len(seq) - rng.randint(0, 120))])
        b = rng.choice([rng.randint(a, len(seq)), len(seq), len(seq) + 500, a])
        argv = [str(path), "--interval", f"{name}:{a}-{b}"]
        if rng.random() < 0.7:
            kmin = rng.choice([1, 2, 3])
            argv += ["-min", str(kmin), "-max", str(kmin + rng.choice([0, 5, 30, 49]))]
        if rng.random() < 0.5:
            argv += ["--min-repeats", str(rng.choice([1, 2, 3, 4])), "--min-span", str(rng.choice([2, 9, 12, 30]))]
        if rng.random() < 0.5:
            argv += ["-o", f"sub/dir/prefix{i}"]                     # the BED file still lands in the working directory
        out.append(argv)
    for raw in ("ACGTACGTACGTTTTTTTTTT", "cagcagcagcagNNNNNNcagcag" * 3, "A" * 40, "ACGT", "nnnnnnnn", "TTTTTTTTTNACACACAC"):
        out.append([raw])
        out.append([raw, "-o", "named", "-min", "2", "-max", "6", "--min-repeats", "2", "--min-span", "4"])
    out += [[str(path), "--interval", "chr1:5"], [str(path), "--interval", "nope:1-50"], ["ACGT", "--interval", "chr1:1-5"],
            ["ACGTXZ"], ["ACGT", "-min", "0"], ["ACGT", "-min", "5", "-max", "4"], ["ACGT", "--min-repeats", "0"],
            ["ACGT", "--min-span", "0"]]
    return out


def _run_dropin_cli(cli_main, argv, cwd, capsys):
    saved = os.getcwd()
    os.chdir(cwd)
    try:
        capsys.readouterr()
        try:
            status = cli_main([str(a) for a in argv])
        except SystemExit as e:
            status = e.code
        return status, capsys.readouterr().out
    finally:
        os.chdir(saved)


def _compare_cli(cli_main, tmp_path, capsys, refusal=()):
    n_files = n_errors = 0
    for gz in (False, True):
        path, records = _cli_fasta(tmp_path, gz)
        for i, argv in enumerate(_cli_argument_lists(path, records)):
            d_ref, d_mine = tmp_path / f"ref{int(gz)}_{i}", tmp_path / f"mine{int(gz)}_{i}"
            d_ref.mkdir(), d_mine.mkdir()
            try:
                want = ref.run_main(argv, d_ref)
            except (AssertionError, IndexError) as e:
                want = type(e).__name__
            try:
                got = _run_dropin_cli(cli_main, argv, d_mine, capsys)
            except (AssertionError, IndexError) as e:
                got = type(e).__name__
            except refusal:
                continue
            assert got == want, argv
            files_ref = sorted(p.name for p in d_ref.iterdir())
            assert sorted(p.name for p in d_mine.iterdir()) == files_ref, argv
            for name in files_ref:
                assert (d_mine / name).read_bytes() == (d_ref / name).read_bytes(), (argv, name)
            n_files += len(files_ref)
            n_errors += want == (2, "") or isinstance(want, str)
    assert n_files >= 80 and n_errors >= 16


@needs_ref
def test_cli_files_and_stdout_equal_the_reference_cli_host_side(tmp_path, capsys, monkeypatch):
    """The reference's own main() (with a line-by-line stand-in for pyfastx) and the drop-in CLI on the same argument lists:
    same exit status, same stdout, same BED / TSV bytes -- --interval on plain and gzipped FASTA (mixed case, N blocks, IUPAC
    letters, intervals clamped to the record, min_repeats 1..4), raw strings, every parser.error path.  Here the GPU scan is
    the closed-form stand-in (host glue only); the -m gpu twin below runs the real thing."""
    from crf_b200 import api, cli
    from tests import test_host_cpu as stand_in
    monkeypatch.setattr(api, "get_context", lambda device=None: stand_in._ClosedFormCtx())
    closed_form = api.scan_arrays

    def scan_arrays(s, kmin, kmax, min_repeats, span, device=None, **kw):
        if min_repeats == 1:
            return stand_in._runs_single_copy(bytes(s), kmin, kmax, span)
        return closed_form(s, kmin, kmax, min_repeats, span, device=device, **kw)

    monkeypatch.setattr(api, "scan_arrays", scan_arrays)
    _compare_cli(cli.main, tmp_path, capsys, refusal=(NotImplementedError,))


@needs_ref
@pytest.mark.gpu
def test_cli_files_and_stdout_equal_the_reference_cli(tmp_path, capsys):
    from crf_b200 import cli
    _compare_cli(cli.main, tmp_path, capsys, refusal=(NotImplementedError,))
