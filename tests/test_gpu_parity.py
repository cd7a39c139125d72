"""Parity of the CUDA path (through the C ABI / ctypes) with the reference's golden vectors and
with the CPU oracle.  Bit-exact: integer coordinates and motif strings must be identical."""
import random

import numpy as np
import pytest

from tests.helpers import exc_of, expected_of, load_golden, ns, random_seq

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def prf():
    import perfect_repeat_finder
    return perfect_repeat_finder


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


KNOB_SETS = [
    {},                                                                      # library defaults (T=8)
    {"words_per_thread": 1},                                                 # 32-base strips, 8 kbp tiles
    {"words_per_thread": 16},
    {"words_per_thread": 1, "tile_out_cap": 2, "walk_limit_words": 1},       # force spill + block walker
]


@pytest.mark.parametrize("knobs", KNOB_SETS)
@pytest.mark.parametrize("name", ["kat.json", "fuzz_full.json", "fuzz_interval.json", "fuzz_interval_long.json"])
def test_golden_vectors(prf, name, knobs):
    for i, case in enumerate(load_golden(name)):
        fs = ns(**case["settings"])
        if "raises" in case:
            with pytest.raises(exc_of(case)):
                prf.detect_repeats(case["seq"], fs, **knobs)
        else:
            got = prf.detect_repeats(case["seq"], fs, **knobs)
            assert got == expected_of(case), f"{name}[{i}] {case['settings']} {case['seq']!r}"


def test_min_repeats_1_is_refused_loudly(prf):
    with pytest.raises(NotImplementedError):
        prf.detect_repeats("ACGTACGTACGT", ns(min_motif_size=1, max_motif_size=4, min_repeats=1, min_span=3))


def test_filter_validation(prf):
    with pytest.raises(ValueError, match="min_motif_size is set to 0"):
        prf.detect_repeats("ACGT", ns(min_motif_size=0, max_motif_size=3, min_repeats=3, min_span=9))
    with pytest.raises(ValueError, match="max_motif_size is set to 1"):
        prf.detect_repeats("ACGT", ns(min_motif_size=2, max_motif_size=1, min_repeats=3, min_span=9))
    with pytest.raises(ValueError, match="min_repeats"):
        prf.detect_repeats("ACGT", ns(min_motif_size=1, max_motif_size=3, min_repeats=0, min_span=9))
    with pytest.raises(ValueError, match="min_span"):
        prf.detect_repeats("ACGT", ns(min_motif_size=1, max_motif_size=3, min_repeats=3, min_span=None))
    with pytest.raises(AttributeError):
        prf.detect_repeats("ACGT", ns(min_motif_size=1, max_motif_size=3, min_repeats=3))


def test_empty_and_all_n(prf):
    fs = ns(min_motif_size=1, max_motif_size=50, min_repeats=3, min_span=9)
    assert prf.detect_repeats("", fs) == []
    assert prf.detect_repeats("N" * 100, fs) == []
    assert prf.detect_repeats("A", fs) == []
    assert prf.detect_repeats("a" * 9, fs) == [(0, 9, "A")]


@pytest.mark.parametrize("knobs", KNOB_SETS)
def test_differential_fuzz_vs_oracle(prf, oracle, knobs):
    rng = random.Random(20261018 + len(knobs))
    sizes = [1, 2, 31, 32, 33, 63, 64, 65, 255, 256, 257, 1000, 8191, 8192, 8193, 20000, 70000]
    for n in sizes:
        for rep in range(2):
            seq = random_seq(rng, n, exotic=(rep == 1))
            kmin = rng.choice([1, 1, 2, 5])
            kmax = kmin + rng.choice([0, 5, 49, 49, 70, 130])
            fs = ns(min_motif_size=kmin, max_motif_size=kmax, min_repeats=rng.choice([2, 3, 3, 4]),
                    min_span=rng.choice([1, 6, 9, 9, 12, 40]))
            want = oracle.detect_repeats_by_k(seq, fs)
            got = prf.detect_repeats(seq, fs, **knobs)
            assert got == want, f"n={n} {fs} knobs={knobs}: first diff " \
                                f"{next((a, b) for a, b in zip(got + [None], want + [None]) if a != b)}"


def test_boundary_sweep_every_offset(prf, oracle):
    """A repeat slid across word / strip / tile boundaries one base at a time (T=1: strips are
    one 32-base word, tiles 8192 bases)."""
    rng = random.Random(5)
    fs = ns(min_motif_size=1, max_motif_size=40, min_repeats=3, min_span=9)
    for unit in ["A", "AC", "ACG", "ACGTT", "ACGTTGCA" + "T" * 9, "ACGGTCATTGCAGGTTACAGTCAGTACCGATGCATTGA"]:
        body = unit * (1 + 140 // len(unit))
        for base in (8192 - len(body) - 3, 8192 - 40, 8192 - 1):
            for off in range(0, 70, 1 if len(unit) < 4 else 7):
                left = "".join(rng.choice("ACGT") for _ in range(base + off))
                seq = left + body + "".join(rng.choice("ACGT") for _ in range(300))
                assert prf.detect_repeats(seq, fs, words_per_thread=1) == oracle.detect_repeats_by_k(seq, fs)


def test_long_runs_cross_tiles(prf, oracle):
    """Runs far longer than a tile (config C4 in miniature): homopolymer, dinucleotide, k=50."""
    rng = random.Random(6)
    fs = ns(min_motif_size=1, max_motif_size=50, min_repeats=3, min_span=9)
    units = ["A", "AT", "ACGTG", "".join(rng.choice("ACGT") for _ in range(50))]
    for knobs in ({"words_per_thread": 1}, {}, {"words_per_thread": 1, "walk_limit_words": 2}):
        parts = []
        for u in units:
            parts.append("".join(rng.choice("ACGT") for _ in range(rng.randint(1, 3000))))
            parts.append(u * (rng.randint(9000, 30000) // len(u)) + u[:rng.randint(0, len(u) - 1)] if len(u) > 1
                         else u * rng.randint(9000, 30000))
        parts.append("N" * 77 + "ACGT" * 5000)
        seq = "".join(parts)
        assert prf.detect_repeats(seq, fs, **knobs) == oracle.detect_repeats_by_k(seq, fs)


def test_megabase_default_settings(prf, oracle):
    rng = np.random.default_rng(11)
    n = 3_000_000
    arr = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
    # plant repeats and N runs
    pos = 1000
    while pos < n - 5000:
        k = int(rng.integers(1, 51))
        copies = int(rng.integers(2, 12))
        unit = rng.choice(np.frombuffer(b"ACGT", np.uint8), k)
        rep = np.tile(unit, copies)
        arr[pos:pos + rep.size] = rep
        pos += rep.size + int(rng.integers(50, 3000))
        if rng.random() < 0.02:
            ln = int(rng.integers(1, 5000))
            arr[pos:pos + ln] = ord("N")
            pos += ln
    lower = rng.random(n) < 0.3
    arr = np.where(lower, arr | 0x20, arr).astype(np.uint8)
    seq = arr.tobytes().decode("ascii")
    fs = ns(min_motif_size=1, max_motif_size=50, min_repeats=3, min_span=9)
    want = oracle.detect_repeats_by_k(seq, fs)
    assert len(want) > 1000
    for knobs in ({}, {"words_per_thread": 16}):
        assert prf.detect_repeats(seq, fs, **knobs) == want


def test_large_motif_range_like_the_hail_pipeline(prf, oracle):
    """run_hail_batch_pipeline.py:33 defaults to motif sizes up to 1000: k >> 5 up to 31, halo of 1000 bases."""
    rng = random.Random(77)
    parts = []
    for _ in range(60):
        parts.append("".join(rng.choice("ACGT") for _ in range(rng.randint(50, 4000))))
        unit = "".join(rng.choice("ACGT") for _ in range(rng.choice([1, 3, 33, 64, 97, 255, 256, 500, 999, 1000])))
        parts.append(unit * rng.randint(3, 6) + unit[:rng.randint(0, len(unit) - 1)] if len(unit) > 1 else unit * 40)
    seq = "".join(parts)
    fs = ns(min_motif_size=1, max_motif_size=1000, min_repeats=3, min_span=9)
    want = oracle.detect_repeats_by_k(seq, fs)
    assert any(len(m) >= 500 for _, _, m in want)
    for knobs in ({}, {"words_per_thread": 1}):
        assert prf.detect_repeats(seq, fs, **knobs) == want


def test_knob_settings_agree_and_stream_binding(prf):
    """Every tile shape gives the same rows; the library runs on a caller-provided CUDA stream."""
    import torch
    from crf_b200 import _cabi
    rng = random.Random(123)
    seq = random_seq(rng, 300_000).encode()
    ctx = _cabi.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    ref = None
    with ctx.load(seq, max_motif_cap=64) as s:
        for wpt in (1, 2, 4, 8, 16):
            n = s.scan(1, 64, 3, 9, words_per_thread=wpt)
            rows = [a.tolist() for a in s.fetch(n)]
            ref = ref or rows
            assert rows == ref and n > 100
        st = s.stats()
        assert st.kernel_ms > 0 and st.scan_ms >= st.kernel_ms and st.launches >= 5
        info = s.info()
        assert info.total_bases == len(seq) and info.n_records == 1 and info.max_motif_cap == 64
    ctx.set_stream(0)
    ctx.close()


def test_c_abi_error_paths(prf):
    import ctypes
    import numpy as np
    from crf_b200 import _cabi
    lib = _cabi.lib()
    ctx = _cabi.Context(0)
    with pytest.raises(ValueError):
        ctx.load(b"ACGT", max_motif_cap=0)
    with pytest.raises(NotImplementedError):
        ctx.load(b"ACGT", max_motif_cap=70000)
    with pytest.raises(ValueError):
        _cabi.Context(1234)
    with ctx.load(b"ACGTACGTACGTACGT", max_motif_cap=8) as s:
        with pytest.raises(ValueError, match="exceeds the max_motif_cap"):
            s.scan(1, 9, 3, 9)
        with pytest.raises(ValueError, match="no scan results"):
            s.fetch(1)
        n = s.scan(1, 8, 3, 9)
        assert n == 1
        small = np.zeros(0, np.uint32)
        rc = lib.crf_fetch(s._h, small.ctypes.data, small.ctypes.data, small.ctypes.data, small.ctypes.data, 0, 0)
        assert rc == _cabi.CRF_ERR_CAPACITY and b"capacity" in lib.crf_last_error()
        with pytest.raises(ValueError):
            s.run_end(0, 99, 2)
        assert s.run_end(0, 0, 4) == 12       # M_4 is one for positions 0..11 of ACGTx4
        with pytest.raises(ValueError):
            s.scan(1, 8, 3, 9, words_per_thread=3)
    ctx.close()


def test_non_ascii_text(prf, oracle):
    fs = ns(min_motif_size=1, max_motif_size=5, min_repeats=3, min_span=6)
    seq = "ACGT" + "é" * 7 + "acacacacac"          # latin-1 symbol: an ordinary letter for the reference
    assert prf.detect_repeats(seq, fs) == [(4, 11, "É"), (11, 21, "AC")]
    with pytest.raises(NotImplementedError):
        prf.detect_repeats("ACGT\u0394\u0394\u0394\u0394\u0394\u0394\u0394", fs)


def test_interval_mode_with_long_runs_past_the_interval_end(prf, oracle):
    """Found by tests/fuzz_gpu.py: the stop position of the reference's loop (prf:73-74) when a repeat runs far
    beyond interval_end (several probe windows), with and without motif size 1 being tracked."""
    rng = random.Random(99)
    for case in range(40):
        seq = random_seq(rng, rng.randint(3000, 30000))
        unit = "".join(rng.choice("ACGT") for _ in range(rng.choice([1, 1, 2, 3, 16])))
        pos = rng.randint(0, len(seq))
        seq = seq[:pos] + unit * rng.randint(100, 3000) + seq[pos:]
        a = rng.randint(max(0, pos - 500), pos + 200)
        b = rng.randint(a, min(len(seq), pos + 600))
        fs = ns(min_motif_size=rng.choice([1, 2, 3]), max_motif_size=rng.choice([6, 20, 64]), min_repeats=rng.choice([2, 3]),
                min_span=rng.choice([9, 100]), interval_start_0based=a, interval_end=b)
        try:
            want, exc = oracle.detect_repeats(seq, fs), None
        except AssertionError:
            want, exc = None, AssertionError
        if exc:
            with pytest.raises(exc):
                prf.detect_repeats(seq, fs)
        else:
            assert prf.detect_repeats(seq, fs) == want, f"case {case}: {fs}"
