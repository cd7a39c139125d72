"""The host-side packer (crf_pack_ascii / crf_fasta_packed; no GPU): ASCII text -> the 2-bit planes + not-ACGT mask +
exotic list that crf_seq_load_packed uploads, against a numpy statement of the format in include/crf.h."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))

from crf_b200 import _cabi  # noqa: E402
from tests.fake_ctx import unpack_planes  # noqa: E402


def reference_planes(arr):
    up = arr.copy()
    low = (up >= 97) & (up <= 122)
    up[low] -= 32                                            # str.upper() on ASCII (prf:33)
    n = arr.size
    nw = (n + 31) // 32
    code = np.zeros(n, np.uint8)
    ok = np.zeros(n, bool)
    for i, c in enumerate(b"ACGT"):
        code[up == c] = i
        ok |= up == c
    pad = nw * 32 - n
    weights = np.uint64(1) << np.arange(32, dtype=np.uint64)

    def plane(bits, fill):
        b = np.concatenate([bits.astype(np.uint64), np.full(pad, fill, np.uint64)]).reshape(nw, 32)
        return (b * weights).sum(1).astype(np.uint32)
    exo_pos = np.flatnonzero(~ok & (up != ord("N")))
    exo = (exo_pos.astype(np.uint64) << np.uint64(8)) | up[exo_pos].astype(np.uint64)
    return plane((code >> 1) & ok, 0), plane((code & 1) & ok, 0), plane(~ok, 1), exo


ALPHABET = np.frombuffer(b"ACGTacgtNnRYk-*\xe1", dtype=np.uint8)
WEIGHTS = [.2, .2, .2, .2, .03, .03, .03, .03, .02, .02, .01, .01, .005, .005, .005, .005]


@pytest.mark.parametrize("scalar", [False, True])
def test_pack_ascii_matches_the_documented_format(scalar, monkeypatch):
    if scalar:
        monkeypatch.setenv("CRF_PACK_SCALAR", "1")          # the portable loop instead of AVX2
    rng = np.random.default_rng(3)
    for n in [0, 1, 31, 32, 33, 63, 64, 1000, 100_003, 9_000_017]:
        arr = rng.choice(ALPHABET, size=n, p=WEIGHTS).astype(np.uint8)
        pk = _cabi.pack_ascii(arr, n_threads=4)
        H, L, NM, exo = reference_planes(arr)
        nw = (n + 31) // 32
        assert np.array_equal(pk.H[:nw], H) and np.array_equal(pk.L[:nw], L) and np.array_equal(pk.NM[:nw], NM), n
        assert np.array_equal(pk.exotic, exo), n
        if n:
            up = arr.copy()
            low = (up >= 97) & (up <= 122)
            up[low] -= 32
            is_plain = np.isin(up, np.frombuffer(b"ACGTN", np.uint8))
            back = unpack_planes(pk)
            assert np.array_equal(back, up), n                # round trip: planes + exotic list hold the whole text
            assert (~is_plain).sum() == pk.exotic.size


def test_mask_runs_are_the_maximal_blocks_of_the_plane():
    rng = np.random.default_rng(8)
    for n in [0, 1, 31, 32, 33, 64, 1000, 100_003, 9_000_017]:
        arr = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=n, p=[.24, .24, .24, .24, .04]).astype(np.uint8)
        if n > 5000:
            arr[100:3000] = ord("N")                          # long blocks, one across the piece boundaries of the threads
            arr[n // 2:n // 2 + 70_000] = ord("n")
            arr[n - 777:] = ord("N")                          # ... and one up to the last base
        pk = _cabi.pack_ascii(arr, n_threads=4).with_runs(4)
        masked = np.isin(arr, [ord("N"), ord("n")])
        d = np.diff(np.concatenate([[0], masked.astype(np.int8), [0]]))
        want = np.stack([np.flatnonzero(d == 1), np.flatnonzero(d == -1)], 1).astype(np.uint64)
        assert np.array_equal(pk.runs, want), n
        assert pk.nbytes == 8 * ((n + 31) // 32) + 16 * len(want)


def test_fasta_reader_hands_out_packed_planes(tmp_path):
    rng = np.random.default_rng(5)
    recs = [rng.choice(ALPHABET, size=n, p=WEIGHTS).astype(np.uint8) for n in (70_001, 0, 33, 120_000)]
    recs = [np.where(np.isin(r, [ord("-"), ord("*"), 0xe1]), ord("W"), r).astype(np.uint8) for r in recs]   # printable
    path = tmp_path / "p.fa"
    with open(path, "wb") as f:
        for i, r in enumerate(recs):
            f.write(b">r%d some words\n" % i)
            for j in range(0, r.size, 61):
                f.write(r[j:j + 61].tobytes() + b"\n")
    with _cabi.Fasta(str(path)) as fa:
        pk = fa.packed()
        text = np.concatenate(recs)
        assert fa.total_bases == text.size and np.array_equal(fa.bases, text)
        H, L, NM, exo = reference_planes(text)
        nw = (text.size + 31) // 32
        got = unpack_planes(pk)
        up = text.copy()
        low = (up >= 97) & (up <= 122)
        up[low] -= 32
        assert np.array_equal(got, up)
        assert np.array_equal(pk.exotic, exo) and exo.size > 100
        again = fa.packed()                                   # cached with the handle
        assert again.H_ptr.value == pk.H_ptr.value
        del H, L, NM, nw
