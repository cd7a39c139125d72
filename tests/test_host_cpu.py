"""Host-side pieces that need no GPU: FASTA reader, CLI argument errors, workload generators."""
import gzip
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))

from crf_b200 import cli, fasta, synth  # noqa: E402


def test_fasta_reader_plain_and_gzip(tmp_path):
    text = ">chr1 some description\nACGT\nacgtNN\n\n>chr2\nTTTT\r\nGG\n>empty\n>last x\nA"
    p = tmp_path / "t.fa"
    p.write_text(text)
    want = [("chr1", b"ACGTacgtNN"), ("chr2", b"TTTTGG"), ("empty", b""), ("last", b"A")]
    assert [(r.name, r.seq) for r in fasta.read_fasta(str(p))] == want
    gz = tmp_path / "t.fa.gz"
    with gzip.open(gz, "wt") as f:
        f.write(text)
    assert [(r.name, r.seq) for r in fasta.read_fasta(str(gz))] == want


@pytest.mark.parametrize("argv", [
    ["--min-motif-size", "0", "ACGT"],
    ["--min-motif-size", "5", "--max-motif-size", "4", "ACGT"],
    ["--min-repeats", "0", "ACGT"],
    ["--min-span", "0", "ACGT"],
    ["ACGTXX"],                                  # neither a file nor a nucleotide string (prf:172-173)
    ["--interval", "chr1:0-10", "ACGT"],         # --interval needs a FASTA file (prf:155-156)
])
def test_cli_argument_errors_exit_2(argv, capsys):
    with pytest.raises(SystemExit) as exc:
        cli.main(argv)
    assert exc.value.code == 2


def test_cli_defaults_match_reference():
    args = cli.build_parser().parse_args(["ACGT"])
    assert (args.min_motif_size, args.max_motif_size, args.min_repeats, args.min_span) == (1, 50, 3, 9)
    assert args.interval is None and args.output_prefix is None and args.plot is None


def test_synth_numpy_and_torch_agree_and_are_deterministic():
    a, off, meta = synth.s38(scale=0.0005)
    b, off2, _ = synth.s38(device="cpu", scale=0.0005)
    assert np.array_equal(a, b.numpy()) and np.array_equal(off, off2)
    c, _, _ = synth.s38(scale=0.0005)
    assert np.array_equal(a, c)
    assert set(np.unique(a).tolist()) <= set(b"ACGTNacgt")
    assert len(off) == 25 and meta["names"][0] == "chr1"


def test_synth_read_set_shape():
    bases, offsets, _ = synth.sr(500)
    assert bases.size == 500 * 150 and offsets[1] == 150 and offsets[-1] == 75000


def test_native_row_writer_bed_and_tsv(tmp_path):
    """crf_write_rows is host-only code in libcrf.so: BED (prf:148-149) and TSV (prf:166-170) formats."""
    from crf_b200 import _cabi, build
    build.build()
    bases, offsets = b"ACGTacgtNNacgtttttt", [0, 10, 19]
    bed = tmp_path / "o.bed"
    n = _cabi.write_rows(str(bed), ["chr1", "chrZ"], bases, offsets, [0, 1, 1], [0, 0, 3], [8, 4, 9], [4, 4, 1])
    assert bed.read_text() == "chr1\t0\t8\tACGT\nchrZ\t0\t4\tACGT\nchrZ\t3\t9\tT\n" and n == 39
    tsv = tmp_path / "o.tsv"
    _cabi.write_rows(str(tsv), None, bases, offsets, [0], [4], [8], [2], tsv=True)
    assert tsv.read_text() == "start_0based\tend\tmotif\n4\t8\tAC\n"
    # many rows: buffer flushes
    k = 5000
    _cabi.write_rows(str(bed), ["a"], b"ACGT" * 2000, [0, 8000], [0] * k, list(range(k)), list(range(1, k + 1)), [1] * k)
    lines = bed.read_text().splitlines()
    assert len(lines) == k and lines[-1].startswith("a\t4999\t5000\t")
