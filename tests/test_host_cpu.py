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


def _simple_fasta(raw):
    """Checker: the obvious line-by-line reading of FASTA text (bytes) -> [(name, seq)]."""
    out = []
    for line in raw.split(b"\n"):
        if line.startswith(b">"):
            fields = line[1:].split()
            out.append([fields[0].decode() if fields else "", []])
        elif out:
            out[-1][1].append(line.replace(b"\r", b""))
    return [(n, b"".join(parts)) for n, parts in out]


@pytest.mark.parametrize("eol", [b"\n", b"\r\n"])
def test_fasta_reader_large_multi_piece(tmp_path, eol):
    """Records larger than one 8 MiB reader piece, ragged line widths, several threads, multi-member gzip."""
    from crf_b200 import _cabi
    rng = np.random.default_rng(5)
    chunks = []
    for r, size in enumerate([0, 1, 59, 60, 61, 20_000_000, 3, 9_000_000]):
        seq = rng.choice(np.frombuffer(b"ACGTacgtNn", dtype=np.uint8), size=size).tobytes()
        width = [60, 61, 1, 70, 80, 60, 7, 10_000_000][r]
        chunks.append(b">rec%d description %d" % (r, r) + eol)
        chunks.extend(seq[i:i + width] + eol for i in range(0, size, width))
    raw = b"".join(chunks)[:-len(eol)]                  # no line end after the last line
    want = _simple_fasta(raw)
    assert sum(len(s) for _, s in want) == 29_000_184
    plain = tmp_path / "big.fa"
    plain.write_bytes(raw)
    cut = len(raw) // 3
    gz = tmp_path / "big.fa.gz"
    gz.write_bytes(gzip.compress(raw[:cut], 1) + gzip.compress(raw[cut:], 1))      # two members, like bgzip
    for path in (plain, gz):
        for threads in (1, 5):
            with _cabi.Fasta(str(path), n_threads=threads) as fa:
                assert fa.names == [n for n, _ in want]
                assert fa.total_bases == int(fa.offsets[-1]) == len(fa.bases)
                for i, (_, seq) in enumerate(want):
                    assert fa.record(i).tobytes() == seq, (path, threads, i)


def _bgzf(raw, block=30000, level=6, strategy=0):
    """bgzip's container: gzip members of <= 64 KiB with a 'BC' extra field holding the member's size, then the empty
    end-of-file block."""
    import struct
    import zlib
    out = []
    for i in list(range(0, len(raw), block)) + [None]:
        chunk = b"" if i is None else raw[i:i + block]
        comp = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        data = comp.compress(chunk) + comp.flush()
        bsize = 12 + 6 + len(data) + 8
        out.append(b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) +
                   data + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    return b"".join(out)


def test_fasta_reader_bgzf_parallel_blocks(tmp_path, monkeypatch):
    from crf_b200 import _cabi
    rng = np.random.default_rng(11)
    chunks = []
    for r, size in enumerate([5_000_000, 0, 123, 9_000_000]):
        seq = rng.choice(np.frombuffer(b"ACGTacgtN", dtype=np.uint8), size=size).tobytes()
        chunks.append(b">c%d x\n" % r + b"".join(seq[i:i + 60] + b"\n" for i in range(0, size, 60)))
    raw = b"".join(chunks)
    want = _simple_fasta(raw)
    blob = _bgzf(raw)
    assert gzip.decompress(blob) == raw                      # the container is valid gzip
    path = tmp_path / "ref.fa.bgz"
    path.write_bytes(blob)
    for threads in (1, 6):
        with _cabi.Fasta(str(path), n_threads=threads) as fa:
            assert [(n, fa.record(i).tobytes()) for i, n in enumerate(fa.names)] == want
    # the blocks go through the reader's own DEFLATE decoder; zlib per block (CRF_GUNZIP_ZLIB=1) must give the same, and so
    # must blocks of every DEFLATE type: stored (level 0), fixed codes (Z_FIXED), full-size 65 280-byte blocks
    import zlib
    monkeypatch.setenv("CRF_GUNZIP_ZLIB", "1")
    with _cabi.Fasta(str(path), n_threads=3) as fa:
        assert [(n, fa.record(i).tobytes()) for i, n in enumerate(fa.names)] == want
    monkeypatch.delenv("CRF_GUNZIP_ZLIB")
    for k, (block, level, strategy) in enumerate([(65280, 0, 0), (65280, 6, zlib.Z_FIXED), (65280, 1, 0), (1000, 9, zlib.Z_HUFFMAN_ONLY)]):
        variant = tmp_path / f"v{k}.fa.bgz"
        part = raw[:3_000_000]
        variant.write_bytes(_bgzf(part, block, level, strategy))
        with _cabi.Fasta(str(variant), n_threads=4) as fa:
            assert [(n, fa.record(i).tobytes()) for i, n in enumerate(fa.names)] == _simple_fasta(part)
    # BGZF blocks followed by an ordinary gzip member: not BGZF throughout, read by the serial path
    mixed = tmp_path / "mixed.fa.gz"
    mixed.write_bytes(_bgzf(raw[:1_000_000])[:-28] + gzip.compress(raw[1_000_000:2_000_000], 1))
    with _cabi.Fasta(str(mixed)) as fa:
        assert [(n, fa.record(i).tobytes()) for i, n in enumerate(fa.names)] == _simple_fasta(raw[:2_000_000])
    # a flipped payload byte is caught by the block's CRC
    broken = bytearray(blob)
    broken[len(blob) // 2] ^= 0x55
    bad = tmp_path / "bad.fa.bgz"
    bad.write_bytes(bytes(broken))
    with pytest.raises(ValueError, match="BGZF|gzip"):
        _cabi.Fasta(str(bad))


def test_fasta_reader_edge_cases(tmp_path):
    from crf_b200 import _cabi
    cases = {
        "empty.fa": (b"", []),
        "nohdr.fa": (b"ACGT\nACGT\n", []),
        "hdronly.fa": (b">x", [("x", b"")]),
        "leading.fa": (b"junk\n>  a b\nAC>GT\n>\nTT\n", [("a", b"AC>GT"), ("", b"TT")]),
    }
    for name, (raw, want) in cases.items():
        (tmp_path / name).write_bytes(raw)
        with _cabi.Fasta(str(tmp_path / name)) as fa:
            assert [(n, fa.record(i).tobytes()) for i, n in enumerate(fa.names)] == want, name
    with pytest.raises(ValueError):
        _cabi.Fasta(str(tmp_path / "missing.fa"))
    bad = tmp_path / "bad.fa.gz"
    bad.write_bytes(gzip.compress(b">a\nACGT\n" * 1000)[:-20])
    with pytest.raises(ValueError):
        _cabi.Fasta(str(bad))


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
    # file -> reader -> writer without Python strings in between: the reader's NUL-separated name table and base buffer
    fa_path = tmp_path / "w.fa"
    fa_path.write_bytes(b">chr1 first\nACGTacgt\nNN\n>chrZ\nacgtttttt\n")
    with _cabi.Fasta(str(fa_path)) as fa:
        assert fa.names_blob == b"chr1\0chrZ\0"
        _cabi.write_rows(str(bed), fa.names_blob, fa.bases, fa.offsets, [0, 1, 1], [0, 0, 3], [8, 4, 9], [4, 4, 1])
    assert bed.read_text() == "chr1\t0\t8\tACGT\nchrZ\t0\t4\tACGT\nchrZ\t3\t9\tT\n"


def test_gunzip_decoder_against_zlib():
    """The FASTA reader's own DEFLATE decoder (csrc/crf_inflate.h, crf_gunzip mode 2 = no zlib behind it) against Python's zlib:
    stored / fixed / dynamic blocks, Huffman-only and RLE strategies, every level and window size, flush points inside the
    stream, FNAME / MTIME header fields, several members, inputs from 0 bytes to 600 kB of DNA, text, runs, zeros and noise.
    A flipped bit or a cut stream is an error (or, where the flip is harmless, the same bytes zlib gives) -- never other bytes."""
    import gzip
    import io
    import random
    import zlib
    from crf_b200 import _cabi
    rng = random.Random(20261018)

    def gz(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=31, memlevel=8, chunks=None):
        c = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
        out, pos = b"", 0
        for n in chunks or []:
            out += c.compress(data[pos:pos + n]) + c.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_NO_FLUSH]))
            pos += n
        return out + c.compress(data[pos:]) + c.flush()

    def make(n):
        kind = rng.choice(["dna", "soft-masked", "text", "zeros", "noise", "runs", "fasta", "two letters", "periodic"])
        if kind == "dna":
            return bytes(rng.choice(b"ACGT") for _ in range(n))
        if kind == "soft-masked":
            return bytes(rng.choice(b"ACGTacgtN") for _ in range(n))
        if kind == "text":
            return " ".join(rng.choice(["the", "quick", "brown", "fox", "jumps", "over", "lazy", "dog", "\n"])
                            for _ in range(n // 4 + 1)).encode()[:n]
        if kind == "zeros":
            return bytes(n)
        if kind == "noise":
            return rng.randbytes(n)
        if kind == "runs":
            return b"".join(bytes([rng.randrange(256)]) * rng.randint(1, 600) for _ in range(n // 100 + 1))[:n]
        if kind == "fasta":
            s = bytes(rng.choice(b"ACGT") for _ in range(n))
            return b">chr\n" + b"\n".join(s[i:i + 60] for i in range(0, len(s), 60))
        if kind == "two letters":
            return bytes(rng.choice(b"AB") for _ in range(n))
        unit = rng.randbytes(rng.randint(1, 40))
        return (unit * rng.randint(1, 3000) + rng.randbytes(rng.randint(0, 3000)))[:max(n, 1)]

    n_cases = n_rejected = 0
    for _ in range(260):
        data = make(rng.choice([0, 1, 2, rng.randint(0, 300), rng.randint(300, 70_000), rng.randint(70_000, 600_000)]))
        mode = rng.random()
        if mode < 0.15:
            comp = gz(data, level=0)
        elif mode < 0.3:
            comp = gz(data, level=rng.randint(1, 9), strategy=zlib.Z_FIXED)
        elif mode < 0.4:
            comp = gz(data, level=rng.randint(1, 9), strategy=zlib.Z_HUFFMAN_ONLY)
        elif mode < 0.5:
            comp = gz(data, level=rng.randint(1, 9), strategy=zlib.Z_RLE)
        elif mode < 0.6:
            comp = gz(data, level=rng.randint(1, 9), chunks=[rng.randint(0, max(1, len(data) // 3)) for _ in range(3)])
        elif mode < 0.7:
            comp = gz(data, level=rng.randint(1, 9), memlevel=rng.randint(1, 9), wbits=16 + rng.randint(9, 15))
        elif mode < 0.8:
            b = io.BytesIO()
            with gzip.GzipFile(filename="some_name.fa", mode="wb", fileobj=b, compresslevel=rng.randint(1, 9),
                               mtime=rng.randint(0, 2 ** 31)) as f:
                f.write(data)
            comp = b.getvalue()
        else:
            comp = gz(data, level=rng.randint(1, 9))
        want = data
        if rng.random() < 0.25:                                         # a second member
            more = make(rng.randint(0, 5000))
            comp, want = comp + gz(more, level=rng.randint(0, 9)), data + more
        assert _cabi.gunzip(comp, use_zlib=2) == want                   # the reader's decoder alone
        assert _cabi.gunzip(comp) == want and _cabi.gunzip(comp, use_zlib=1) == want
        n_cases += 1
        if len(comp) > 20:
            bad = bytearray(comp)
            bad[rng.randrange(10, len(bad))] ^= 1 << rng.randrange(8)
            try:
                got = _cabi.gunzip(bytes(bad), use_zlib=2)
                ref = zlib.decompressobj(31)
                try:
                    z = ref.decompress(bytes(bad))
                except zlib.error:
                    z = None
                if z is not None and ref.eof and not ref.unused_data:
                    assert got == z
            except NotImplementedError:
                n_rejected += 1
            try:                                                        # a cut stream: an error, unless the cut is the end of member one
                cut = _cabi.gunzip(comp[:rng.randrange(0, len(comp) - 1)], use_zlib=rng.choice([0, 2]))
                assert cut == data and want != data
            except (ValueError, NotImplementedError):
                pass
    assert n_cases == 260 and n_rejected > 150


def test_gunzip_one_stream_on_several_threads(monkeypatch, capfd):
    """A plain gzip stream decoded by several threads (crf_inflate.h: gunzip_parallel -- chunks that find a block boundary by
    trial, decode into 16-bit marker symbols without knowing the 32 KB before them, and are stitched and resolved afterwards).
    Small chunk sizes (CRF_GUNZIP_CHUNK_KB) send streams of 0.3 - 2 MB through it: DNA, text, runs, zeros and noise mixed in
    one stream (dynamic, stored and -- Z_FIXED -- fixed blocks, flush points); the bytes must be zlib's, the trace must show
    that the threaded path did the work, and that the case "a chunk starts later than the previous one stopped" came up."""
    import random
    import re
    import zlib
    from crf_b200 import _cabi
    rng = random.Random(77)
    monkeypatch.setenv("CRF_GUNZIP_TRACE", "1")

    def piece(kind, n):
        if kind == "dna":
            return bytes(rng.choice(b"ACGT") for _ in range(n))
        if kind == "soft-masked":
            return bytes(rng.choice(b"ACGTacgtN\n") for _ in range(n))
        if kind == "text":
            return " ".join(rng.choice(["the", "quick", "brown", "fox", "jumps", "over", "lazy", "dog", "\n"])
                            for _ in range(n // 4 + 1)).encode()[:n]
        if kind == "zeros":
            return bytes(n)
        if kind == "noise":
            return rng.randbytes(n)
        unit = rng.randbytes(rng.randint(1, 40))
        return (unit * (n // len(unit) + 1))[:n]

    n_parallel = n_chunks = n_carried = 0
    for case in range(30):
        if case < 4:                                     # stored blocks in the middle: chunks that start there find a later block
            kinds = ["dna", "noise", "dna", "noise", "soft-masked"]
        else:
            kinds = [rng.choice(["dna", "soft-masked", "text", "zeros", "noise", "periodic"]) for _ in range(rng.randint(1, 6))]
        data = b"".join(piece(k, rng.randint(50_000, 400_000)) for k in kinds)
        strategy = rng.choice([zlib.Z_DEFAULT_STRATEGY] * 5 + [zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE])
        c = zlib.compressobj(rng.randint(1, 9), zlib.DEFLATED, 31, rng.randint(4, 9), strategy)
        comp, pos = b"", 0
        for _ in range(rng.choice([0, 0, 2])):
            n = rng.randint(0, len(data) // 3)
            comp += c.compress(data[pos:pos + n]) + c.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH]))
            pos += n
        comp += c.compress(data[pos:]) + c.flush()
        monkeypatch.setenv("CRF_GUNZIP_CHUNK_KB", str(rng.choice([4, 16, 64])))
        capfd.readouterr()
        assert _cabi.gunzip(comp, use_zlib=2) == data, (case, kinds)
        m = re.search(r"parallel: (\d+) round\(s\), (\d+) marker chunk\(s\) stitched, (\d+) carried", capfd.readouterr().err)
        if m:
            n_parallel += 1
            n_chunks += int(m.group(2))
            n_carried += int(m.group(3))
    assert n_parallel >= 20 and n_chunks >= 60 and n_carried >= 3


def test_gunzip_fuzz_under_sanitizers(tmp_path):
    """tests/fuzz_inflate.cpp built with AddressSanitizer + UBSan against csrc/crf_inflate.h: random zlib-made streams (a
    quarter with flipped bits) through the several-threads decoder with tiny chunks -- decoders that start on a phantom block
    boundary decode garbage, and must do so without touching a byte outside their buffers."""
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("no g++")
    here = os.path.dirname(os.path.abspath(__file__))
    csrc = os.path.join(os.path.dirname(here), "colab-repeat-finder_b200", "csrc")
    exe = str(tmp_path / "fuzz_inflate")
    # (a chunk budget of 3 x the expected output and no floor: the N runs of the test data outgrow it, so "a chunk stops, the
    # decoder in front covers its range with the same budget again, else one thread does the job" is exercised too)
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-pthread",
           "-DCRF_INFLATE_MARKER_BUDGET=1", "-I", csrc, os.path.join(here, "fuzz_inflate.cpp"), "-o", exe, "-lz"]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0:
        pytest.skip("sanitizer build not available here: " + built.stderr[-300:])
    run = subprocess.run([exe, "2026", "24"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-4000:]
    assert "24 cases" in run.stdout


def test_reader_and_packer_under_sanitizers(tmp_path, monkeypatch):
    """tests/fuzz_reader.cpp: csrc/crf_fasta.h + csrc/crf_pack.h + csrc/crf_rows.h compiled without the CUDA runtime under ASan + UBSan.  Plain,
    gzip (several threads, small chunks), multi-member and BGZF files with ragged lines, CRLF, IUPAC letters, empty records and
    30 000 short reads: the harness checks the planes, the exotic list and the mask runs against the text base by base; its
    record table (name, length, FNV-1a of the bases) is compared here with the line-by-line reading of the same bytes."""
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("no g++")
    here = os.path.dirname(os.path.abspath(__file__))
    csrc = os.path.join(os.path.dirname(here), "colab-repeat-finder_b200", "csrc")
    exe = str(tmp_path / "fuzz_reader")
    built = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                            "-pthread", "-I", csrc, "-I", os.path.join(os.path.dirname(here), "include"),
                            os.path.join(here, "fuzz_reader.cpp"), "-o", exe, "-lz"], capture_output=True, text=True)
    if built.returncode != 0:
        pytest.skip("sanitizer build not available here: " + built.stderr[-300:])
    rng = np.random.default_rng(2027)

    def text(n_records, max_len, alphabet, eol):
        parts = []
        for r in range(n_records):
            size = int(rng.integers(0, max_len + 1))
            seq = rng.choice(np.frombuffer(alphabet, dtype=np.uint8), size=size).tobytes()
            width = int(rng.choice([60, 61, 70, 1, 1_000_000]))
            parts.append(b">r%d some text" % r + eol)
            parts.extend(seq[i:i + width] + eol for i in range(0, size, width))
        return b"".join(parts)

    files = {}
    big = text(5, 3_500_000, b"ACGTacgtNNn", b"\n")
    files["big.fa"] = big
    files["big.fa.gz"] = gzip.compress(big, 1)
    files["big.fa.bgz"] = _bgzf(big, 65280)
    files["members.fa.gz"] = gzip.compress(big[:4_000_000], 6) + gzip.compress(big[4_000_000:], 1)
    files["iupac_crlf.fa"] = text(40, 30_000, b"ACGTRYKMSWBDHVNacgtn", b"\r\n")
    reads = text(30_000, 150, b"ACGTN", b"\n")
    files["reads.fa"] = reads
    files["reads.fa.gz"] = gzip.compress(reads, 4)
    for name, raw in files.items():
        (tmp_path / name).write_bytes(raw)
    monkeypatch.setenv("CRF_GUNZIP_CHUNK_KB", "256")
    monkeypatch.setenv("FUZZ_READER_ROWS", "150000")    # the row writer too (csrc/crf_rows.h), on several threads
    run = subprocess.run([exe] + [str(tmp_path / n) for n in files], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-4000:]
    # the BED / TSV files the harness wrote for one input, against the same rows formatted here
    recs = [(n, seq) for n, seq in _simple_fasta(files["iupac_crlf.fa"])]
    bed, tsv, i = [], ["start_0based\tend\tmotif\n"], 0
    while len(bed) < 150000:
        name, seq = recs[i % len(recs)]
        if seq:
            k = min(1 + i % 50, len(seq))
            st = (i * 7919) % (len(seq) - k + 1)
            motif = seq[st:st + k].decode().upper()
            bed.append(f"{name}\t{st}\t{st + 3 * k}\t{motif}\n")
            tsv.append(f"{st}\t{st + 3 * k}\t{motif}\n")
        i += 1
    assert (tmp_path / "iupac_crlf.fa.bed").read_text() == "".join(bed)
    assert (tmp_path / "iupac_crlf.fa.tsv").read_text() == "".join(tsv)

    def fnv(seq):
        h = 1469598103934665603
        for block in range(0, len(seq), 1 << 16):       # (python ints: keep the loop out of the per-byte path where possible)
            for c in seq[block:block + (1 << 16)]:
                h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h

    plain_of = {"big.fa": big, "big.fa.gz": big, "big.fa.bgz": big, "members.fa.gz": big, "iupac_crlf.fa": files["iupac_crlf.fa"],
                "reads.fa": reads, "reads.fa.gz": reads}
    lines = [ln for ln in run.stdout.splitlines() if " records " in ln]
    assert len(lines) == 2 * len(files)
    want_cache = {}
    for ln in lines:
        head, _, table = ln.partition(";")
        name = os.path.basename(head.split(" threads ")[0])
        if name not in want_cache:
            recs = _simple_fasta(plain_of[name])
            hashed = name in ("iupac_crlf.fa",)          # hash one file in full (pure-Python FNV is slow), lengths for all
            want_cache[name] = (recs, hashed)
        recs, hashed = want_cache[name]
        got = table.split()
        assert len(got) == len(recs), name
        for entry, (rec_name, seq) in zip(got, recs):
            g_name, g_len, g_hash = entry.rsplit(":", 2)
            assert g_name == rec_name and int(g_len) == len(seq), (name, entry)
            if hashed:
                assert int(g_hash, 16) == fnv(seq), (name, entry)
    # the same bytes whatever the container and the thread count
    by_file = {}
    for ln in lines:
        head, _, table = ln.partition(";")
        by_file.setdefault(os.path.basename(head.split(" threads ")[0]), set()).add(table)
    assert all(len(v) == 1 for v in by_file.values())
    assert by_file["big.fa"] == by_file["big.fa.gz"] == by_file["big.fa.bgz"] == by_file["members.fa.gz"]
    assert by_file["reads.fa"] == by_file["reads.fa.gz"]


def test_fasta_reader_gzip_decoders_agree(tmp_path, monkeypatch):
    """A gzip FASTA through the reader's own decoder and through zlib alone (CRF_GUNZIP_ZLIB=1): same records, same bases."""
    import gzip
    import random
    from crf_b200 import _cabi
    rng = random.Random(3)
    recs = [(f"r{i} x", "".join(rng.choice("ACGTacgtN") for _ in range(rng.randint(0, 200_000)))) for i in range(5)]
    text = "".join(f">{n}\n" + "".join(s[i:i + 70] + "\n" for i in range(0, len(s), 70)) for n, s in recs)
    path = tmp_path / "x.fa.gz"
    for level in (1, 6, 9):
        with gzip.open(path, "wt", compresslevel=level) as f:
            f.write(text)
        out = []
        for use_zlib in (False, True):
            if use_zlib:
                monkeypatch.setenv("CRF_GUNZIP_ZLIB", "1")
            else:
                monkeypatch.delenv("CRF_GUNZIP_ZLIB", raising=False)
            with _cabi.Fasta(str(path)) as fa:
                out.append((list(fa.names), fa.offsets.tolist(), fa.bases.tobytes()))
        assert out[0] == out[1]
        assert out[0][0] == [n.split()[0] for n, _ in recs] and out[0][2] == "".join(s for _, s in recs).encode()


def test_native_row_writer_many_rows_threaded(tmp_path):
    """Above 131 072 rows crf_write_rows formats slices of rows on several threads and writes the buffers in row order: 300 000
    rows (ragged motif sizes, three records with names of different length, lower-case text) against a Python formatter,
    BED, appended BED and TSV; a row count that is not a multiple of the slice size."""
    import random
    from crf_b200 import _cabi
    rng = random.Random(12)
    text = "".join(rng.choice("ACGTacgtN") for _ in range(30_000)).encode()
    offsets = [0, 9_000, 9_000, 30_000]                       # the middle record is empty
    names = ["chr1", "an_empty_record", "chrUn_KI270742v1"]
    n = 300_001
    rec = np.sort(np.array([rng.choice([0, 2]) for _ in range(n)], dtype=np.uint32))
    length = np.where(rec == 0, 9_000, 21_000)
    k = np.array([rng.randint(1, 50) for _ in range(n)], dtype=np.uint32)
    start = np.array([rng.randint(0, int(length[i]) - 60) for i in range(n)], dtype=np.uint32)
    end = start + k * 3

    def expected(tsv):
        out = []
        for r, s0, e0, k0 in zip(rec.tolist(), start.tolist(), end.tolist(), k.tolist()):
            motif = text[offsets[r] + s0:offsets[r] + s0 + k0].decode().upper()
            out.append(f"{s0}\t{e0}\t{motif}\n" if tsv else f"{names[r]}\t{s0}\t{e0}\t{motif}\n")
        return "".join(out)

    bed = tmp_path / "big.bed"
    nbytes = _cabi.write_rows(str(bed), names, text, offsets, rec, start, end, k)
    want = expected(False)
    assert bed.read_text() == want and nbytes == len(want)
    _cabi.write_rows(str(bed), names, text, offsets, rec[:5], start[:5], end[:5], k[:5], append=True)
    assert bed.read_text() == want + "".join(want.splitlines(keepends=True)[:5])
    tsv = tmp_path / "big.tsv"
    _cabi.write_rows(str(tsv), None, text, offsets, rec, start, end, k, tsv=True)
    assert tsv.read_text() == "start_0based\tend\tmotif\n" + expected(True)


# ---- min_repeats == 1: the host half (position-0 wrap-around, early-break selection) with the GPU scan replaced by
# ---- a numpy statement of what crf_scan documents for that setting

def _runs_single_copy(s, kmin, kmax, min_span):
    """include/crf.h, crf_scan with min_repeats == 1: maximal runs of >= max(min_span-k, k-1) matches, motif N-free
    and primitive."""
    from crf_b200 import api
    L, arr = len(s), np.frombuffer(s, dtype=np.uint8)
    rows = []
    for k in range(kmin, min(kmax, L - 1) + 1):
        m = (arr[:L - k] == arr[k:]) & (arr[:L - k] != ord("N"))
        edge = np.diff(np.concatenate([[0], m.astype(np.int8), [0]]))
        for st, i0 in zip(np.flatnonzero(edge == 1).tolist(), np.flatnonzero(edge == -1).tolist()):
            if i0 - st >= max(min_span - k, k - 1, 1) and s[st + k - 1] != ord("N") and api._is_primitive(s[st:st + k]):
                rows.append((st, i0 + k, k))
    rows.sort()
    cols = np.array(rows, dtype=np.uint32).reshape(-1, 3)
    return cols[:, 0], cols[:, 1], cols[:, 2]


def _stop_literal(s, end_position, kmin, kmax):
    """prf:66-74 word for word, trackers reduced to their run length."""
    L = len(s)
    run = {k: 1 for k in range(kmin, kmax + 1)}
    for t in range(L):
        for k in run:
            if t < L - k:
                run[k] = run[k] + 1 if (s[t] == s[t + k] and s[t] != ord("N")) else 1
        if t > end_position and not any(v >= k + 1 for k, v in run.items()):
            return t
    return None


def test_single_copy_host_logic_against_reference_vectors(monkeypatch):
    from crf_b200 import api
    from tests.helpers import load_golden, ns
    monkeypatch.setattr(api, "get_context", lambda device=None: None)
    monkeypatch.setattr(api, "scan_arrays",
                        lambda s, kmin, kmax, mr, span, device=None, **kw: _runs_single_copy(bytes(s), kmin, kmax, span))
    monkeypatch.setattr(api, "_interval_stop_position",
                        lambda ctx, s, e, kmin, kmax, knobs: _stop_literal(bytes(s), e, kmin, kmax))
    checked = refused = raising = 0
    for case in load_golden("fuzz_minrep1.json"):
        fs = ns(**case["settings"])
        if fs.min_motif_size == 1 and fs.min_span <= 1:
            with pytest.raises(NotImplementedError):
                api.detect_repeats(case["seq"], fs)
            refused += 1
            continue
        try:
            got, exc = api.detect_repeats(case["seq"], fs), None
        except (AssertionError, IndexError) as e:
            got, exc = None, type(e).__name__
        assert exc == case.get("raises"), case["settings"]
        if exc is None:
            assert got == [tuple(r) for r in case["result"]], case["settings"]
        raising += exc is not None
        checked += 1
    assert checked >= 280 and raising >= 10 and refused <= 15


def test_fasta_reader_random_files(tmp_path):
    """Seeded fuzz of the native reader against the line-by-line checker: ragged widths, LF / CRLF mixed, blank lines,
    '>' inside sequence lines, empty records, text before the first header; plain, gzip and BGZF containers."""
    import random
    from crf_b200 import _cabi
    rng = random.Random(2026)
    for case in range(120):
        parts = []
        if rng.random() < 0.2:
            parts.append(b"stray text\n")
        for r in range(rng.randint(0, 6)):
            eol = rng.choice([b"\n", b"\r\n"])
            parts.append(b">" + rng.choice([b"", b" ", b"\t"]) + b"r%d" % r + rng.choice([b"", b" desc here", b"\tx"]) + eol)
            for _ in range(rng.randint(0, 8)):
                width = rng.choice([0, 1, 7, 60, 61, 200])
                line = bytes(rng.choice(b"ACGTacgtNnRY>") for _ in range(width))
                if line.startswith(b">"):
                    line = b"A" + line[1:]
                parts.append(line + rng.choice([eol, b"\n"]))
        raw = b"".join(parts)
        if raw and rng.random() < 0.5:
            raw = raw.rstrip(b"\r\n")                     # no line end after the last line
        want = _simple_fasta(raw)
        container = rng.choice(["plain", "gzip", "bgzf"])
        blob = raw if container == "plain" else gzip.compress(raw) if container == "gzip" else _bgzf(raw, rng.choice([50, 4000]))
        path = tmp_path / f"f{case}.fa"
        path.write_bytes(blob)
        with _cabi.Fasta(str(path), n_threads=rng.choice([1, 3])) as fa:
            got = [(n, fa.record(i).tobytes()) for i, n in enumerate(fa.names)]
        assert got == want, (case, container, raw[:200])


def test_cli_whole_fasta_glue_with_a_stand_in_scan(tmp_path, monkeypatch, capsys):
    """The CLI's FASTA -> BED plumbing (native reader, grouping of records into loads, native writer, stdout lines) with
    the GPU scan replaced by tests/fake_ctx.py (the oracle).  The real scan is covered by the -m gpu CLI test."""
    import random
    from crf_b200 import api
    from tests.fake_ctx import FakeContext
    from tests.helpers import ns, random_seq
    from oracle import oracle
    rng = random.Random(4)
    recs = [("chrA", random_seq(rng, 3000)), ("chrB extra words", random_seq(rng, 50)), ("empty", ""),
            ("chrC", random_seq(rng, 7000)), ("chrD", random_seq(rng, 900))]
    fa = tmp_path / "glue.fasta"
    fa.write_text("".join(f">{n}\n" + "".join(s[i:i + 50] + "\n" for i in range(0, len(s), 50)) for n, s in recs))
    want = []
    for name, seq in recs:
        fs = ns(min_motif_size=1, max_motif_size=12, min_repeats=3, min_span=9)
        want += [f"{name.split()[0]}\t{s}\t{e}\t{m}\n" for s, e, m in oracle.detect_repeats(seq, fs)]
    monkeypatch.setattr(api, "get_context", lambda device=None: FakeContext())
    monkeypatch.chdir(tmp_path)
    from crf_b200 import _cabi
    real_limit = _cabi.load_limit(12)
    assert 4_200_000_000 < real_limit < 2 ** 32           # what libcrf itself enforces (crf_load_limit)
    for limit in (real_limit, 3500, 1):                   # one load / several groups / one record per load
        monkeypatch.setattr(_cabi, "load_limit", lambda cap, limit=limit: limit)
        assert cli.main([str(fa), "-max", "12", "-o", str(tmp_path / "sub" / "glue_out")]) == 0
        assert (tmp_path / "glue_out.bed").read_text() == "".join(want)          # basename(prefix), in the CWD (prf:116)
        out = capsys.readouterr().out
        assert "Processing chrB (50 bp)" in out and "Processing empty (0 bp)" in out and "Found 0 repeats" in out
        assert out.rstrip().endswith("Wrote results to glue_out.bed")
    assert len(want) > 30


def test_cli_groups_records_by_layout_positions_not_bases():
    """A load lays every record out as length + max_motif_size positions (inter-record gap), so many short records need
    far more positions than bases: 150-bp reads with -max 50 fill a load at 3/4 of its bases.  The groups must respect
    the library's own limit (crf_load_limit), whatever the record shape."""
    from crf_b200 import _cabi
    limit = _cabi.load_limit(50)
    n = 30_000_000                                          # 4.5 Gbp of reads -> 6.0e9 positions: two loads
    groups = cli.group_records(np.full(n, 150, dtype=np.int64), 50)
    assert len(groups) == 2 and groups[0][0] == 0 and groups[-1][1] == n
    for first, last in groups:
        assert (last - first) * 200 <= limit
    assert groups[0][1] * 200 + 200 > limit                 # ... and the first one is as full as it can be
    # patched small limit, ragged records, one record larger than the limit stays alone
    lengths = [10, 500, 20, 20, 20, 3000, 5]
    groups = cli.group_records(lengths, 50, limit=600)
    assert groups == [(0, 1), (1, 2), (2, 5), (5, 6), (6, 7)]
    assert cli.group_records([], 50) == []


def test_tracker_module_compat():
    """utils.perfect_repeat_tracker stays importable: the string helper answers like the reference's (vectors generated
    from trk:108-142 by tests/golden/make_golden.py, also used by api._is_primitive), the tracker class keeps the reference's constructor (its scan is covered by the -m gpu tests)."""
    from crf_b200 import api
    from tests.helpers import load_golden
    from utils.perfect_repeat_tracker import PerfectRepeatTracker, consists_of_perfect_repeats
    cases = load_golden("primitivity.json")
    assert len(cases) == 600 and sum(c["unit"] is not None for c in cases) > 200
    for c in cases:
        assert consists_of_perfect_repeats(c["seq"]) == c["unit"], c
        assert api._is_primitive(c["seq"].encode()) == (c["unit"] is None), c
    with pytest.raises(NotImplementedError, match="detect_repeats"):
        PerfectRepeatTracker(3, 1, 9, "ACGT", {})           # the single-copy corner is detect_repeats()'s
    t = PerfectRepeatTracker(3, 3, 9, "ACGT", {})            # (the GPU scan behind it only runs on first use)
    assert t.current_position == 0 and t.motif_size == 3


# ---- interval mode: the host half (N-trim, probe scan -> stop position, cut, selection, AssertionError / IndexError) with
# ---- the scan replaced by a numpy statement of the closed form crf_scan documents

class _ClosedFormSeq:
    def __init__(self, raw):
        self.s = bytes(raw).upper()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass

    def scan(self, kmin, kmax, min_repeats, min_span, flags=0, **knobs):
        from crf_b200 import _cabi, api
        s, L = self.s, len(self.s)
        arr = np.frombuffer(s, dtype=np.uint8)
        rows = []
        for k in range(kmin, min(kmax, L - 1) + 1):
            m = (arr[:L - k] == arr[k:]) & (arr[:L - k] != ord("N"))
            edge = np.diff(np.concatenate([[0], m.astype(np.int8), [0]]))
            rmin = max(min_span - k, (min_repeats - 1) * k, 1)
            for st, i0 in zip(np.flatnonzero(edge == 1).tolist(), np.flatnonzero(edge == -1).tolist()):
                if i0 - st >= rmin and ((flags & _cabi.SCAN_NO_PRIMITIVITY) or api._is_primitive(s[st:st + k])):
                    rows.append((st, i0 + k, k))
        rows.sort()
        self.rows = np.array(rows, dtype=np.uint32).reshape(-1, 3)
        return len(rows)

    def fetch(self, n):
        return np.zeros(n, np.uint32), self.rows[:, 0], self.rows[:, 1], self.rows[:, 2]


class _ClosedFormCtx:
    def load(self, raw, offsets=None, max_motif_cap=50, on_device=False):
        return _ClosedFormSeq(raw)


@pytest.mark.parametrize("name", ["kat.json", "fuzz_full.json", "fuzz_interval.json", "fuzz_interval_long.json"])
def test_detect_repeats_host_logic_against_reference_vectors(monkeypatch, name):
    from crf_b200 import api
    from tests.helpers import load_golden, ns
    monkeypatch.setattr(api, "get_context", lambda device=None: _ClosedFormCtx())
    n_rows = n_exc = 0
    for i, case in enumerate(load_golden(name)):
        fs = ns(**case["settings"])
        try:
            got, exc = api.detect_repeats(case["seq"], fs), None
        except (AssertionError, IndexError, ValueError, AttributeError) as e:
            got, exc = None, type(e).__name__
        assert exc == case.get("raises"), (name, i, case["settings"])
        if exc is None:
            assert got == [tuple(r) for r in case["result"]], (name, i, case["settings"])
            n_rows += len(got)
        n_exc += exc is not None
    assert n_rows > 30


def test_interval_ending_near_the_end_of_the_sequence(monkeypatch):
    """interval_end within 2 * max_motif_size of the END of a periodic tail: trackers of large motif sizes have stopped
    (trk:50) but still count as mid-repeat, so prf:70-74 runs the loop to the end.  The probe window of
    api._interval_stop_position must reach back 2k from the end for them; vectors from the reference, all three
    min_repeats modes, then the stop position alone against the literal loop on seeded strings."""
    import random
    from crf_b200 import api
    from tests.helpers import load_golden, ns
    monkeypatch.setattr(api, "get_context", lambda device=None: _ClosedFormCtx())
    closed_form_scan = api.scan_arrays

    def scan_arrays(s, kmin, kmax, min_repeats, span, device=None, **kw):
        if min_repeats == 1:
            return _runs_single_copy(bytes(s), kmin, kmax, span)
        return closed_form_scan(s, kmin, kmax, min_repeats, span, device=device, **kw)

    monkeypatch.setattr(api, "scan_arrays", scan_arrays)
    n_rows = n_single = 0
    for i, case in enumerate(load_golden("fuzz_interval_tail.json")):
        fs = ns(**case["settings"])
        assert "raises" not in case
        assert api.detect_repeats(case["seq"], fs) == [tuple(r) for r in case["result"]], (i, case["settings"])
        n_rows += len(case["result"])
        n_single += fs.min_repeats == 1
    assert n_rows > 1000 and n_single > 40
    rng = random.Random(5)
    n_to_end = 0
    for _ in range(1500):
        s = "".join(rng.choice(rng.choice(["AC", "ACGT", "ACGTN"])) for _ in range(rng.randint(2, 300)))
        if rng.random() < 0.7:
            unit = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 40)))
            pos = rng.randint(0, len(s))
            s = s[:pos] + unit * rng.randint(2, 30) + ("" if rng.random() < 0.5 else s[pos:])
        kmin = rng.choice([1, 2, 3, 7])
        kmax = kmin + rng.choice([0, 1, 5, 15, 30, 49])
        E = rng.choice([rng.randint(0, len(s)), max(0, len(s) - rng.randint(0, 2 * kmax + 2))])
        want = _stop_literal(s.encode(), E, kmin, kmax)
        assert api._interval_stop_position(_ClosedFormCtx(), s.encode(), E, kmin, kmax, {}) == want, (s, E, kmin, kmax)
        n_to_end += want is None and E < len(s) - 1
    assert n_to_end > 100


def test_write_rows_reports_a_short_write(tmp_path):
    """crf_write_rows (host only): rows as text, and a device that takes no more bytes is an OSError, not a silently
    truncated BED file."""
    from crf_b200 import _cabi
    bases = np.frombuffer(b"ACGTACGTACGTttttttttNN", dtype=np.uint8)
    path = tmp_path / "w.bed"
    n = _cabi.write_rows(str(path), ["chrA"], bases, [0, 22], [0, 0], [0, 12], [12, 20], [4, 1])
    assert path.read_text() == "chrA\t0\t12\tACGT\nchrA\t12\t20\tT\n" and n == 28
    n = _cabi.write_rows(str(path), None, bases, [0, 22], [0], [12], [20], [2], tsv=True)
    assert path.read_text() == "start_0based\tend\tmotif\n12\t20\tTT\n" and n == 23 + 9
    if os.path.exists("/dev/full"):
        with pytest.raises(OSError, match="failed"):
            _cabi.write_rows("/dev/full", ["chrA"], bases, [0, 22], [0] * 100_000, [0] * 100_000, [12] * 100_000, [4] * 100_000)


GOLDEN_BED = "/root/reference/benchmark/repeat_finder/chr22_repeats.bed"


@pytest.mark.skipif(not os.path.isfile(GOLDEN_BED), reason="the reference checkout is not on this machine")
def test_reference_golden_bed_obeys_the_closed_form():
    """The one full-size output the reference ships (benchmark/repeat_finder/chr22_repeats.bed: chr22, motif 1-6,
    --min-repeats 3 --min-span 9; its input chr22.fa.gz is missing from the checkout).  Every row must satisfy what the
    kernels compute: sorted by (start, end), unique keys, span >= max(min_span, min_repeats * k), a
    primitive N-free motif, and the repeat really being periodic cannot be checked without the bases -- the rest pins our
    reading of the BED format and the filters."""
    from utils.perfect_repeat_tracker import consists_of_perfect_repeats
    rows = [ln.rstrip("\n").split("\t") for ln in open(GOLDEN_BED)]
    assert len(rows) == 67638 and all(len(r) == 4 and r[0] == "chr22" for r in rows)
    keys = [(int(r[1]), int(r[2])) for r in rows]
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
    for (s0, e0), r in zip(keys, rows):
        motif = r[3]
        k = len(motif)
        assert 1 <= k <= 6 and set(motif) <= set("ACGT")
        assert e0 - s0 >= max(9, 3 * k)
        assert consists_of_perfect_repeats(motif) is None


@pytest.mark.gpu
@pytest.mark.skipif(synth.chr22_path() is None or not os.path.isfile(GOLDEN_BED),
                    reason="benchmark/chr22.fa.gz is not in the reference checkout (.MISSING_LARGE_BLOBS); set CRF_CHR22_FASTA")
def test_config_c1_real_chr22_bed_equals_the_reference_golden(tmp_path, monkeypatch):
    """Config C1 proper: the CLI on the real chr22 FASTA against the BED the reference produced from it."""
    monkeypatch.chdir(tmp_path)
    assert cli.main([synth.chr22_path(), "-min", "1", "-max", "6", "--min-repeats", "3", "--min-span", "9", "-o", "c1"]) == 0
    assert open(tmp_path / "c1.bed").read() == open(GOLDEN_BED).read()
