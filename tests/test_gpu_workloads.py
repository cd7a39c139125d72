"""GPU tests at the BASELINE.json workload shapes: CLI files, partitioned scans with stitching, the
chr22-sized record against the oracle, the hg38-sized genome through size-independent properties
plus sampled windows against the oracle, and the many-reads path."""
import argparse
import gzip
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))

from tests.helpers import ns, random_seq  # noqa: E402

pytestmark = pytest.mark.gpu

DEFAULTS = dict(min_motif_size=1, max_motif_size=50, min_repeats=3, min_span=9)


@pytest.fixture(scope="module")
def mods():
    import torch
    import perfect_repeat_finder as prf
    from crf_b200 import _cabi, api, cli, partition, synth
    from oracle import oracle
    return argparse.Namespace(torch=torch, prf=prf, cabi=_cabi, api=api, cli=cli, partition=partition, synth=synth,
                              oracle=oracle)


def _oracle_rows(oracle, seq, **kw):
    fs = ns(**{**DEFAULTS, **kw})
    return oracle.detect_repeats_by_k(seq, fs)


def test_cli_fasta_to_bed_and_raw_to_tsv(mods, tmp_path, monkeypatch, capsys):
    rng = random.Random(1)
    recs = [("chrA", random_seq(rng, 30000)), ("chrB desc", random_seq(rng, 100)), ("e", ""), ("chrC", random_seq(rng, 70000))]
    fa = tmp_path / "toy.fa.gz"
    with gzip.open(fa, "wt") as f:
        for name, seq in recs:
            f.write(f">{name}\n")
            for i in range(0, len(seq), 60):
                f.write(seq[i:i + 60] + "\n")
    monkeypatch.chdir(tmp_path)
    assert mods.cli.main([str(fa), "-min", "1", "-max", "20"]) == 0
    out = capsys.readouterr().out
    assert "Processing chrA (30,000 bp)" in out and "Wrote results to toy.bed" in out
    want = []
    for name, seq in recs:
        for s, e, m in _oracle_rows(mods.oracle, seq, max_motif_size=20):
            want.append(f"{name.split()[0]}\t{s}\t{e}\t{m}\n")
    assert open(tmp_path / "toy.bed").read() == "".join(want)
    assert len(want) > 50

    # --interval on one record: identical to the reference semantics (oracle with interval attrs)
    assert mods.cli.main([str(fa), "-max", "20", "--interval", "chrC:1000-9000", "-o", str(tmp_path / "iv")]) == 0
    fs = ns(**{**DEFAULTS, "max_motif_size": 20}, interval_start_0based=1000, interval_end=9000)
    want_iv = "".join(f"chrC\t{s}\t{e}\t{m}\n" for s, e, m in mods.oracle.detect_repeats(recs[3][1], fs))
    assert open(tmp_path / "iv.bed").read() == want_iv and want_iv

    # raw sequence -> TSV with header
    assert mods.cli.main(["ACGT" * 5 + "nnnn" + "a" * 12, "-o", str(tmp_path / "raw")]) == 0
    assert open(tmp_path / "raw.tsv").read() == "start_0based\tend\tmotif\n0\t20\tACGT\n24\t36\tA\n"


def test_partitioned_scan_with_stitching_on_gpu(mods):
    """Config C4 in miniature: repeats of 0.5 / 1 / 2.5 chunks around unit boundaries; the chunked,
    halo-limited scan must equal the whole-record scan."""
    chunk, halo = 65536, 4096
    bases, offsets, meta = mods.synth.sx(4_000_000, chunk, device=None)
    ctx = mods.api.get_context()
    whole = mods.api.scan_arrays(bases, 1, 50, 3, 9)
    rec, st, en, k = mods.partition.scan_partitioned(ctx, bases, [0], [bases.size], 1, 50, 3, 9, chunk=chunk, halo=halo)
    assert np.array_equal(st, whole[0]) and np.array_equal(en, whole[1]) and np.array_equal(k, whole[2])
    spans = (whole[1].astype(np.int64) - whole[0])
    assert (spans > chunk + halo).sum() >= 5, "the workload must contain runs that need stitching"
    # and the whole-record scan itself equals the oracle
    o_s, o_e, o_m, _ = mods.oracle.detect_repeats_by_k(bases, ns(**DEFAULTS), arrays=True)
    assert np.array_equal(whole[0], o_s) and np.array_equal(whole[1], o_e) and np.array_equal(whole[2], o_m)


def test_output_map_open_rows_and_patch(mods):
    """The library-side variant of the same bookkeeping: chromosome coordinates via crf_seq_set_output_map,
    open-ended results via crf_fetch_open, ends fixed with crf_run_end + crf_patch_end."""
    chunk, halo = 65536, 4096
    bases, offsets, meta = mods.synth.sx(3_000_000, chunk, device=None)
    ctx = mods.api.get_context()
    whole = mods.api.scan_arrays(bases, 1, 50, 3, 9)
    plan = mods.partition.Plan([bases.size], 1, chunk, halo, 50, 3, 9)
    starts, lens, own_lo, own_hi = plan.load_args(0, [0])
    units = plan.units
    with ctx.load_ranges(bases, starts, lens, own_lo, own_hi, max_motif_cap=50) as seq:
        seq.set_output_map(out_record=[u.record for u in units], out_shift=[u.d0 for u in units],
                           open_ended=[int(u.d1 < u.rec_len) for u in units])
        n = seq.scan(1, 50, 3, 9)
        assert n == len(whole[0])
        open_rows = seq.fetch_open()
        assert len(open_rows) == seq.stats().n_open and len(open_rows) >= 3
        fixed = mods.partition.stitch(plan, [tuple(int(x) for x in r[1:]) for r in open_rows],
                                      lambda unit, lp, k: seq.run_end(unit.index, lp, k))
        for row, (_r, _s, e, _k) in zip(open_rows, fixed):
            seq.patch_end(int(row[0]), e)
        rec, st, en, k = seq.fetch(n)
    assert rec.max() == 0
    assert np.array_equal(st, whole[0]) and np.array_equal(en, whole[1]) and np.array_equal(k, whole[2])


def test_c4_chunk_crossing_repeats_256mbp(mods):
    """Config C4 at full size: one 256 Mbp record with perfect repeats of 0.5 / 1 / 2.5 chunks (2-10 Mbp) whose
    ends sit within +-(k+1) of chunk boundaries, a run to the last base and runs abutting N.  Whole-record scan
    vs the oracle, then the same record cut into 62 units with a 64 kbp halo, every long repeat stitched."""
    chunk, halo = 1 << 22, 1 << 16
    bases, offsets, meta = mods.synth.sx(256_000_000, chunk, device="cuda:0")
    assert len(meta["planted"]) >= 20
    host = bases.cpu().numpy()
    ctx = mods.api.get_context()
    with ctx.load(bases.data_ptr(), offsets, max_motif_cap=50, on_device=True) as seq:
        n = seq.scan(1, 50, 3, 9)
        _, st, en, k = seq.fetch(n)
        assert seq.stats().n_long >= 20
    o_s, o_e, o_m, _ = mods.oracle.detect_repeats_by_k(host, ns(**DEFAULTS), arrays=True)
    assert np.array_equal(st, o_s) and np.array_equal(en, o_e) and np.array_equal(k, o_m)
    spans = en.astype(np.int64) - st
    assert (spans > 2 * chunk).sum() >= 5
    # partitioned: units of 4 Mbp + halo, results in record coordinates, open-ended rows fixed in place
    plan = mods.partition.Plan([host.size], 1, chunk, halo, 50, 3, 9)
    starts, lens, own_lo, own_hi = plan.load_args(0, [0])
    with ctx.load_ranges(bases.data_ptr(), starts, lens, own_lo, own_hi, max_motif_cap=50, on_device=True) as seq:
        seq.set_output_map(out_record=[u.record for u in plan.units], out_shift=[u.d0 for u in plan.units],
                           open_ended=[int(u.d1 < u.rec_len) for u in plan.units])
        n2 = seq.scan(1, 50, 3, 9)
        assert n2 == n
        open_rows = seq.fetch_open()
        assert len(open_rows) >= 20
        fixed = mods.partition.stitch(plan, [tuple(int(x) for x in r[1:]) for r in open_rows],
                                      lambda unit, lp, kk: seq.run_end(unit.index, lp, kk))
        for row, (_r, _s, e, _k) in zip(open_rows, fixed):
            seq.patch_end(int(row[0]), e)
        _, st2, en2, k2 = seq.fetch(n2)
    assert np.array_equal(st2, st) and np.array_equal(en2, en) and np.array_equal(k2, k)


def test_chr22_sized_record_bit_exact(mods):
    """Configs C1 / C2 on the chr22-shaped stand-in (benchmark/chr22.fa.gz is not in the reference checkout)."""
    bases, offsets, meta = mods.synth.s22(device="cuda:0")
    host = bases.cpu().numpy()
    for kw in (dict(min_motif_size=2, max_motif_size=6), dict()):
        fs = {**DEFAULTS, **kw}
        st, en, k = mods.api.scan_arrays(host, fs["min_motif_size"], fs["max_motif_size"], 3, 9)
        o_s, o_e, o_m, _ = mods.oracle.detect_repeats_by_k(host, ns(**fs), arrays=True)
        assert len(o_s) > 20000
        assert np.array_equal(st, o_s) and np.array_equal(en, o_e) and np.array_equal(k, o_m)


def test_hg38_sized_genome_every_record_row_by_row(mods):
    """Config C3 at full size: ordering / threshold properties on every row, then ALL 24 records (3.09 Gbp, every one of
    the ~6.4 M rows) compared row by row with the oracle (one tracker thread per motif size, ~1 min of host time)."""
    bases, offsets, meta = mods.synth.s38(device="cuda:0")
    ctx = mods.api.get_context()
    with ctx.load(bases.data_ptr(), offsets, max_motif_cap=50, on_device=True) as seq:
        n = seq.scan(1, 50, 3, 9)
        rec, st, en, k = seq.fetch(n)
    assert n > 5_000_000
    rec64, st64, en64, k64 = (a.astype(np.int64) for a in (rec, st, en, k))
    key = (rec64 << 40) | st64
    assert np.all(np.diff(key) >= 0)                                      # sorted by (record, start)
    same = np.diff(key) == 0
    assert np.all(np.diff(en64)[same] > 0)                                 # ... then by end, keys unique
    assert np.all(en64 - st64 >= np.maximum(9, 3 * k64))                   # span >= max(min_span, min_repeats*k)
    assert k64.min() >= 1 and k64.max() <= 50
    lengths = np.diff(offsets.astype(np.int64))
    assert np.all(en64 <= lengths[rec64])
    bounds = np.searchsorted(rec64, np.arange(25))
    checked = 0
    for r in range(24):
        host = bases[int(offsets[r]):int(offsets[r + 1])].cpu().numpy()
        o_s, o_e, o_m, _ = mods.oracle.detect_repeats_by_k(host, ns(**DEFAULTS), arrays=True)
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        assert hi - lo == len(o_s), f"record {r}: {hi - lo} rows vs {len(o_s)} from the oracle"
        assert np.array_equal(st64[lo:hi], o_s) and np.array_equal(en64[lo:hi], o_e) and np.array_equal(k64[lo:hi], o_m), f"record {r}"
        checked += hi - lo
    assert checked == n
    # the hash bench.py compares its gathered rows with at every N (tests/golden/workload_rows_sha256.json): these rows have
    # just been checked against the oracle one by one
    import hashlib
    import json
    sha = hashlib.sha256()
    for a in (rec, st, en, k):
        sha.update(np.ascontiguousarray(a, dtype=np.uint32).tobytes())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "workload_rows_sha256.json"), "w") as f:
        json.dump({"s38:1.0:1-50": sha.hexdigest(), "rows": int(n)}, f)
    golden = os.path.join(ROOT, "tests", "golden", "workload_rows_sha256.json")
    if os.path.exists(golden):
        known = json.load(open(golden)).get("s38:1.0:1-50")
        assert known in (None, sha.hexdigest())


def test_all_ten_million_reads_row_by_row(mods):
    """Config C5 at full size: 10 M reads x 150 bp, motif 1-20, every read compared with the oracle.  The oracle scans the
    reads as one string with max_motif_size N's between them (N never matches, trk:53, so no run can leave its read)."""
    torch = mods.torch
    n_reads, rl, gap = 10_000_000, 150, 20
    bases, offsets, meta = mods.synth.sr(n_reads, device="cuda:0")
    ctx = mods.api.get_context()
    with ctx.load(bases.data_ptr(), offsets, max_motif_cap=20, on_device=True) as seq:
        n = seq.scan(1, 20, 3, 9)
        rec, st, en, k = seq.fetch(n)
    assert n > 500_000 and en.max() <= rl
    gapped = torch.full((n_reads, rl + gap), ord("N"), dtype=torch.uint8, device="cuda:0")
    gapped[:, :rl] = bases.view(n_reads, rl)
    host = gapped.flatten().cpu().numpy()
    del gapped
    o_s, o_e, o_m, _ = mods.oracle.detect_repeats_by_k(host, ns(**{**DEFAULTS, "max_motif_size": 20}), arrays=True)
    assert len(o_s) == n
    o_rec = o_s // (rl + gap)
    assert np.array_equal(rec.astype(np.int64), o_rec)
    assert np.array_equal(st.astype(np.int64), o_s - o_rec * (rl + gap))
    assert np.array_equal(en.astype(np.int64), o_e - o_rec * (rl + gap))
    assert np.array_equal(k.astype(np.int64), o_m.astype(np.int64))


def test_many_reads_path(mods):
    """Config C5 in miniature: 200 000 independent 150-bp reads, motif 1-20; coordinates are per read and no
    repeat may span two reads."""
    n_reads = 200_000
    bases, offsets, meta = mods.synth.sr(n_reads, device="cuda:0")
    ctx = mods.api.get_context()
    with ctx.load(bases.data_ptr(), offsets, max_motif_cap=20, on_device=True) as seq:
        n = seq.scan(1, 20, 3, 9)
        rec, st, en, k = seq.fetch(n)
    assert n > 10_000 and en.max() <= 150
    host = bases.cpu().numpy().reshape(n_reads, 150)
    fs = ns(**{**DEFAULTS, "max_motif_size": 20})
    rng = np.random.default_rng(9)
    for r in rng.choice(n_reads, 3000, replace=False).tolist():
        want = mods.oracle.detect_repeats_by_k(np.ascontiguousarray(host[r]), fs, arrays=True)
        sel = rec == r
        assert np.array_equal(st[sel], want[0]) and np.array_equal(en[sel], want[1]) and np.array_equal(k[sel], want[2])
    # every read with a result: cross-check a few hundred against the oracle too
    for r in np.unique(rec)[:300].tolist():
        want = mods.oracle.detect_repeats_by_k(np.ascontiguousarray(host[r]), fs, arrays=True)
        sel = rec == r
        assert np.array_equal(st[sel], want[0]) and np.array_equal(en[sel], want[1])


def test_millions_of_records_table_paths(mods):
    """The per-record tables of a load with more than 2^20 records are built on several threads in the context's page-locked
    arena, the host copies crf_run_end / crf_seq_set_output_map need are read back on first use, boundaries are taken as
    they are and records that do not end in ascending source order fall back to the running-maximum table: 1.2 M reads
    of 150 bp from the host, as boundaries, as (start, length) ranges in order and as ranges in REVERSE order (pipelined
    upload, ends unsorted) -- all against the load of the same bytes already in HBM, plus run ends and shifted output."""
    n_reads = 1_200_000
    bases, offsets, _ = mods.synth.sr(n_reads, device="cuda:0")
    host = bases.cpu().numpy()
    ctx = mods.api.get_context()

    def rows(seq):
        n = seq.scan(1, 20, 3, 9)
        return seq.fetch(n)

    with ctx.load(bases.data_ptr(), offsets, max_motif_cap=20, on_device=True) as seq:
        want = rows(seq)
    assert want[0].size > 100_000
    starts, lens = offsets[:-1].copy(), np.diff(offsets)
    with ctx.load(host, offsets, max_motif_cap=20) as seq:                                   # boundaries, host text
        got = rows(seq)
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
        r = int(want[0][1000])                                                               # lazy host tables: a run end ...
        st, k = int(want[1][1000]), int(want[3][1000])
        assert seq.run_end(r, st, k) == int(want[2][1000]) - k                                # where the run of matches stops: end - k
        with pytest.raises(ValueError):
            seq.run_end(r, 150, k)
        seq.set_output_map(out_shift=np.full(n_reads, 7, np.uint64))                         # ... and shifted coordinates
        shifted = rows(seq)
        assert np.array_equal(shifted[1], want[1] + 7) and np.array_equal(shifted[2], want[2] + 7)
    with ctx.load_ranges(host, starts, lens, max_motif_cap=20) as seq:                       # ranges in source order
        assert all(np.array_equal(a, b) for a, b in zip(rows(seq), want))
    pk = mods.cabi.pack_ascii(host).with_runs()
    with ctx.load_packed(pk, offsets, max_motif_cap=20) as seq:                              # planes + mask runs, boundaries
        assert all(np.array_equal(a, b) for a, b in zip(rows(seq), want))
    with ctx.load_ranges(host, starts[::-1].copy(), lens[::-1].copy(), max_motif_cap=20) as seq:   # reverse order
        rec, st, en, k = rows(seq)
    order = np.lexsort((en, st, n_reads - 1 - rec.astype(np.int64)))
    assert np.array_equal(n_reads - 1 - rec[order].astype(np.int64), want[0].astype(np.int64))
    assert np.array_equal(st[order], want[1]) and np.array_equal(en[order], want[2]) and np.array_equal(k[order], want[3])


def test_pipelined_host_upload_equals_device_load(mods):
    """Host buffers above 128 MB go up in 64 MB chunks with the pack kernel running behind the copy; the packed
    planes must not depend on how the bytes arrived: same rows as a load of the same bytes already in HBM, for
    records in source order, for overlapping out-of-order ranges with empty records in between, and for reads."""
    torch = mods.torch
    bases, offsets, _ = mods.synth.s38(device="cuda:0", scale=0.1)          # ~309 Mbp, 24 records
    host = bases.cpu().numpy()
    ctx = mods.api.get_context()

    def rows(seq, kmax=50):
        n = seq.scan(1, kmax, 3, 9)
        out = seq.fetch(n)
        seq.close()
        return out

    want = rows(ctx.load(bases.data_ptr(), offsets, max_motif_cap=50, on_device=True))
    got = rows(ctx.load(host, offsets, max_motif_cap=50))
    assert len(want[0]) > 500000
    assert all(np.array_equal(a, b) for a, b in zip(want, got))

    rng = np.random.default_rng(3)
    n = len(host)
    starts = rng.integers(0, n - 40_000_000, size=12).astype(np.uint64)     # out of order, overlapping
    lens = rng.integers(1, 40_000_000, size=12).astype(np.uint64)
    lens[[2, 7]] = 0
    want = rows(ctx.load_ranges(bases.data_ptr(), starts, lens, max_motif_cap=50, on_device=True))
    got = rows(ctx.load_ranges(host, starts, lens, max_motif_cap=50))
    assert len(want[0]) > 100000
    assert all(np.array_equal(a, b) for a, b in zip(want, got))

    n_reads = 2_000_000                                                     # 300 MB of 150-base records
    r_off = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(150))
    want = rows(ctx.load(bases.data_ptr(), r_off, max_motif_cap=20, on_device=True), 20)
    got = rows(ctx.load(host[:n_reads * 150], r_off, max_motif_cap=20), 20)
    assert len(want[0]) > 10000 and all(np.array_equal(a, b) for a, b in zip(want, got))
    del torch


def _hail_fanout(cli_main, oracle, tmp_path, recs, batch, kmax):
    """The reference's scale-out (hail_batch_pipeline/run_hail_batch_pipeline.py:76-77, 115-123, 148-153): one
    `--interval chrom:start-end` run per batch, the BED files concatenated, then `sort -k1,1 -k2,2n | uniq`.
    Returns (lines from the CLI runs, lines the reference would produce: the oracle per interval)."""
    fa = tmp_path / "fan.fa"
    with open(fa, "wt") as f:
        for name, seq in recs:
            f.write(f">{name}\n")
            for i in range(0, len(seq), 70):
                f.write(seq[i:i + 70] + "\n")
    got, want = [], []
    for name, seq in recs:
        for b0 in range(0, len(seq), batch):
            b1 = min(len(seq), b0 + batch)
            prefix = tmp_path / f"fan.{name}_{b0}"
            assert cli_main([str(fa), "-max", str(kmax), "--interval", f"{name}:{b0}-{b1}", "-o", str(prefix)]) == 0
            got += open(tmp_path / f"{prefix.name}.bed").read().splitlines()
            fs = ns(**{**DEFAULTS, "max_motif_size": kmax}, interval_start_0based=b0, interval_end=b1)
            want += [f"{name}\t{s}\t{e}\t{m}" for s, e, m in oracle.detect_repeats(seq, fs)]

    def sort_uniq(lines):
        return sorted(set(lines), key=lambda ln: (ln.split("\t")[0], int(ln.split("\t")[1]), ln))
    return sort_uniq(got), sort_uniq(want), len(got)


def _fanout_records():
    rng = random.Random(9)
    a = list(random_seq(rng, 45000))
    for pos, unit, copies in [(9990, "CAG", 12), (19950, "AT", 400), (29999, "A", 30), (39990, "GATTACA", 5)]:
        a[pos:pos + len(unit) * copies] = unit * copies        # repeats across the 10 kb batch boundaries
    b = random_seq(rng, 10000)[:4000] + "N" * 2100 + random_seq(rng, 10000)[:12000]   # an N block over a boundary
    return [("chr1", "".join(a)[:45000]), ("chr2", b)]


def test_hail_style_interval_fanout_then_sort_uniq(mods, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    got, want, n_raw = _hail_fanout(mods.cli.main, mods.oracle, tmp_path, _fanout_records(), batch=10000, kmax=30)
    assert got == want and len(got) > 500 and n_raw >= len(got)
    # a repeat straddling a batch boundary comes out whole from the batch it starts in (the loop runs on while a tracker is
    # mid-repeat, prf:70-74) and left-clipped from the next one: chr1 9990-10026 (CAG)x12 and its twin 10000-10026
    assert "chr1\t9990\t10026\tCAG" in got and "chr1\t10000\t10026\tAGC" in got


def test_packed_load_equals_ascii_load_equals_oracle(mods):
    """crf_seq_load_packed (planes packed on the host, 0.375 B/bp over PCIe, re-laid out by repack_kernel) must give the
    rows of crf_seq_load_ascii on the same text, which are the oracle's: ragged multi-record input with lower case, N
    runs and IUPAC letters (equal exotic letters match, trk:53), reads, overlapping out-of-order ranges with owned
    sub-ranges, and a text large enough for the chunked, pipelined upload."""
    cabi, oracle = mods.cabi, mods.oracle
    ctx = mods.api.get_context()
    rng = random.Random(12)

    def rows(seq, kmax=50):
        n = seq.scan(1, kmax, 3, 9)
        out = seq.fetch(n)
        seq.close()
        return out

    # 1. ragged records with exotic letters, against the oracle per record
    recs = [random_seq(rng, n, exotic=True) for n in (1, 31, 32, 33, 5000, 0, 64, 70_001, 12_345)]
    recs.append("R" * 40 + "ACRACRACRACRACR" + "Y" * 9)          # exotic letters that repeat (they match each other)
    text = np.frombuffer("".join(recs).encode(), dtype=np.uint8)
    offsets = np.concatenate([[0], np.cumsum([len(r) for r in recs])]).astype(np.uint64)
    pk = cabi.pack_ascii(text)
    assert pk.exotic.size > 50
    a = rows(ctx.load(text, offsets, max_motif_cap=50))
    b = rows(ctx.load_packed(pk, offsets, max_motif_cap=50))
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and len(a[0]) > 300
    c = rows(ctx.load_packed(pk.with_runs(), offsets, max_motif_cap=50))      # the mask as runs instead of a plane
    assert pk.runs.shape[0] > 20 and all(np.array_equal(x, y) for x, y in zip(a, c))
    want = []
    for r, rec in enumerate(recs):
        want += [(r, s, e, len(m)) for s, e, m in oracle.detect_repeats_by_k(rec, ns(**DEFAULTS))]
    assert list(zip(*(x.tolist() for x in b))) == want

    # 2. reads (many short records: layout words that span several records and gaps)
    n_reads = 50_000
    bases, r_off, _ = mods.synth.sr(n_reads, device=None)
    pk = cabi.pack_ascii(bases)
    a = rows(ctx.load(bases, r_off, max_motif_cap=20), 20)
    b = rows(ctx.load_packed(pk, r_off, max_motif_cap=20), 20)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and len(a[0]) > 2000
    c = rows(ctx.load_packed(pk.with_runs(), r_off, max_motif_cap=20), 20)    # thousands of single-position runs
    assert pk.runs.shape[0] > 5000 and all(np.array_equal(x, y) for x, y in zip(a, c))

    # 3. the pipelined upload (> 1 Gbp of planes' positions), records in order; then overlapping out-of-order ranges
    #    with owned sub-ranges and empty records
    big, offs, _ = mods.synth.s38(device="cuda:0", scale=0.4)                 # ~1.24 Gbp, 24 records
    host = big.cpu().numpy()
    pk = cabi.pack_ascii(host)
    a = rows(ctx.load(big.data_ptr(), offs, max_motif_cap=50, on_device=True))
    b = rows(ctx.load_packed(pk, offs, max_motif_cap=50))
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and len(a[0]) > 2_000_000
    nrng = np.random.default_rng(3)
    n = len(host)
    starts = nrng.integers(0, n - 90_000_000, size=14).astype(np.uint64)
    lens = nrng.integers(1, 90_000_000, size=14).astype(np.uint64)
    lens[[2, 7]] = 0
    own_lo = (lens // np.uint64(7)).astype(np.uint64)
    own_hi = (lens - lens // np.uint64(5)).astype(np.uint64)
    a = rows(ctx.load_ranges(big.data_ptr(), starts, lens, own_lo, own_hi, max_motif_cap=50, on_device=True))
    b = rows(ctx.load_packed(pk, max_motif_cap=50, ranges=(starts, lens, own_lo, own_hi)))
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and len(a[0]) > 500_000
    pk.with_runs()                                                             # runs mode: pipelined, then ranges
    c = rows(ctx.load_packed(pk, max_motif_cap=50, ranges=(starts, lens, own_lo, own_hi)))
    assert all(np.array_equal(x, y) for x, y in zip(a, c))
    d = rows(ctx.load_packed(pk, offs, max_motif_cap=50))
    want_all = rows(ctx.load(big.data_ptr(), offs, max_motif_cap=50, on_device=True))
    assert all(np.array_equal(x, y) for x, y in zip(want_all, d))
