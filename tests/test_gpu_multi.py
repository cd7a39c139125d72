"""The multi-GPU path (crf_b200/multi.py + csrc/crf_xchg.cuh) on the GPU box.  The driver's test box has ONE GPU, so
most of these run N ranks as N threads that all use device 0 ("loopback"): the exchange blocks, the count / done
protocol, the peer stores, the void-step fallback and the stitch of open-ended rows are exactly the code N GPUs run --
only the stores stay inside one HBM.  The tests that need distinct GPUs are skipped below 2 devices."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))

from tests.helpers import ns  # noqa: E402

pytestmark = pytest.mark.gpu

DEFAULTS = dict(min_motif_size=1, max_motif_size=50, min_repeats=3, min_span=9)


@pytest.fixture(scope="module")
def mods():
    import argparse
    import torch
    from crf_b200 import _cabi, api, cli, multi, partition, synth
    from oracle import oracle
    return argparse.Namespace(torch=torch, cabi=_cabi, api=api, cli=cli, multi=multi, partition=partition, synth=synth,
                              oracle=oracle)


def _whole(mods, bases, offsets, kmax=50):
    ctx = mods.api.get_context()
    with ctx.load(bases, offsets, max_motif_cap=kmax) as seq:
        n = seq.scan(1, kmax, 3, 9)
        return seq.fetch(n)


@pytest.mark.parametrize("world,phases", [(2, 1), (2, 2), (3, 3), (8, 2)])
def test_loopback_ranks_gather_equals_single_scan(mods, world, phases):
    """24 ragged records over `world` ranks (threads on device 0), each step in `phases` phases whose pushes overlap the
    next phase's scan: rank 0's gathered rows == the one-load scan."""
    bases, offsets, _ = mods.synth.s38(device=None, scale=0.002)            # ~6 Mbp, 24 records
    want = _whole(mods, bases, offsets)
    lengths = np.diff(offsets.astype(np.int64))
    got = mods.multi.scan_on_devices([0] * world, bases, offsets[:-1], lengths, 1, 50, 3, 9, chunk=1 << 17, halo=1 << 12,
                                     timeout_s=30, phases=phases)
    assert len(want[0]) > 5000
    for a, b in zip(want, got):
        assert np.array_equal(a, b)


def test_loopback_stitches_runs_longer_than_the_halo(mods):
    """Config C4 in miniature over 4 ranks: repeats of 0.5 / 1 / 2.5 chunks cross unit AND rank boundaries, so rows
    arrive open-ended and are patched inside rank 0's buffer (crf_xchg_patch_end)."""
    chunk, halo = 65536, 4096
    bases, offsets, meta = mods.synth.sx(4_000_000, chunk, device=None)
    want = _whole(mods, bases, offsets)
    got = mods.multi.scan_on_devices([0] * 4, bases, [0], [bases.size], 1, 50, 3, 9, chunk=chunk, halo=halo, timeout_s=30)
    assert (want[1].astype(np.int64) - want[0] > chunk + halo).sum() >= 5
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    o_s, o_e, o_m, _ = mods.oracle.detect_repeats_by_k(bases, ns(**DEFAULTS), arrays=True)
    assert np.array_equal(got[1], o_s) and np.array_equal(got[2], o_e) and np.array_equal(got[3], o_m)


def test_loopback_void_step_falls_back_and_recovers(mods):
    """A result buffer that is too small makes the asynchronous step void on EVERY rank; all ranks repeat it through
    crf_scan (which grows the buffer) + crf_xchg_push."""
    bases, offsets, _ = mods.synth.s38(device=None, scale=0.002)
    want = _whole(mods, bases, offsets)
    lengths = np.diff(offsets.astype(np.int64))
    got = mods.multi.scan_on_devices([0, 0], bases, offsets[:-1], lengths, 1, 50, 3, 9, chunk=1 << 18, halo=1 << 12,
                                     knobs={"result_cap": 64}, timeout_s=30)
    for a, b in zip(want, got):
        assert np.array_equal(a, b)


def test_loopback_reads_split_by_count(mods):
    n_reads = 60_000
    bases, offsets, _ = mods.synth.sr(n_reads, device=None)
    want = _whole(mods, bases, offsets, kmax=20)
    got = mods.multi.scan_on_devices([0, 0, 0], bases, offsets[:-1], np.full(n_reads, 150), 1, 20, 3, 9, reads=True,
                                     timeout_s=30)
    assert len(want[0]) > 3000
    for a, b in zip(want, got):
        assert np.array_equal(a, b)


def test_many_steps_in_flight_then_one_wait(mods):
    """bench.py's pattern: K asynchronous steps back to back, one wait; every step's status is checked."""
    bases, offsets, _ = mods.synth.s38(device=None, scale=0.002)
    want = _whole(mods, bases, offsets)
    lengths = np.diff(offsets.astype(np.int64))
    world = 2
    comms = mods.multi.ThreadComm.split(world)
    import threading
    out = {}

    def work(rank):
        rs = mods.multi.RankScan(mods.cabi.Context(0), comms[rank], bases, offsets[:-1], lengths, 1, 50, 3, 9,
                                 chunk=1 << 18, halo=1 << 12, timeout_s=30, phases=2)
        rs.step_async()
        rs.finish()                                   # sizes the buffers
        for _ in range(20):
            rs.step_async()
        n = rs.finish()
        out[rank] = (n, rs.last.steps_checked, rs.last.worst_status, rs.fetch() if rank == 0 else None)
        rs.close()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert out[0][0] == out[1][0] == len(want[0])
    assert out[0][1] == 40 and out[0][2] == 0          # 20 jobs x 2 phases, every status word checked
    for a, b in zip(want, out[0][3]):
        assert np.array_equal(a, b)


def test_rebalanced_shares_give_the_same_rows(mods):
    """RankScan.rebalance(): shares re-cut by measured scan time (boundaries inside chunks), sequences reloaded; the rows
    gathered on rank 0 stay the single-scan rows."""
    import threading
    bases, offsets, _ = mods.synth.s38(device=None, scale=0.004)
    bases = bases.copy()
    bases[: len(bases) // 3] = ord("N")                    # one third costs nothing: the first ranks' shares are cheap
    want = _whole(mods, bases, offsets)
    lengths = np.diff(offsets.astype(np.int64))
    world = 3
    comms = mods.multi.ThreadComm.split(world)
    out = {}

    def work(rank):
        rs = mods.multi.RankScan(mods.cabi.Context(0), comms[rank], bases, offsets[:-1], lengths, 1, 50, 3, 9,
                                 chunk=1 << 19, halo=1 << 12, timeout_s=30)
        rs.step_async()
        rs.finish()
        before = rs.plan.share_range(rank)
        gains = []
        for _ in range(2):
            gains.append(rs.rebalance(min_gain=0.0))
            rs.step_async()
            rs.finish()
        out[rank] = (before, rs.plan.share_range(rank), gains, rs.fetch() if rank == 0 else None)
        rs.close()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert len(out) == world
    for a, b in zip(want, out[0][3]):
        assert np.array_equal(a, b)
    assert any(out[r][0] != out[r][1] for r in range(world))       # the boundaries did move
    spans = sorted(out[r][1] for r in range(world))
    assert spans[0][0] == 0 and spans[-1][1] == int(lengths.sum()) and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_cli_devices_writes_the_same_bed(mods, tmp_path, monkeypatch):
    """--devices: the BED of a multi-rank run equals the single-GPU BED byte for byte (loopback: both ranks on GPU 0;
    with 2+ GPUs also on distinct devices)."""
    bases, offsets, meta = mods.synth.s38(device=None, scale=0.001)
    fa = tmp_path / "g.fa"
    with open(fa, "wt") as f:
        for r in range(len(offsets) - 1):
            seq = bytes(bases[int(offsets[r]):int(offsets[r + 1])]).decode()
            f.write(f">chr{r + 1} test\n")
            for i in range(0, len(seq), 80):
                f.write(seq[i:i + 80] + "\n")
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(mods.partition, "DEFAULT_CHUNK", 1 << 16)
    monkeypatch.setattr(mods.partition, "DEFAULT_HALO", 1 << 12)
    assert mods.cli.main([str(fa), "-o", "one"]) == 0
    assert mods.cli.main([str(fa), "-o", "two", "--devices", "0,0"]) == 0
    one = open(tmp_path / "one.bed", "rb").read()
    assert len(one) > 50_000 and open(tmp_path / "two.bed", "rb").read() == one
    if mods.torch.cuda.device_count() >= 2:
        assert mods.cli.main([str(fa), "-o", "three", "--devices", "0-1"]) == 0
        assert open(tmp_path / "three.bed", "rb").read() == one


@pytest.mark.skipif("__import__('torch').cuda.device_count() < 2")
def test_two_real_gpus_gather_over_nvlink(mods):
    bases, offsets, _ = mods.synth.s38(device=None, scale=0.01)
    want = _whole(mods, bases, offsets)
    lengths = np.diff(offsets.astype(np.int64))
    n_dev = min(mods.torch.cuda.device_count(), 8)
    got = mods.multi.scan_on_devices(list(range(n_dev)), bases, offsets[:-1], lengths, 1, 50, 3, 9, chunk=1 << 20,
                                     halo=1 << 14, timeout_s=30)
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
