"""Multi-GPU host logic on CPU: the partition plan, ownership, and the stitching of runs that leave a
unit -- with a fake context that answers scans from the CPU oracle, single process and 2 ranks over
gloo (torch.distributed)."""
import argparse
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))

from crf_b200 import partition  # noqa: E402
from oracle import oracle  # noqa: E402
from tests.fake_ctx import FakeContext  # noqa: E402
from tests.helpers import random_seq  # noqa: E402

KMIN, KMAX, MR, MS = 1, 12, 3, 9


def make_records(seed):
    rng = random.Random(seed)
    recs = []
    for n in (5000, 1, 0, 2300, 777):
        recs.append(random_seq(rng, n))
    # runs far longer than chunk + halo, crossing several unit boundaries, and one reaching the record end
    long1 = "ACG" * 900
    long2 = "T" * 1500
    recs[0] = recs[0][:700] + long1 + recs[0][700 + len(long1):]
    recs[3] = recs[3][:300] + long2
    return [r.encode() for r in recs]


def expected(records):
    fs = argparse.Namespace(min_motif_size=KMIN, max_motif_size=KMAX, min_repeats=MR, min_span=MS)
    rows = []
    for r, rec in enumerate(records):
        for s, e, m in oracle.detect_repeats_by_k(rec, fs):
            rows.append((r, s, e, len(m)))
    return rows


def test_plan_covers_every_base_once_and_balances():
    lengths = [5000, 1, 0, 2300, 777]
    for world in (1, 2, 3, 8):
        plan = partition.Plan(lengths, world, chunk=256, halo=64, kmax=KMAX, min_repeats=MR, min_span=MS)
        owned = [0] * len(lengths)
        seen = []
        for rank in range(world):
            for u in plan.units_of(rank):
                owned[u.record] += u.u1 - u.u0
                assert u.d0 == max(0, u.u0 - 1) and u.d1 == min(u.rec_len, u.u1 + 64)
                assert plan.rank_of_unit(u.index) == rank
                assert plan.unit_owning(u.record, u.u0).index == u.index
                seen.append(u.index)
        assert owned == lengths and seen == list(range(len(plan.units)))
        per_rank = [sum(u.u1 - u.u0 for u in plan.units_of(r)) for r in range(world)]
        assert max(per_rank) - min(per_rank) <= 2 * 256
    with pytest.raises(ValueError):
        partition.Plan(lengths, 2, chunk=256, halo=16, kmax=KMAX, min_repeats=MR, min_span=MS)


def test_plan_balances_by_cost_density_with_cuts_inside_chunks():
    lengths = [5000, 1, 0, 2300, 777, 40_000]
    total = sum(lengths)
    # the second half of the genome costs three times as much per base
    density = [(0, total // 2, 1.0), (total // 2, total, 3.0)]
    for world in (2, 3, 8):
        plan = partition.Plan(lengths, world, chunk=4096, halo=64, kmax=KMAX, min_repeats=MR, min_span=MS, density=density)
        seen, owned = [], [0] * len(lengths)
        costs = []
        for rank in range(world):
            lo, hi = plan.share_range(rank)
            costs.append(sum((min(hi, h) - max(lo, l)) * c for l, h, c in density if min(hi, h) > max(lo, l)))
            for u in plan.units_of(rank):
                owned[u.record] += u.u1 - u.u0
                assert plan.rank_of_unit(u.index) == rank and plan.unit_owning(u.record, u.u0).index == u.index
                assert plan.unit_owning(u.record, u.u1 - 1).index == u.index
                seen.append(u.index)
        assert owned == lengths and seen == list(range(len(plan.units)))
        assert max(costs) - min(costs) <= 0.05 * sum(costs) / world + 3 * 600     # equal cost, up to the minimum split distance
        assert max(u.u1 - u.u0 for u in plan.units) <= 4096


@pytest.mark.parametrize("chunk,halo", [(256, 64), (500, 100), (4096, 64), (1 << 20, 64)])
def test_single_rank_chunked_scan_equals_whole_record_scan(chunk, halo):
    records = make_records(3)
    lengths = [len(r) for r in records]
    starts = np.concatenate([[0], np.cumsum(lengths)])[:-1]
    rec, st, en, k = partition.scan_partitioned(FakeContext(), b"".join(records), starts, lengths, KMIN, KMAX, MR, MS,
                                                chunk=chunk, halo=halo)
    got = list(zip(rec.tolist(), st.tolist(), en.tolist(), k.tolist()))
    assert got == expected(records)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    records = make_records(3)
    lengths = [len(r) for r in records]
    starts = np.concatenate([[0], np.cumsum(lengths)])[:-1]
    rec, st, en, k = partition.scan_partitioned(FakeContext(), b"".join(records), starts, lengths, KMIN, KMAX, MR, MS,
                                                rank=rank, world=world, chunk=300, halo=64, dist=dist)
    first = list(zip(rec.tolist(), st.tolist(), en.tolist(), k.tolist()))
    # the same with the fixed-size tensor collectives bench.py uses for the stitch (CPU tensors over gloo)
    rec, st, en, k = partition.scan_partitioned(FakeContext(), b"".join(records), starts, lengths, KMIN, KMAX, MR, MS,
                                                rank=rank, world=world, chunk=300, halo=64, dist=dist,
                                                tensor_device="cpu")
    assert list(zip(rec.tolist(), st.tolist(), en.tolist(), k.tolist())) == first
    q.put((rank, first))
    dist.barrier()
    dist.destroy_process_group()


# ---- crf_b200.multi.RankScan (the path bench.py and the CLI's --devices use) with the library replaced by CPU stand-ins:
# ---- shares, phases, the offsets inside rank 0's buffer, the stitch of open-ended rows and their patching ----
def _rankscan_rows(multi, comm, records, chunk, halo, phases, rebalance=False):
    lengths = [len(r) for r in records]
    starts = np.concatenate([[0], np.cumsum(lengths)])[:-1]
    rs = multi.RankScan(FakeContext(), comm, np.frombuffer(b"".join(records), dtype=np.uint8), starts, lengths,
                        KMIN, KMAX, MR, MS, chunk=chunk, halo=halo, phases=phases)
    rs.step_async()
    n = rs.finish()
    if rebalance:
        rs.rebalance(min_gain=-1.0)
        rs.step_async()
        n = rs.finish()
    rows = None
    if comm.rank == 0:
        rec, st, en, k = rs.fetch()
        rows = list(zip(rec.tolist(), st.tolist(), en.tolist(), k.tolist()))
        assert len(rows) == n
    rs.close()
    return rows


@pytest.mark.parametrize("world,phases,chunk,halo", [(1, 1, 300, 64), (2, 1, 300, 64), (3, 2, 256, 64), (4, 3, 500, 100)])
def test_rankscan_with_thread_ranks(world, phases, chunk, halo, monkeypatch):
    import threading
    from crf_b200 import multi
    from tests.fake_ctx import FakeXchg
    monkeypatch.setattr(multi._cabi, "Xchg", FakeXchg)
    monkeypatch.setattr(FakeXchg, "dist", None)
    records = make_records(3)
    comms = multi.ThreadComm.split(world)
    out, errors = {}, []

    def work(rank):
        try:
            out[rank] = _rankscan_rows(multi, comms[rank], records, chunk, halo, phases, rebalance=(phases == 1 and world > 1))
        except BaseException as exc:      # noqa: BLE001
            errors.append(exc)
            comms[rank]._s.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert out[0] == expected(records)


def _rankscan_worker(rank, world, port, q):
    import torch.distributed as dist
    from crf_b200 import multi
    from tests.fake_ctx import FakeXchg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    multi._cabi.Xchg = FakeXchg
    FakeXchg.dist = dist
    rows = _rankscan_rows(multi, multi.DistComm(dist, rank, world), make_records(3), 300, 64, 2)
    q.put((rank, rows))
    dist.barrier()
    dist.destroy_process_group()


def test_rankscan_two_process_ranks_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randint(2001, 4000)
    procs = [ctx.Process(target=_rankscan_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0] == expected(make_records(3)) and results[1] is None


def test_two_ranks_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randint(0, 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = expected(make_records(3))
    assert results[0] == want and results[1] == want
