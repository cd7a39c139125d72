#!/usr/bin/env python3
"""Differential fuzz campaign: CUDA path vs the CPU oracle on random sequences, filters and kernel knobs.
Not collected by pytest (long-running); run on a GPU box:  python tests/fuzz_gpu.py --seconds 300 --seed 1"""
import argparse
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))

from oracle import oracle  # noqa: E402
import perfect_repeat_finder as prf  # noqa: E402
from tests.helpers import ns, random_seq  # noqa: E402


def packed_vs_ascii(seq, kmin, kmax, min_repeats, min_span, rng):
    """The same text split into random records, loaded as ASCII, as host-packed planes and as planes + mask runs
    (crf_seq_load_ascii / crf_seq_load_packed / crf_seq_load_packed_runs): the raw rows of one scan must be identical."""
    import numpy as np
    from crf_b200 import _cabi, api
    raw = np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)
    cuts = sorted({0, len(seq), *(rng.randint(0, len(seq)) for _ in range(rng.choice([0, 1, 3, 9])))})
    offsets = np.array(cuts, dtype=np.uint64)
    ctx = api.get_context()

    def rows(s):
        with s:
            n = s.scan(kmin, kmax, min_repeats, min_span)
            return [a.tolist() for a in s.fetch(n)]

    a = rows(ctx.load(raw, offsets, max_motif_cap=kmax))
    pk = _cabi.pack_ascii(raw)
    b = rows(ctx.load_packed(pk, offsets, max_motif_cap=kmax))
    c = rows(ctx.load_packed(pk.with_runs(), offsets, max_motif_cap=kmax))
    if a != b or a != c:
        path = f"/tmp/fuzz_fail_packed_{len(seq)}.txt"
        with open(path, "w") as f:
            f.write(repr((seq, cuts, kmin, kmax, min_repeats, min_span)))
        print("MISMATCH packed vs ascii", len(seq), cuts, kmin, kmax, min_repeats, min_span, "saved", path)
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = random.Random(args.seed)
    t_end = time.time() + args.seconds
    n_cases = n_rows = n_packed = 0
    while time.time() < t_end:
        n = rng.choice([rng.randint(0, 300), rng.randint(300, 5000), rng.randint(5000, 120000), rng.randint(60000, 400000)])
        seq = random_seq(rng, n, exotic=rng.random() < 0.3)
        if rng.random() < 0.3:                       # long exact repeats crossing strips / tiles
            unit = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 60)))
            pos = rng.randint(0, max(0, len(seq) - 1))
            seq = seq[:pos] + unit * rng.randint(3, 3000) + seq[pos:]
        kmin = rng.choice([1, 1, 1, 2, 3, 7, 20])
        kmax = kmin + rng.choice([0, 1, 5, 15, 30, 49, 63, 64, 100, 200])
        fs = dict(min_motif_size=kmin, max_motif_size=kmax, min_repeats=rng.choice([1, 2, 2, 3, 3, 3, 4, 6]),
                  min_span=rng.choice([1, 2, 5, 9, 9, 9, 12, 33, 100]))
        if fs["min_repeats"] == 1 and kmin == 1 and fs["min_span"] == 1:
            fs["min_span"] = 2                       # the one refused setting (every base is a "repeat")
        interval = rng.random() < 0.15 and len(seq) > 10
        if interval:
            a = rng.randint(0, len(seq))
            fs.update(interval_start_0based=a, interval_end=rng.randint(a, len(seq)))
        knobs = rng.choice([{}, {}, {"words_per_thread": 1}, {"words_per_thread": 2}, {"words_per_thread": 4},
                            {"words_per_thread": 16}, {"words_per_thread": 1, "tile_out_cap": 3},
                            {"walk_limit_words": 1}, {"tile_out_cap": 1, "walk_limit_words": 2},
                            {"flags": 2}, {"flags": 8}, {"flags": 2, "tile_out_cap": 2, "walk_limit_words": 1}])
        try:
            by_k = not interval and fs["min_repeats"] > 1     # the by-k variant shares no dict: no keep-shorter rule
            want = (oracle.detect_repeats_by_k if by_k else oracle.detect_repeats)(seq, ns(**fs))
            exc = None
        except (AssertionError, IndexError) as e:
            want, exc = None, type(e)
        try:
            got = prf.detect_repeats(seq, ns(**fs), **knobs)
            gexc = None
        except (AssertionError, IndexError) as e:
            got, gexc = None, type(e)
        if got != want or exc != gexc:
            print("MISMATCH", fs, knobs, len(seq), exc, gexc)
            path = f"/tmp/fuzz_fail_{args.seed}_{n_cases}.txt"
            with open(path, "w") as f:
                f.write(repr((seq, fs, knobs)))
            if got is not None and want is not None:
                d = next((i for i, (x, y) in enumerate(zip(got + [None], want + [None])) if x != y), None)
                print(" first difference at row", d, (got + [None])[d], (want + [None])[d], "saved", path)
            sys.exit(1)
        if want is not None and not interval and len(seq) > kmax and rng.random() < 0.35:
            packed_vs_ascii(seq, kmin, kmax, fs["min_repeats"], fs["min_span"], rng)
            n_packed += 1
        n_cases += 1
        n_rows += len(want) if want else 0
    print(f"fuzz ok: {n_cases} cases ({n_packed} also through the packed loaders), {n_rows} rows compared, seed {args.seed}")


if __name__ == "__main__":
    main()
