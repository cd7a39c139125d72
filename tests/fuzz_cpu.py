#!/usr/bin/env python3
"""CPU-only differential campaigns (not collected by pytest; the GPU one is tests/fuzz_gpu.py):

  --pair oracle-ref   the C oracle against the Python reference itself (oracle/_ref, build container only)
  --pair host-oracle  the host half of the drop-in (crf_b200/api.py: interval stop position, min_repeats == 1 corner cases,
                      N-trimming, exceptions) against the oracle, with the GPU scan replaced by the closed-form stand-in of
                      tests/test_host_cpu.py

Small sequences with planted periodic pieces at the start / end, intervals that begin or end near the sequence ends, all
min_repeats modes.  Example:  python tests/fuzz_cpu.py --pair host-oracle --seconds 150 --seed 3"""
import argparse
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))

from oracle import oracle, ref  # noqa: E402
from tests.helpers import ns, random_seq  # noqa: E402

RAISES = (AssertionError, IndexError, ValueError, AttributeError, NotImplementedError)


def host_detect():
    from crf_b200 import api
    from tests import test_host_cpu as stand_in
    api.get_context = lambda device=None: stand_in._ClosedFormCtx()
    closed_form = api.scan_arrays

    def scan_arrays(s, kmin, kmax, min_repeats, span, device=None, **kw):
        if min_repeats == 1:
            return stand_in._runs_single_copy(bytes(s), kmin, kmax, span)
        return closed_form(s, kmin, kmax, min_repeats, span, device=device, **kw)

    api.scan_arrays = scan_arrays
    return api.detect_repeats


def one_case(rng, refused_ok):
    m = rng.choice([rng.randint(0, 60), rng.randint(0, 300), rng.randint(300, 1200)])
    seq = random_seq(rng, m, exotic=rng.random() < 0.3)
    for _ in range(rng.choice([0, 0, 1, 1, 2])):
        unit = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 60)))
        pos = rng.choice([rng.randint(0, len(seq)), len(seq), 0])
        seq = seq[:pos] + unit * rng.randint(2, 12) + unit[:rng.randint(0, len(unit))] + seq[pos:]
    kmin = rng.choice([1, 1, 1, 2, 3, 7, 20])
    kmax = kmin + rng.choice([0, 1, 5, 15, 30, 49, 63, 64, 100])
    fs = dict(min_motif_size=kmin, max_motif_size=kmax, min_repeats=rng.choice([1, 1, 2, 2, 3, 3, 4]),
              min_span=rng.choice([1, 2, 5, 9, 9, 12, 33, 100]))
    if not refused_ok and fs["min_repeats"] == 1 and kmin == 1 and fs["min_span"] == 1:
        fs["min_span"] = 2                               # the one setting the drop-in refuses
    if rng.random() < 0.8:
        n = len(seq)
        a = rng.choice([rng.randint(0, n), 0, max(0, n - rng.randint(0, 150))])
        b = rng.choice([rng.randint(a, max(a, n)), n, max(a, n - rng.randint(0, 2 * kmax + 3)), a])
        if rng.random() < 0.8:
            fs["interval_start_0based"] = a
        if rng.random() < 0.9:
            fs["interval_end"] = b
    return seq, fs


def run(fn, seq, fs):
    try:
        return fn(seq, ns(**fs)), None
    except RAISES as e:
        return None, type(e)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pair", choices=["oracle-ref", "host-oracle"], default="host-oracle")
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    if args.pair == "oracle-ref":
        if not ref.available():
            sys.exit("oracle/_ref is not there (run oracle/make_ref.sh where /root/reference exists)")
        truth, tested = ref.detect_repeats, oracle.detect_repeats
    else:
        truth, tested = oracle.detect_repeats, host_detect()
    rng = random.Random(args.seed)
    t_end = time.time() + args.seconds
    n_cases = n_rows = n_raising = 0
    while time.time() < t_end:
        seq, fs = one_case(rng, refused_ok=args.pair == "oracle-ref")
        want, exc = run(truth, seq, fs)
        got, gexc = run(tested, seq, fs)
        if got != want or exc != gexc:
            path = f"/tmp/fuzz_cpu_fail_{args.pair}_{args.seed}.txt"
            with open(path, "w") as f:
                f.write(repr((seq, fs)))
            print("MISMATCH", args.pair, fs, len(seq), exc, gexc, "saved", path)
            if got is not None and want is not None:
                print(" missing", [x for x in want if x not in got][:5], "extra", [x for x in got if x not in want][:5])
            sys.exit(1)
        n_cases += 1
        n_rows += len(want) if want else 0
        n_raising += exc is not None
    print(f"fuzz ok ({args.pair}): {n_cases} cases, {n_rows} rows, {n_raising} raising, seed {args.seed}")


if __name__ == "__main__":
    main()
