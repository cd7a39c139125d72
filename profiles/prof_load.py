#!/usr/bin/env python3
"""Where does a load spend its time?   CRF_LOAD_TRACE=1 python profiles/prof_load.py --workload sr [--scale 1] [--reps 3]
Generates the workload in HBM, brings the text to page-locked host memory, packs it on the host, then times
Context.load (ASCII from the host), Context.load_packed (planes) and load_packed with the mask as runs, wall clock around
the call; with CRF_LOAD_TRACE set the library prints its own stage times to stderr."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))
import torch  # noqa: E402
from crf_b200 import _cabi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="sr", choices=["s38", "s22", "sr"])
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
kmax = 20 if args.workload == "sr" else 50
if args.workload == "s38":
    bases, offsets, meta = synth.s38(device="cuda:0", scale=args.scale)
elif args.workload == "s22":
    bases, offsets, meta = synth.chr22(device="cuda:0", scale=args.scale)
else:
    bases, offsets, meta = synth.sr(int(10_000_000 * args.scale), device="cuda:0")
host = torch.empty(bases.shape, dtype=torch.uint8, pin_memory=True)
host.copy_(bases)
torch.cuda.synchronize()
del bases
text = host.numpy()
t0 = time.perf_counter()
pk = _cabi.pack_ascii(text)
pk_runs = _cabi.pack_ascii(text).with_runs()
print(f"host pack x2: {(time.perf_counter() - t0) * 1e3:.1f} ms; records {len(offsets) - 1}, bases {int(offsets[-1])}", flush=True)
ctx = _cabi.Context(0)
for name, fn in (("ascii", lambda: ctx.load(text, offsets, max_motif_cap=kmax)),
                 ("packed", lambda: ctx.load_packed(pk, offsets, max_motif_cap=kmax)),
                 ("packed + runs", lambda: ctx.load_packed(pk_runs, offsets, max_motif_cap=kmax))):
    for i in range(args.reps):
        t0 = time.perf_counter()
        seq = fn()
        t1 = time.perf_counter()
        device_ms = seq.info().load_ms
        seq.close()
        t2 = time.perf_counter()
        print(f"{name:14s} rep {i}: load {(t1 - t0) * 1e3:8.2f} ms (device side {device_ms:.2f}), "
              f"close {(t2 - t1) * 1e3:6.2f} ms", flush=True)
        sys.stderr.flush()
