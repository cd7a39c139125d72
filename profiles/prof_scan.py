#!/usr/bin/env python3
"""Small driver for ncu / timing experiments on the scan kernel alone:
   python profiles/prof_scan.py [--scale 0.25] [--reps 3] [--fast-only] [--wpt 8]
Generates S38 at the given scale in HBM, loads it once, runs `reps` scans and prints the kernel times."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))
import torch  # noqa: E402
from crf_b200 import _cabi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.25)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--fast-only", action="store_true")
ap.add_argument("--no-sup", action="store_true")
ap.add_argument("--outcap", type=int, default=0)
ap.add_argument("--wpt", type=int, default=0)
ap.add_argument("--flags", type=int, default=0, help="crf_scan_params.flags (2: warp tiles, 8: two strips per thread)")
ap.add_argument("--kmin", type=int, default=1)
ap.add_argument("--kmax", type=int, default=0)
ap.add_argument("--workload", default="s38", choices=["s38", "s22", "sr"])
args = ap.parse_args()
if not args.kmax:
    args.kmax = 20 if args.workload == "sr" else 50
if args.workload == "s38":
    bases, offsets, meta = synth.s38(device="cuda:0", scale=args.scale)
elif args.workload == "s22":
    bases, offsets, meta = synth.chr22(device="cuda:0", scale=args.scale)
else:
    bases, offsets, meta = synth.sr(int(10_000_000 * args.scale), device="cuda:0")
torch.cuda.synchronize()
ctx = _cabi.Context(0)
seq = ctx.load(bases.data_ptr(), offsets, max_motif_cap=args.kmax, on_device=True)
flags = ((1 << 16) if args.fast_only else 0) | ((1 << 17) if args.no_sup else 0) | args.flags
for i in range(args.reps):
    n = seq.scan(args.kmin, args.kmax, 3, 9, flags=flags, words_per_thread=args.wpt, tile_out_cap=args.outcap)
    st = seq.stats()
    print(f"rep {i}: kernel {st.kernel_ms:.3f} ms scan {st.scan_ms:.3f} ms results {n} cand {st.n_candidates} "
          f"bp {int(offsets[-1])} -> {int(offsets[-1]) / st.kernel_ms / 1e6:.1f} Gbp/s (kernel)")
