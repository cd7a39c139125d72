#!/usr/bin/env python3
"""End-to-end wall time of the command line, FASTA file on disk -> BED file on disk (reported separately from the
device-time metric).  The reference's own benchmark is this, on chr22: `--min-motif-size 1 --max-motif-size 6 --min-repeats 3
--min-span 9` took 4 min 08 s (benchmark/repeat_finder/repeat_finder.log of the reference).  benchmark/chr22.fa.gz is not in the
checkout, so the input is the chr22-shaped record S22 (or the real file via $CRF_CHR22_FASTA), written as a 60-column FASTA.
   python profiles/prof_cli_wall.py [--scale 1.0]"""
import argparse
import contextlib
import io
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from crf_b200 import cli, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--workload", default="s22", choices=["s22", "s38"])
args = ap.parse_args()
if args.workload == "s22":
    bases, offsets, meta = synth.chr22(device="cuda:0", scale=args.scale)
else:
    bases, offsets, meta = synth.s38(device="cuda:0", scale=args.scale)
seq = bases.cpu().numpy()
del bases
torch.cuda.synchronize()
work = tempfile.mkdtemp(prefix="crf_cli_")
fa = os.path.join(work, "chr22.fa" if args.workload == "s22" else "genome.fa")
t0 = time.perf_counter()
with open(fa, "wb") as f:
    for r, name in enumerate(meta["names"]):
        rec = seq[int(offsets[r]):int(offsets[r + 1])]
        f.write(b">" + name.encode() + b"\n")
        n_full = rec.size // 60
        lines = np.empty((n_full, 61), dtype=np.uint8)
        lines[:, :60] = rec[:n_full * 60].reshape(n_full, 60)
        lines[:, 60] = 10
        lines.tofile(f)
        del lines
        if rec.size % 60:
            f.write(rec[n_full * 60:].tobytes() + b"\n")
print(f"{meta['workload']}: {seq.size} bp written to {fa} ({os.path.getsize(fa) / 1e6:.1f} MB) in {time.perf_counter() - t0:.2f} s",
      flush=True)
os.chdir(work)
settings = (("motif 1-6 (the reference's benchmark setting)", ["-min", "1", "-max", "6"]),
            ("motif 2-6 (BASELINE config C1)", ["-min", "2", "-max", "6"]),
            ("motif 1-50 (defaults, config C2)", []))
if args.workload == "s38":
    settings = (("motif 1-50 (defaults, config C3)", []),)
for label, extra in settings:
    for rep in range(2 if args.workload == "s38" else 3):
        out = io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(out):
            rc = cli.main([fa, "--min-repeats", "3", "--min-span", "9", "-o", "out"] + extra)
        dt = time.perf_counter() - t0
        rows = sum(1 for _ in open("out.bed", "rb"))
        print(f"{label}: run {rep}: {dt * 1e3:8.1f} ms wall, FASTA on disk -> {rows} BED rows on disk "
              f"({os.path.getsize('out.bed') / 1e6:.1f} MB), rc {rc}; {out.getvalue().splitlines()[1]}", flush=True)
import shutil  # noqa: E402
os.chdir(ROOT)
shutil.rmtree(work, ignore_errors=True)
