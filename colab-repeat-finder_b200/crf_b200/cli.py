"""Command line of the drop-in.

Same flags, defaults, stdout lines and output files as the reference's main()
(/root/reference/perfect_repeat_finder.py:83-179), re-organised around two code paths:

  FASTA file  -> <basename(prefix)>.bed in the working directory (chrom, start_0based, end, motif)
  raw string  -> <prefix>.tsv with a header line (start_0based, end, motif)

Deliberate deviations:
  * a FASTA file WITHOUT --interval works (the reference raises AttributeError at :139): all records go to the GPU
    in one load and one scan, rows are written by the native writer (crf_write_rows);
  * --plot is accepted but plotting is out of scope here (matplotlib is not a dependency).
"""
import argparse
import os
import re
import sys

import numpy as np

from . import _cabi, api, fasta

# (flags, keyword arguments) -- the reference's options, prf:86-98
_FILTER_OPTIONS = [
    (("-min", "--min-motif-size"), dict(default=1, type=int, help="Smallest motif size (bp) to look for.")),
    (("-max", "--max-motif-size"), dict(default=50, type=int, help="Largest motif size (bp) to look for.")),
    (("--min-repeats",), dict(default=3, type=int, help="Report a locus only if the motif occurs at least this many "
                                                         "times in a row.")),
    (("--min-span",), dict(default=9, type=int, help="Report a locus only if it covers at least this many bases.")),
]
_OTHER_OPTIONS = [
    (("-i", "--interval"), dict(help="Restrict the scan to chrom:start_0based-end (FASTA input only).")),
    (("-p", "--plot"), dict(help="Plot file name (accepted for compatibility; plotting is not part of this build).")),
    (("-o", "--output-prefix"), dict(help="Prefix of the output file: <prefix>.tsv for a sequence given on the command "
                                          "line, <basename(prefix)>.bed for a FASTA file.")),
    (("--verbose",), dict(action="store_true", help="Accepted for compatibility.")),
    (("--debug",), dict(action="store_true", help="Accepted for compatibility.")),
    (("--show-progress-bar",), dict(action="store_true", help="Accepted for compatibility.")),
    # not in the reference: its scale-out is a separate Hail Batch pipeline (run_hail_batch_pipeline.py:96-123)
    (("--devices",), dict(default=os.environ.get("CRF_DEVICES"),
                          help="GPUs for whole-FASTA runs, e.g. 0-7 or 0,2,3 (default: $CRF_DEVICES, else one GPU). "
                               "The records are cut into chunks, every GPU scans its share and the rows are gathered on "
                               "the first one over NVLink; the BED file is identical to a single-GPU run.")),
]


def build_parser():
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter,
                                     description="Find perfect tandem repeats (B200 build).")
    filters = parser.add_argument_group("Repeat Filters")
    for flags, kw in _FILTER_OPTIONS:
        filters.add_argument(*flags, **kw)
    for flags, kw in _OTHER_OPTIONS:
        parser.add_argument(*flags, **kw)
    parser.add_argument("input_sequence", help="A nucleotide string, or the path of a FASTA file (plain or .gz)")
    return parser


def _check_filters(parser, args):
    """prf:102-109 -- same messages, exit status 2."""
    for value, flag, floor in ((args.min_motif_size, "--min-motif-size", "1"),):
        if value < 1:
            parser.error(f"{flag} is set to {value}. It must be at least {floor}.")
    if args.max_motif_size < args.min_motif_size:
        parser.error(f"--max-motif-size is set to {args.max_motif_size}. It must be at least --min-motif-size.")
    if args.min_repeats < 1:
        parser.error(f"--min-repeats is set to {args.min_repeats}. It must be at least 1.")
    if args.min_span < 1:
        parser.error(f"--min-span is set to {args.min_span}. It must be at least 1.")


def parse_devices(text):
    """"0-3" / "0,2,5" / "1" -> list of device indices (None or "" -> one device: api.default_device())."""
    if text is None or str(text).strip() == "":
        return [api.default_device()]
    out = []
    for part in str(text).split(","):
        part = part.strip()
        if "-" in part:
            lo, hi = part.split("-", 1)
            out.extend(range(int(lo), int(hi) + 1))
        else:
            out.append(int(part))
    if not out or min(out) < 0:
        raise ValueError(f"invalid device list {text!r}")
    return out


def group_records(lengths, max_motif_size, limit=None):
    """Cut the record list into consecutive groups that each fit one device load: a load lays every record out
    as length + max_motif_size positions and holds at most crf_load_limit() of them (32-bit positions on the
    device).  A single record above the limit is its own group (and is refused by the load with a clear message).
    Returns [(first, last)), ...]."""
    limit = _cabi.load_limit(max_motif_size) if limit is None else limit
    groups, first, n = [], 0, len(lengths)
    while first < n:
        last, need = first, 0
        while last < n and (last == first or need + int(lengths[last]) + max_motif_size <= limit):
            need += int(lengths[last]) + max_motif_size
            last += 1
        groups.append((first, last))
        first = last
    return groups


def _whole_fasta_to_bed(fa, args, bed_path):
    """Every record of the file: as few loads as the position limit allows (one for a human genome), one scan per
    load (on one GPU, or split over --devices), native row writer.  The groups are slices of the reader's buffer (no
    copies on the Python side).  Returns the row count per record."""
    devices = parse_devices(getattr(args, "devices", None))
    lengths = np.diff(fa.offsets.astype(np.int64))
    counts = np.zeros(fa.n_records, dtype=np.int64)
    open(bed_path, "wb").close()
    # the text is packed to 2-bit planes + mask on the host (threaded, at ingest) and uploaded at 0.375 B/bp; the ASCII
    # buffer stays on the host for the motif column
    planes = fa.packed() if fa.total_bases else None
    for first, last in group_records(lengths, args.max_motif_size):
        total = int(lengths[first:last].sum())
        base0 = int(fa.offsets[first])
        blob = fa.bases[base0:base0 + total]
        offsets = fa.offsets[first:last + 1] - np.uint64(base0)
        if total:
            if len(devices) > 1:
                from . import multi
                many_short = (last - first) > 4096 and int(lengths[first:last].max()) < (1 << 20)
                rec, start, end, k = multi.scan_on_devices(
                    devices, planes, fa.offsets[first:last], lengths[first:last], args.min_motif_size, args.max_motif_size,
                    args.min_repeats, args.min_span, reads=many_short, contexts=[api.get_context(d) for d in devices])
            else:
                with api.get_context(devices[0]).load_packed(planes, fa.offsets[first:last + 1],
                                                             max_motif_cap=args.max_motif_size) as seq:
                    n = seq.scan(args.min_motif_size, args.max_motif_size, args.min_repeats, args.min_span)
                    rec, start, end, k = seq.fetch(n)
            whole_file = first == 0 and last == fa.n_records          # the usual case: hand the name table over as it is
            _cabi.write_rows(bed_path, fa.names_blob if whole_file else fa.names[first:last], blob, offsets, rec, start,
                             end, k, append=True)
            counts[first:last] += np.bincount(rec.astype(np.int64, copy=False), minlength=last - first)
    return counts


def _run_fasta(parser, args):
    if not args.output_prefix:
        args.output_prefix = re.sub(".fa(sta)?(.gz)?", "", args.input_sequence)   # the (unanchored) regex of prf:114
    bed_path = f"{os.path.basename(args.output_prefix)}.bed"                       # always in the working directory
    # (page-locked planes only: the planes are what is uploaded, the text stays on the host for the motif column)
    with fasta.open_fasta(args.input_sequence, pinned=2) as fa:
        if args.interval:
            fields = re.split("[:-]", args.interval)
            if len(fields) != 3:
                parser.error("Invalid --interval format. Must be chrom:start_0based-end")
            args.interval_chrom = fields[0]
            args.interval_start_0based, args.interval_end = int(fields[1]), int(fields[2])
            if args.interval_chrom not in fa.names:
                parser.error(f"Chromosome {args.interval_chrom} not found in the input FASTA file")
            chosen = fa.record(fa.names.index(args.interval_chrom))
            args.interval_end = min(args.interval_end, len(chosen))                    # prf:139-140
            print(f"Processing {args.interval_chrom} ({args.interval_end - args.interval_start_0based:,d} bp)")
            rows = api.detect_repeats(chosen.tobytes(), args)
            print(f"Found {len(rows):,d} repeats")
            with open(bed_path, "wt") as bed_file:
                bed_file.write("".join(f"{args.interval_chrom}\t{s}\t{e}\t{m}\n" for s, e, m in rows))
        elif args.min_repeats == 1:
            # the reference's single-copy quirks are per record (wrap-around at each record's position 0): one call each
            with open(bed_path, "wt") as bed_file:
                for i, name in enumerate(fa.names):
                    print(f"Processing {name} ({len(fa.record(i)):,d} bp)")
                    rows = api.detect_repeats(fa.record(i).tobytes(), args)
                    print(f"Found {len(rows):,d} repeats")
                    bed_file.write("".join(f"{name}\t{s}\t{e}\t{m}\n" for s, e, m in rows))
        else:
            counts = _whole_fasta_to_bed(fa, args, bed_path)
            lengths = np.diff(fa.offsets.astype(np.int64))
            names, lengths, counts = fa.names, lengths.tolist(), counts.tolist()
            for lo in range(0, len(names), 65536):            # the reference's two lines per record; batched, a reads
                sys.stdout.write("".join(                     # file has millions of records
                    f"Processing {n_} ({l_:,d} bp)\nFound {c_:,d} repeats\n"
                    for n_, l_, c_ in zip(names[lo:lo + 65536], lengths[lo:lo + 65536], counts[lo:lo + 65536])))
    print(f"Wrote results to {bed_path}")


def _run_raw(parser, args):
    if args.interval:
        parser.error("The --interval option is only supported for FASTA files.")
    if not args.output_prefix:
        args.output_prefix = "repeats"
    tsv_path = f"{args.output_prefix}.tsv"
    rows = api.detect_repeats(args.input_sequence, args)
    print(f"Found {len(rows):,d} repeats")
    with open(tsv_path, "wt") as tsv_file:
        tsv_file.write("start_0based\tend\tmotif\n")
        tsv_file.writelines(f"{s}\t{e}\t{m}\n" for s, e, m in rows)
    print(f"Wrote results to {tsv_path}")
    if args.plot:
        if len(args.input_sequence) > 5_000:
            print(f"Warning: The input sequence is too long ({len(args.input_sequence):,d} bp). Skipping plot...")
        else:
            print("Warning: plotting is not part of this build (matplotlib is not a dependency). Skipping plot...")


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    _check_filters(parser, args)
    if os.path.isfile(args.input_sequence):
        _run_fasta(parser, args)
    elif set(args.input_sequence.upper()) <= set("ACGTN"):
        _run_raw(parser, args)
    else:
        parser.error(f"Invalid input: {args.input_sequence}. This should be a FASTA file path or a string of nucleotides.")
    return 0


if __name__ == "__main__":
    sys.exit(main())
