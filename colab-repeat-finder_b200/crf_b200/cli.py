"""Command line of the drop-in: same flags, defaults, messages and output files as the reference's
main() (perfect_repeat_finder.py:83-179).  Deliberate deviations:
  * a FASTA file WITHOUT --interval works (the reference raises AttributeError at :139): every record is
    scanned as detect_repeats(record.seq, filters) -- all records in one GPU load;
  * --plot is accepted but plotting is out of scope here (matplotlib is not a dependency).
"""
import argparse
import os
import re
import sys

import numpy as np

from . import _cabi, api, fasta


def build_parser():
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    group = parser.add_argument_group("Repeat Filters")
    group.add_argument("-min", "--min-motif-size", default=1, type=int, help="Minimum motif size in base pairs.")
    group.add_argument("-max", "--max-motif-size", default=50, type=int, help="Maximum motif size in base pairs.")
    group.add_argument("--min-repeats", default=3, type=int, help="The minimum number of repeats to look for.")
    group.add_argument("--min-span", default=9, type=int, help="The repeats should span at least this many consecutive "
                                                               "bases in the input sequence.")
    parser.add_argument("-i", "--interval", help="Only consider sequence from this interval (chrom:start_0based-end).")
    parser.add_argument("-p", "--plot", help="Write out a plot with this filename.")
    parser.add_argument("-o", "--output-prefix", help="The output filename prefix for the output TSV file. If the input "
                                                      "is a FASTA file, a BED file will also be generated.")
    parser.add_argument("--verbose", action="store_true", help="Print verbose output.")
    parser.add_argument("--debug", action="store_true", help="Print debugging output.")
    parser.add_argument("--show-progress-bar", action="store_true", help="Show progress bar.")
    parser.add_argument("input_sequence", help="The nucleotide sequence, or a FASTA file path")
    return parser


def _scan_records_to_bed(records, args, bed_path):
    """All records in one load / one scan; rows written by the native writer.  Returns rows per record."""
    ctx = api.get_context()
    lengths = np.array([len(r.seq) for r in records], dtype=np.uint64)
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    blob = b"".join(r.seq for r in records)
    counts = np.zeros(len(records), dtype=np.int64)
    if not len(blob):
        open(bed_path, "wb").close()
        return counts
    with ctx.load(blob, offsets, max_motif_cap=args.max_motif_size) as seq:
        n = seq.scan(args.min_motif_size, args.max_motif_size, args.min_repeats, args.min_span)
        rec, start, end, k = seq.fetch(n)
    _cabi.write_rows(bed_path, [r.name for r in records], blob, offsets, rec, start, end, k)
    np.add.at(counts, rec.astype(np.int64), 1)
    return counts


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)

    if args.min_motif_size < 1:
        parser.error(f"--min-motif-size is set to {args.min_motif_size}. It must be at least 1.")
    if args.max_motif_size < args.min_motif_size:
        parser.error(f"--max-motif-size is set to {args.max_motif_size}. It must be at least --min-motif-size.")
    if args.min_repeats < 1:
        parser.error(f"--min-repeats is set to {args.min_repeats}. It must be at least 1.")
    if args.min_span < 1:
        parser.error(f"--min-span is set to {args.min_span}. It must be at least 1.")

    interval_sequence = None
    if os.path.isfile(args.input_sequence):
        if not args.output_prefix:
            args.output_prefix = re.sub(".fa(sta)?(.gz)?", "", args.input_sequence)   # same (unanchored) regex as prf:114

        output_bed_path = f"{os.path.basename(args.output_prefix)}.bed"
        fasta_entries = fasta.read_fasta(args.input_sequence)
        if args.interval:
            interval = re.split("[:-]", args.interval)
            if len(interval) != 3:
                parser.error("Invalid --interval format. Must be chrom:start_0based-end")
            args.interval_chrom, args.interval_start_0based, args.interval_end = interval
            args.interval_start_0based = int(args.interval_start_0based)
            args.interval_end = int(args.interval_end)
            matches = [e for e in fasta_entries if e.name == args.interval_chrom]
            if not matches:
                parser.error(f"Chromosome {args.interval_chrom} not found in the input FASTA file")
            fasta_entries = [matches[0]]

        if args.interval:
            with open(output_bed_path, "wt") as bed_file:
                entry = fasta_entries[0]
                seq_len = len(entry.seq)
                if args.interval_end > seq_len:
                    args.interval_end = seq_len
                seq_len = args.interval_end - args.interval_start_0based
                print(f"Processing {entry.name} ({seq_len:,d} bp)")
                output_intervals = api.detect_repeats(entry.seq, args)
                print(f"Found {len(output_intervals):,d} repeats")
                bed_file.write("".join(f"{entry.name}\t{s}\t{e}\t{m}\n" for s, e, m in output_intervals))
        else:
            counts = _scan_records_to_bed(fasta_entries, args, output_bed_path)
            for entry, n_found in zip(fasta_entries, counts.tolist()):
                print(f"Processing {entry.name} ({len(entry.seq):,d} bp)")
                print(f"Found {n_found:,d} repeats")

        print(f"Wrote results to {output_bed_path}")

    elif set(args.input_sequence.upper()) <= set("ACGTN"):
        if args.interval:
            parser.error("The --interval option is only supported for FASTA files.")

        interval_sequence = args.input_sequence
        if not args.output_prefix:
            args.output_prefix = "repeats"
        output_tsv_path = f"{args.output_prefix}.tsv"

        output_intervals = api.detect_repeats(args.input_sequence, args)
        print(f"Found {len(output_intervals):,d} repeats")

        with open(output_tsv_path, "wt") as tsv_file:
            tsv_file.write("\t".join(["start_0based", "end", "motif"]) + "\n")
            for start_0based, end, motif in output_intervals:
                tsv_file.write("\t".join([str(start_0based), str(end), motif]) + "\n")
        print(f"Wrote results to {output_tsv_path}")
    else:
        parser.error(f"Invalid input: {args.input_sequence}. This should be a FASTA file path or a string of nucleotides.")

    if args.plot and interval_sequence:
        if len(interval_sequence) > 5_000:
            print(f"Warning: The input sequence is too long ({len(interval_sequence):,d} bp). Skipping plot...")
        else:
            print("Warning: plotting is not part of this build (matplotlib is not a dependency). Skipping plot...")
    return 0


if __name__ == "__main__":
    sys.exit(main())
