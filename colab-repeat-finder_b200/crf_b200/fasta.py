"""FASTA input through the library's native reader (crf_fasta_open), standing in for pyfastx
(perfect_repeat_finder.py:117-137 uses only `.name` -- the first whitespace-delimited token of the header --
and `.seq`, the record's lines joined without line ends, case preserved).  Plain or gzip input."""
import os

from . import _cabi


class FastaRecord:
    __slots__ = ("name", "seq")

    def __init__(self, name, seq):
        self.name = name      # str
        self.seq = seq        # uint8 view into the file's base buffer, one byte per base, case preserved


def open_fasta(path, pinned=False):
    """The whole file as a _cabi.Fasta (bases back to back + offsets + names)."""
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    return _cabi.Fasta(path, pinned=pinned)


def read_fasta(path):
    """All records of a FASTA file, in file order, as (name, bytes) records."""
    with open_fasta(path) as fa:
        return [FastaRecord(name, fa.record(i).tobytes()) for i, name in enumerate(fa.names)]
