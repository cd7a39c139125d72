"""Minimal FASTA reader standing in for pyfastx (perfect_repeat_finder.py:117-137 uses only
`.name` -- the first whitespace-delimited token of the header -- and `.seq`, the record's lines joined
without line ends, case preserved).  Plain or gzip input."""
import gzip
import os


class FastaRecord:
    __slots__ = ("name", "seq")

    def __init__(self, name, seq):
        self.name = name      # str
        self.seq = seq        # bytes, one byte per base, case preserved


def _read_all(path):
    with open(path, "rb") as f:
        magic = f.read(2)
    opener = gzip.open if magic == b"\x1f\x8b" else open
    with opener(path, "rb") as f:
        return f.read()


def read_fasta(path):
    """All records of a FASTA file, in file order."""
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    raw = _read_all(path)
    records = []
    pos = raw.find(b">") if not raw.startswith(b">") else 0
    # a '>' only starts a record at the beginning of a line
    while pos != -1 and pos > 0 and raw[pos - 1:pos] not in (b"\n", b"\r"):
        pos = raw.find(b">", pos + 1)
    while pos != -1 and pos < len(raw):
        eol = raw.find(b"\n", pos)
        if eol == -1:
            eol = len(raw)
        header = raw[pos + 1:eol].decode("utf-8", "replace").strip()
        name = header.split()[0] if header.split() else ""
        nxt = raw.find(b"\n>", eol)
        body = raw[eol + 1:(nxt + 1 if nxt != -1 else len(raw))]
        records.append(FastaRecord(name, body.translate(None, b"\n\r")))
        pos = nxt + 1 if nxt != -1 else -1
    return records
