"""Host-side timing of the FASTA reader on plain gzip input (profiles/r02_gunzip_threads.md): a FASTA text of --mbp Mbp is
written, compressed with `gzip -1`, and opened with the reader's decoder on all threads, on one thread, and with zlib.
--cli also runs the command line on the .fa.gz (needs the GPU)."""
import argparse
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from crf_b200 import _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=int, default=120)
    ap.add_argument("--cli", action="store_true")
    a = ap.parse_args()
    rng = np.random.default_rng(5)
    n = a.mbp * 1_000_000 // 60 * 60
    text = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n, dtype=np.uint8)]
    for s in rng.integers(0, n - 5000, n // 24000):
        text[s:s + int(rng.integers(100, 5000))] |= 0x20                       # soft-masked stretches
    text[n // 10:n // 10 + n // 240] = ord("N")
    for s in rng.integers(0, n - 40000, n // 16000):                            # copied segments: matches beyond 3-mers
        ln, d = int(rng.integers(50, 3000)), int(rng.integers(1, 30000))
        text[s + d:s + d + ln] = text[s:s + ln]
    lines = np.concatenate([text.reshape(-1, 60), np.full((n // 60, 1), 10, np.uint8)], axis=1)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "t.fa")
        with open(path, "wb") as f:
            f.write(b">chrT\n")
            f.write(lines.tobytes())
        t = time.time()
        subprocess.check_call(["gzip", "-1", "-k", path])
        gz = path + ".gz"
        print(f"{os.path.getsize(path) / 1e6:.0f} MB of FASTA text -> {os.path.getsize(gz) / 1e6:.0f} MB (gzip -1, {time.time() - t:.1f} s), "
              f"{os.cpu_count()} host threads", flush=True)
        for label, env in [("own decoder, all threads", {}), ("own decoder, one thread", {"CRF_GUNZIP_CHUNK_KB": "100000000"}),
                           ("zlib", {"CRF_GUNZIP_ZLIB": "1"})]:
            for k in ("CRF_GUNZIP_CHUNK_KB", "CRF_GUNZIP_ZLIB"):
                os.environ.pop(k, None)
            os.environ.update(env)
            for rep in range(2 if label != "zlib" else 1):
                t = time.time()
                fa = _cabi.Fasta(gz)
                dt = time.time() - t
                print(f"{label}: crf_fasta_open {dt * 1e3:.0f} ms", flush=True)
                del fa
        for k in ("CRF_GUNZIP_CHUNK_KB", "CRF_GUNZIP_ZLIB"):
            os.environ.pop(k, None)
        if a.cli:
            t = time.time()
            rc = subprocess.call([sys.executable, os.path.join(ROOT, "perfect_repeat_finder.py"), "-min", "1", "-max", "50", gz], cwd=tmp)
            print(f"command line on the .fa.gz: rc {rc}, {time.time() - t:.2f} s wall, BED {os.path.getsize(os.path.join(tmp, 't.bed'))} bytes", flush=True)


if __name__ == "__main__":
    main()
