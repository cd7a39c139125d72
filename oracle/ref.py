"""Loader of the UNMODIFIED Python reference copied to oracle/_ref/ by oracle/make_ref.sh.

TEST INFRASTRUCTURE ONLY (same rule as oracle.py): tests/, __graft_entry__ and bench.py's cpu_baseline /
`--impl reference` legs.  This is tier L0 of SURVEY.md section 8c -- the reference's own
perfect_repeat_finder.detect_repeats (prf:10-81) driving its own PerfectRepeatTracker (trk:3-105).

The reference's module names (`perfect_repeat_finder`, `utils.*`) are the names of this repo's drop-in
modules too, so the copy is imported under a swap of sys.modules / sys.path and its functions are handed
out as plain objects; afterwards the drop-in modules are back in place and both can be used side by side.
"""
import os
import subprocess
import sys
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_lock = threading.Lock()
_mods = None


def available():
    return os.path.isfile(os.path.join(REF_DIR, "perfect_repeat_finder.py"))


def build(source="/root/reference"):
    """Run make_ref.sh when the reference checkout is present (the build container); the GPU box only has the
    prebuilt copy.  Returns True when oracle/_ref is usable afterwards."""
    if os.path.isfile(os.path.join(source, "perfect_repeat_finder.py")):
        subprocess.check_call(["bash", os.path.join(_HERE, "make_ref.sh"), source], stdout=subprocess.DEVNULL)
    return available()


_SHADOWED = ("perfect_repeat_finder", "perfect_repeat_finder_tests", "utils", "pyfastx", "matplotlib")


def _is_shadowed(name):
    return any(name == s or name.startswith(s + ".") for s in _SHADOWED)


def load():
    """(perfect_repeat_finder module, its tests module) of the reference copy."""
    global _mods
    with _lock:
        if _mods is not None:
            return _mods
        if not available():
            raise FileNotFoundError(f"{REF_DIR} is missing: run oracle/make_ref.sh where /root/reference exists")
        saved = {n: m for n, m in sys.modules.items() if _is_shadowed(n)}
        for n in saved:
            del sys.modules[n]
        saved_path = list(sys.path)
        sys.path[:0] = [REF_DIR, os.path.join(REF_DIR, "_stubs")]
        try:
            import perfect_repeat_finder as ref_prf
            import perfect_repeat_finder_tests as ref_tests
            assert os.path.dirname(os.path.abspath(ref_prf.__file__)) == REF_DIR
        finally:
            sys.path[:] = saved_path
            for n in [n for n in sys.modules if _is_shadowed(n)]:
                del sys.modules[n]
            sys.modules.update(saved)
        _mods = (ref_prf, ref_tests)
        return _mods


def detect_repeats(input_sequence, filter_settings):
    """The reference's detect_repeats on a str (exactly as shipped, one core)."""
    return load()[0].detect_repeats(input_sequence, filter_settings)


# ---- timing helper: P worker processes, one slice each (the reference's own scale-out is 1-CPU jobs over
# ---- disjoint intervals, hail_batch_pipeline/run_hail_batch_pipeline.py:101) --------------------------------
def _worker(job):
    import argparse
    import time
    seq, fs = job
    t0 = time.perf_counter()
    rows = detect_repeats(seq, argparse.Namespace(**fs))
    return time.perf_counter() - t0, rows


def make_pool(processes):
    import multiprocessing as mp
    return mp.get_context("fork").Pool(processes)


def run_slices(pool, slices, fs):
    """One timed step: every slice through the reference, one process each.  Returns (wall seconds,
    per-slice seconds, per-slice rows)."""
    import time
    t0 = time.perf_counter()
    out = pool.map(_worker, [(s, fs) for s in slices], chunksize=1)
    wall = time.perf_counter() - t0
    return wall, [o[0] for o in out], [o[1] for o in out]


# ---- the reference's command line (prf:83-179), for file-level parity of the CLI ------------------------------
class _MiniFasta:
    """What main() uses of pyfastx.Fasta (prf:117-137): iteration over records with .name / .seq, `name in fa`,
    fa[name].seq.  pyfastx is not installed; this reads the file line by line (plain or gzip), names are the first
    word of the header, sequence text is kept as written (case included)."""

    def __init__(self, path):
        import argparse
        import gzip
        with open(path, "rb") as f:
            magic = f.read(2)
        opener = gzip.open if magic == b"\x1f\x8b" else open
        self._entries, name, parts = [], None, []
        with opener(path, "rt") as f:
            for line in f:
                line = line.rstrip("\r\n")
                if line.startswith(">"):
                    if name is not None:
                        self._entries.append(argparse.Namespace(name=name, seq="".join(parts)))
                    name, parts = (line[1:].split() or [""])[0], []
                elif name is not None:
                    parts.append(line.strip())
        if name is not None:
            self._entries.append(argparse.Namespace(name=name, seq="".join(parts)))

    def __iter__(self):
        return iter(self._entries)

    def __contains__(self, name):
        return any(e.name == name for e in self._entries)

    def __getitem__(self, name):
        return next(e for e in self._entries if e.name == name)


def run_main(argv, cwd):
    """The reference's main() with sys.argv = argv, run in directory `cwd` (it writes <prefix>.bed / .tsv there).
    Returns (exit status, stdout text); parser.error() exits with 2, an uncaught exception propagates."""
    import argparse
    import contextlib
    import io
    ref_prf, _ = load()
    with _lock:
        saved_argv, saved_cwd, saved_fastx = sys.argv, os.getcwd(), ref_prf.pyfastx
        out, status = io.StringIO(), 0
        try:
            sys.argv = ["perfect_repeat_finder.py"] + [str(a) for a in argv]
            os.chdir(cwd)
            ref_prf.pyfastx = argparse.Namespace(Fasta=_MiniFasta)
            with contextlib.redirect_stdout(out), contextlib.redirect_stderr(io.StringIO()):
                try:
                    ref_prf.main()
                except SystemExit as e:
                    status = e.code if isinstance(e.code, int) else 1
        finally:
            sys.argv = saved_argv
            os.chdir(saved_cwd)
            ref_prf.pyfastx = saved_fastx
        return status, out.getvalue()
