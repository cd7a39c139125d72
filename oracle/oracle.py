"""ctypes front end of the CPU oracle (oracle/crf_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of crf_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  It mirrors the call shape of the reference's detect_repeats
(/root/reference/perfect_repeat_finder.py:10-81): same arguments, same exceptions.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcrf_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/crf_oracle.c -> oracle/libcrf_oracle.so (gcc, pthreads)."""
    src = os.path.join(_HERE, "crf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call(
            [cc, "-O3", "-march=x86-64-v2", "-pthread", "-fPIC", "-shared", "-o", _LIB_PATH, src])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        lib.crf_oracle_detect.restype = p
        lib.crf_oracle_detect.argtypes = [p, i64, i64, i64, i64, i64, ctypes.c_int, i64, i64]
        lib.crf_oracle_detect_by_k.restype = p
        lib.crf_oracle_detect_by_k.argtypes = [p, i64, i64, i64, i64, i64, ctypes.c_int]
        lib.crf_oracle_count.restype = ctypes.c_size_t
        lib.crf_oracle_count.argtypes = [p]
        lib.crf_oracle_status.restype = ctypes.c_int
        lib.crf_oracle_status.argtypes = [p]
        lib.crf_oracle_steps.restype = i64
        lib.crf_oracle_steps.argtypes = [p]
        lib.crf_oracle_fetch.restype = None
        lib.crf_oracle_fetch.argtypes = [p, p, p, p]
        lib.crf_oracle_free.restype = None
        lib.crf_oracle_free.argtypes = [p]
        lib.crf_oracle_max_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def max_threads():
    return _load().crf_oracle_max_threads()


def validate_filters(fs):
    """The four checks of perfect_repeat_finder.py:23-30, same messages."""
    if not getattr(fs, "min_motif_size") or fs.min_motif_size < 1:
        raise ValueError(f"min_motif_size is set to {fs.min_motif_size}. It must be at least 1.")
    if not getattr(fs, "max_motif_size") or fs.max_motif_size < fs.min_motif_size:
        raise ValueError(f"max_motif_size is set to {fs.max_motif_size}. It must be at least min_motif_size.")
    if not getattr(fs, "min_repeats") or fs.min_repeats < 1:
        raise ValueError(f"min_repeats is set to {fs.min_repeats}. It must be at least 1.")
    if not getattr(fs, "min_span") or fs.min_span < 1:
        raise ValueError(f"min_span is set to {fs.min_span}. It must be at least 1.")


def _as_bytes(seq):
    if isinstance(seq, (bytes, bytearray, memoryview)):
        return bytes(seq)
    if isinstance(seq, np.ndarray):
        return seq
    return seq.encode("latin-1")


def _collect(lib, h, raw, want_arrays):
    try:
        status = lib.crf_oracle_status(h)
        if status == 2:
            raise IndexError("string index out of range")
        if status == 3:
            raise AssertionError("RepeatTracker did not reach end of the sequence")
        if status != 0:
            raise MemoryError(f"oracle status {status}")
        n = lib.crf_oracle_count(h)
        start = np.empty(n, np.int64)
        end = np.empty(n, np.int64)
        mlen = np.empty(n, np.int32)
        if n:
            lib.crf_oracle_fetch(h, start.ctypes.data, end.ctypes.data, mlen.ctypes.data)
        steps = lib.crf_oracle_steps(h)
    finally:
        lib.crf_oracle_free(h)
    if want_arrays:
        return start, end, mlen, steps
    out = []
    for s, e, m in zip(start.tolist(), end.tolist(), mlen.tolist()):
        motif = bytes(raw[s:s + m]).decode("latin-1").upper()
        out.append((s, e, motif))
    return out


def _ptr_len(raw):
    if isinstance(raw, np.ndarray):
        assert raw.dtype == np.uint8 and raw.flags.c_contiguous
        return raw.ctypes.data, raw.size, raw
    buf = ctypes.create_string_buffer(raw, len(raw))
    return ctypes.addressof(buf), len(raw), buf


def detect_repeats(input_sequence, filter_settings, arrays=False):
    """Lock-step literal port; supports interval attrs and min_repeats == 1."""
    validate_filters(filter_settings)
    lib = _load()
    raw = _as_bytes(input_sequence)
    ptr, n, keep = _ptr_len(raw)
    has_iv = hasattr(filter_settings, "interval_start_0based") or hasattr(filter_settings, "interval_end")
    ivs = getattr(filter_settings, "interval_start_0based", 0)
    ive = getattr(filter_settings, "interval_end", n)
    if has_iv and (ivs < 0 or ive < 0):
        raise NotImplementedError("oracle: negative interval coordinates are not modelled")
    h = lib.crf_oracle_detect(ptr, n, filter_settings.min_motif_size, filter_settings.max_motif_size,
                              filter_settings.min_repeats, filter_settings.min_span, int(has_iv), ivs, ive)
    del keep
    return _collect(lib, h, raw, arrays)


def detect_repeats_by_k(input_sequence, filter_settings, threads=0, arrays=False):
    """Full-sequence mode, one thread per motif size (crf_oracle_detect_by_k)."""
    validate_filters(filter_settings)
    if hasattr(filter_settings, "interval_start_0based") or hasattr(filter_settings, "interval_end"):
        raise ValueError("detect_repeats_by_k is full-sequence mode only")
    lib = _load()
    raw = _as_bytes(input_sequence)
    ptr, n, keep = _ptr_len(raw)
    h = lib.crf_oracle_detect_by_k(ptr, n, filter_settings.min_motif_size, filter_settings.max_motif_size,
                                   filter_settings.min_repeats, filter_settings.min_span, int(threads))
    del keep
    return _collect(lib, h, raw, arrays)
