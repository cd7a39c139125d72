"""crf_b200 -- B200-native perfect-tandem-repeat scan behind the reference's detect_repeats() API.

Layout: csrc/ (CUDA kernels + C ABI, built into crf_b200/libcrf.so), _cabi (ctypes binding),
api (detect_repeats), fasta / cli (the reference's command line), partition (multi-GPU).
"""
from .api import detect_repeats, get_context, scan_arrays, validate_filter_settings  # noqa: F401

__all__ = ["detect_repeats", "get_context", "scan_arrays", "validate_filter_settings"]
