"""CPU-side checks of the boundary: the library builds for sm_100a, loads, and exports every
symbol include/crf.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "colab-repeat-finder_b200"))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "crf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crf_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from crf_b200 import _cabi, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert declared, "no declarations found in include/crf.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in crf.h but not exported"
    assert sorted(_cabi.EXPORTS) == declared
    lib.crf_abi_version.restype = ctypes.c_int
    assert lib.crf_abi_version() == 1


def test_struct_layouts_match_header():
    from crf_b200 import _cabi
    assert ctypes.sizeof(_cabi.ScanParams) == 9 * 4
    assert ctypes.sizeof(_cabi.SeqInfo) == 5 * 8 + 2 * 4 + 8
    assert ctypes.sizeof(_cabi.ScanStats) == 2 * 8 + 6 * 8 + 2 * 4 + 8


def test_sass_is_sm100a_only():
    import subprocess
    from crf_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "colab-repeat-finder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"crf_oracle", r"libcrf_oracle", r"oracle\.py"):
                    assert not re.search(pat, text, flags=re.M), f"{f} reaches into oracle/ ({pat})"
