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


def test_ctypes_structs_have_the_layout_a_c_compiler_gives_the_header(tmp_path):
    """include/crf.h compiled as C99 (-pedantic: the boundary is plain C); sizeof and every field offset of its four structs,
    as gcc lays them out, against the ctypes mirrors in _cabi."""
    import shutil
    import subprocess
    import pytest
    from crf_b200 import _cabi
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    mirrors = {"crf_scan_params": _cabi.ScanParams, "crf_seq_info_t": _cabi.SeqInfo, "crf_scan_stats_t": _cabi.ScanStats,
               "crf_xchg_result_t": _cabi.XchgResult}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "crf.h"', "int main(void) {"]
    for c_name, mirror in mirrors.items():
        lines.append(f'    printf("{c_name} sizeof %zu\\n", sizeof({c_name}));')
        for field, _ in mirror._fields_:
            lines.append(f'    printf("{c_name} {field} %zu\\n", offsetof({c_name}, {field}));')
    lines += ["    return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines) + "\n")
    exe = str(tmp_path / "probe")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe])
    seen = 0
    for line in subprocess.check_output([exe], text=True).splitlines():
        c_name, what, value = line.split()
        mirror = mirrors[c_name]
        assert int(value) == (ctypes.sizeof(mirror) if what == "sizeof" else getattr(mirror, what).offset), line
        seen += 1
    assert seen == sum(1 + len(m._fields_) for m in mirrors.values())
    assert _cabi.XCHG_MAX_WORLD == int(re.search(r"#define CRF_XCHG_MAX_WORLD (\d+)", open(os.path.join(ROOT, "include", "crf.h")).read()).group(1))


def test_c_example_builds_against_the_header_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/fasta_to_bed.c -- the whole path through the C ABI from plain C99 -- compiles with -pedantic -Werror, links
    against libcrf.so and runs: the host-only half (FASTA reader, packer) works anywhere; without a GPU the context call must
    return a status and a message (exit 1), with one the BED file must be there."""
    import shutil
    import subprocess
    import pytest
    from crf_b200 import build
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    lib_dir = os.path.dirname(build.build())
    exe = str(tmp_path / "fasta_to_bed")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "fasta_to_bed.c"), "-o", exe, "-L", lib_dir, "-l:libcrf.so",
                           "-Wl,-rpath," + lib_dir])
    fa = tmp_path / "tiny.fa"
    fa.write_text(">a desc\nACGTACGTACGTACGTACGTNNNN\nacacacacacacacacac\n>b\nTTTTTTTTTTTTTTTTTTTT\n")
    bed = tmp_path / "tiny.bed"
    run = subprocess.run([exe, str(fa), str(bed)], capture_output=True, text=True)
    assert "2 records, 62 bp" in run.stdout
    if run.returncode == 0:                              # a GPU is present
        assert bed.read_text() == "a\t0\t20\tACGT\na\t24\t42\tAC\nb\t0\t20\tT\n"
    else:
        assert run.returncode == 1 and "crf_ctx_create" in run.stderr and "status" in run.stderr
    assert subprocess.run([exe], capture_output=True).returncode == 2


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
