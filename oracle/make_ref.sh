#!/usr/bin/env bash
# make_ref.sh -- put the UNMODIFIED Python reference where the GPU box can run it.
#
# TEST INFRASTRUCTURE ONLY.  The reference (broadinstitute/colab-repeat-finder) is pure Python: there is
# nothing to compile, so its "build" is a copy of the three things the hot path consists of
#   perfect_repeat_finder.py, utils/ (tracker + helpers), perfect_repeat_finder_tests.py
# from /root/reference into oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the box like a built
# .so; no reference source ever enters the repo history), plus two stub packages for the imports the
# reference makes at module top but the scan never touches (pyfastx, matplotlib -- neither is installed and
# there is no network).  oracle/ref.py imports the copy; tests/ and bench.py's reference legs use it as the
# L0 truth and as the timed CPU baseline ("kind": "reference").
#
#   usage: oracle/make_ref.sh [reference_dir]      (default /root/reference)
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref"
if [ ! -f "$SRC/perfect_repeat_finder.py" ]; then
    echo "make_ref.sh: $SRC/perfect_repeat_finder.py not found (the GPU box uses the prebuilt oracle/_ref)" >&2
    exit 3
fi
rm -rf "$DST"
mkdir -p "$DST/utils" "$DST/_stubs/pyfastx" "$DST/_stubs/matplotlib"
cp "$SRC/perfect_repeat_finder.py" "$SRC/perfect_repeat_finder_tests.py" "$DST/"
cp "$SRC"/utils/*.py "$DST/utils/"
# stubs: import-time placeholders only; calling into them fails loudly
cat > "$DST/_stubs/pyfastx/__init__.py" <<'EOF'
"""Stub: pyfastx is not installed; the reference only uses it in main() (FASTA parsing), never in detect_repeats()."""
def __getattr__(name):
    raise ImportError("pyfastx stub (oracle/_ref/_stubs): FASTA parsing of the reference is not available here")
EOF
cat > "$DST/_stubs/matplotlib/__init__.py" <<'EOF'
"""Stub: matplotlib is not installed; the reference imports it for plotting only."""
EOF
cat > "$DST/_stubs/matplotlib/pyplot.py" <<'EOF'
def __getattr__(name):
    raise ImportError("matplotlib stub (oracle/_ref/_stubs): plotting is not available here")
EOF
cat > "$DST/_stubs/matplotlib/colors.py" <<'EOF'
class ListedColormap:                      # only referenced inside plotting functions
    def __init__(self, *a, **k):
        raise ImportError("matplotlib stub (oracle/_ref/_stubs): plotting is not available here")
EOF
( cd "$SRC" && sha256sum perfect_repeat_finder.py perfect_repeat_finder_tests.py utils/*.py ) > "$DST/SHA256SUMS"
echo "oracle/_ref: reference copied from $SRC ($(wc -l < "$DST/SHA256SUMS") files)"
