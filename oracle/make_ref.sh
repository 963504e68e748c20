#!/bin/bash
# Installs the UNMODIFIED reference (A-CGray/FEMpy, pure Python + Numba) into oracle/_ref with pip, so that the timed CPU
# baseline of bench.py (`cpu_baseline`, `--impl reference`) and one GPU test can run the reference's own
# FEMpyProblem._assembleMatrix / _assembleResidual on the GPU box, where /root/reference does not exist.
# oracle/_ref is git-ignored (no reference source enters the history) but travels with the gpurun snapshot.
# TEST / BENCH INFRASTRUCTURE: nothing under fempy_b200/ imports it.  The reference's un-installed dependencies
# (mdolab-baseclasses, meshio, parameterized) are replaced by the stubs in oracle/stubs.
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${FEMPY_REFERENCE_ROOT:-/root/reference}"
if [ ! -d "$REF/FEMpy" ]; then
    echo "make_ref: no reference at $REF -- keeping whatever oracle/_ref holds"
    exit 0
fi
if [ -f "$HERE/_ref/FEMpy/__init__.py" ] && [ "$1" != "--force" ]; then
    exit 0
fi
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/src"   # the reference tree is read-only and setuptools writes its egg-info next to setup.py
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --no-compile --target "$HERE/_ref" "$TMP/src"
test -f "$HERE/_ref/FEMpy/Problem.py"
echo "make_ref: installed FEMpy $(python -c "import re;print(re.findall(r'__version__ = .([0-9.]*)', open('$HERE/_ref/FEMpy/__init__.py').read())[0])") into oracle/_ref"
