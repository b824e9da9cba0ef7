#!/usr/bin/env bash
# Build recipe for oracle/_ref: the reference's ONLY native component, compiled from the
# UNMODIFIED source where it lies (/root/reference/gpdemo/kernels.pyx).
#
# The shipped, pre-generated gpdemo/kernels.c (Cython 0.22) does not compile against CPython 3.12
# (SURVEY.md §8c), so the .pyx is re-cythonized.  Nothing from /root/reference is copied into the
# repository: the intermediate C file and the shared object live only in oracle/_ref/ (git-ignored,
# NOT gpurun-ignored, so the built module travels to the GPU box like our own .so files).
#
# TEST INFRASTRUCTURE ONLY.  Used by tests/, bench.py's cpu_baseline leg and oracle/gen_golden.py.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${APM_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
PY="${PYTHON:-python}"
if [ ! -f "$REF/gpdemo/kernels.pyx" ]; then
  echo "oracle/build_ref.sh: $REF/gpdemo/kernels.pyx not present; keeping any prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
EXT="$($PY -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")"
INC="$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")"
# cython reads the .pyx in place and writes the generated C into oracle/_ref only
$PY -m cython -3 "$REF/gpdemo/kernels.pyx" -o "$OUT/kernels.c"
# default reference build has no -ffast-math (setup.py:58 only adds it with -use-gcc-opts)
gcc -O2 -fPIC -shared -I"$INC" "$OUT/kernels.c" -o "$OUT/kernels$EXT"
rm -f "$OUT/kernels.c"
echo "built $OUT/kernels$EXT"
