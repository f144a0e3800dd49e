#!/bin/bash
# Build the UNMODIFIED reference (Fabian2598/SchwingerModel) as a shared library for one
# lattice size, from the sources where they lie under $SM_REFERENCE (default /root/reference).
# Nothing is copied: g++ reads the reference files in place; the only output is
#   oracle/_ref/libref_<NS>x<NT>.so          (git-ignored, travels to the GPU box)
# Flags follow the reference's CMakeLists.txt:30-31 (-O3, C++20, no -march).
#   usage: build_ref.sh NS NT [--exe]      (--exe also builds the reference's own main.cpp)
set -euo pipefail
NS=${1:?NS}; NT=${2:?NT}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SM_REFERENCE:-/root/reference}"
OUT="$HERE/../_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref: reference tree $REF not present (fine on the GPU box: prebuilt files are used)" >&2
  exit 3
fi
mkdir -p "$OUT"
LIB="$OUT/libref_${NS}x${NT}.so"
SRCS="$REF/src/statistics.cpp $REF/src/variables.cpp $REF/src/gauge_conf.cpp $REF/src/dirac_operator.cpp $REF/src/conjugate_gradient.cpp $REF/src/hmc.cpp"
FLAGS="-std=c++20 -O3 -w -DCONFIG_H -DNS=$NS -DNT=$NT -I$HERE/minimpi -I$REF/include"
if [ ! -f "$LIB" ] || [ "$HERE/ref_harness.cpp" -nt "$LIB" ] || [ "$HERE/minimpi/mpi.h" -nt "$LIB" ]; then
  g++ $FLAGS -fPIC -shared -o "$LIB" "$HERE/ref_harness.cpp" $SRCS
  echo "built $LIB"
fi
if [ "${3:-}" = "--exe" ]; then
  EXE="$OUT/SM_${NS}x${NT}"
  if [ ! -f "$EXE" ]; then
    g++ $FLAGS -o "$EXE" "$REF/src/main.cpp" $SRCS
    echo "built $EXE"
  fi
fi
