#!/bin/bash
# Build libschwinger_b200.so for sm_100a only (in-tree; the .so is git-ignored but travels to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libschwinger_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
if [ "${1:-}" != "--force" ] && [ -f "$OUT" ]; then
  newer=$(find "$HERE" "$HERE/../../include" -newer "$OUT" \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' -o -name 'build.sh' \) | head -1)
  if [ -z "$newer" ]; then echo "up to date: $OUT"; exit 0; fi
fi
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --shared -o "$OUT" "$HERE/sm_abi.cu" -ldl 2> "$HERE/ptxas.log" \
  || { cat "$HERE/ptxas.log" >&2; exit 1; }
echo "built $OUT"
