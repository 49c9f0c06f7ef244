#!/usr/bin/env bash
# Build libegm_b200.so for sm_100a (cross-compiles without a GPU).  Usage: build.sh [-j N]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libegm_b200.so"
OBJ="$HERE/../build"
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr)
pids=()
for f in "$HERE"/*.cu; do
  o="$OBJ/$(basename "${f%.cu}").o"
  if [[ ! -f "$o" || "$f" -nt "$o" || "$HERE/common.cuh" -nt "$o" || -n "$(find "$HERE" -name "*.cuh" -newer "$o" -print -quit)" || "$HERE/../../include/egm_b200.h" -nt "$o" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" > "$o.log" 2>&1 || { cat "$o.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$OBJ"/*.o -lcudart
echo "built $OUT"
