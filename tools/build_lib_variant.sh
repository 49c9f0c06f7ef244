#!/usr/bin/env bash
# Tuning aid: build a complete variant of libegm_b200 with extra -D flags into egm-unet_b200/variants/libegm_<name>.so
# (select it at run time with EGM_LIB=<path>).  Usage: tools/build_lib_variant.sh <name> [-DMACRO=V ...]
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
name="$1"; shift
SRC="$ROOT/egm-unet_b200/csrc"; OBJ="/tmp/egm_variant_$name"; OUT="$ROOT/egm-unet_b200/variants"
mkdir -p "$OBJ" "$OUT"
NVCC=/usr/local/cuda/bin/nvcc
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr)
for f in "$SRC"/*.cu; do
  ( "$NVCC" "${FLAGS[@]}" "$@" -c "$f" -o "$OBJ/$(basename "${f%.cu}").o" ) &
done
wait
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libegm_$name.so" "$OBJ"/*.o -lcudart
echo "built $OUT/libegm_$name.so"
