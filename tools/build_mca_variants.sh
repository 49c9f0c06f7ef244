#!/usr/bin/env bash
# Tuning aid: build libegm_b200 variants that differ only in the MCALayer row-walk kernel configuration (csrc/mca_fused.cu macros).
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="$ROOT/egm-unet_b200/csrc/mca_fused.cu"; OBJ="$ROOT/egm-unet_b200/build"; OUT="$ROOT/egm-unet_b200/variants"
NVCC=/usr/local/cuda/bin/nvcc
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr)
others=$(ls "$OBJ"/*.o | grep -v mca_fused.o)
build() {  # name, defines...
  local name="$1"; shift
  "$NVCC" "${FLAGS[@]}" "$@" -c "$SRC" -o "/tmp/mcav_$name.o"
  "$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libegm_$name.so" $others "/tmp/mcav_$name.o" -lcudart
  echo "built $name"
}
build A &
build B -DMF_MINB_FWD=2 -DMF_MINB_BWD=2 &
build C -DMF_CC=32 -DMF_MINB_FWD=4 -DMF_MINB_BWD=3 &
build D -DMF_CC=32 -DMF_MINB_FWD=3 -DMF_MINB_BWD=2 &
build E -DMF_CC=32 -DMF_MINB_FWD=2 -DMF_MINB_BWD=2 &
wait
