#!/usr/bin/env bash
# Thin / non-DoubleConv tcgen05 conv shapes of cfg2 (GRFB branches, FusionConv, edge enhancers, RGA): per-kernel times, fwd + wgrad.
cd "$(dirname "${BASH_SOURCE[0]}")/.."
echo "== 1x1"; python tools/conv_bench.py --k 1 --shapes 16,16,240 64,64,240 64,16,240 16,64,240 16,112,240 32,32,120 128,128,120 256,256,60 256,256,30
echo "== 3x3"; python tools/conv_bench.py --k 3 --shapes 16,16,240 64,16,240 16,32,480 32,32,120 64,64,120 64,64,60
echo "== 3x3 dil 12"; python tools/conv_bench.py --k 3 --dil 12 --shapes 16,16,240 32,32,120
echo "== 7x7"; python tools/conv_bench.py --k 7 --shapes 16,16,240 32,32,120 64,64,60
