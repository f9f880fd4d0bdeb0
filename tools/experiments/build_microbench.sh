#!/bin/sh
# Diagnostics only (exchange / synchronisation micro-benchmarks used by tools/microbench*.py): NOT part of the product
# library.  Builds tools/experiments/libphx_microbench.so against the in-tree libphoenix_b200.so.
set -e
cd "$(dirname "$0")/../.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -shared -Xcompiler -fPIC \
     tools/experiments/phx_microbench.cu -o tools/experiments/libphx_microbench.so \
     -Lphoenix_b200 -lphoenix_b200 -Xlinker -rpath -Xlinker "$(pwd)/phoenix_b200"
