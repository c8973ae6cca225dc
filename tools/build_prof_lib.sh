#!/bin/sh
# Builds gmeta_b200/libgmeta_b200_prof.so: the library with the CTA-pair kernel's per-role cycle counters
# compiled in (GMETA_PAIR_PROF=1).  Use with GMETA_B200_LIB=... tools/layer_bench.py --impls 3 --profile.
set -e
cd "$(dirname "$0")/../gmeta_b200/csrc"
make >/dev/null
mkdir -p build/prof
for f in build/*.o; do cp "$f" build/prof/; done
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DGMETA_PAIR_PROF=1 \
  -c gcn_layer_pair.cu -o build/prof/gcn_layer_pair.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libgmeta_b200_prof.so build/prof/*.o -cudart static -Xcompiler -pthread
