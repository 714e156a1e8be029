#!/bin/bash
# Timing-experiment builds of the library (wrong results by design): build/lib_dbg<bits>.so with -DMCEIK_DBG=<bits>
# (see fsm_bricks16.cu).  Usage: tools/build_variants.sh 1 2 3 ...
set -e
cd "$(dirname "$0")/../mceik_b200/csrc"
mkdir -p ../../build
for v in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
    -Xcompiler -fPIC,-ffp-contract=off -cudart static -DMCEIK_DBG=$v -shared -o ../../build/lib_dbg$v.so \
    abi.cu comm.cu fsm.cu fsm_bricks.cu fsm_bricks16.cu gs.cu -ldl &
done
wait
ls -la ../../build/lib_dbg*.so
