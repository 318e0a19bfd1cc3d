#!/bin/bash
# tuning helper: build the forward TU with a different MUFU/polynomial split and link a side-by-side library
# usage: scripts/build_variant.sh <mufu_cols>   ->  coordx_b200/_build/libcoordx_nm<cols>.so
set -e
NM=$1
cd "$(dirname "$0")/.."
B=coordx_b200/_build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
  -DCOORDX_FWD_MUFU_COLS=$NM -c -o $B/trunk_fwd_tc_nm$NM.o coordx_b200/csrc/trunk_fwd_tc.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $B/libcoordx_nm$NM.so $B/coordx_abi.o $B/trunk_fwd_tc_nm$NM.o $B/trunk_bwd_tc.o
echo built $B/libcoordx_nm$NM.so
