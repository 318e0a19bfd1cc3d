#!/bin/bash
# tuning helper: build trunk_bwd_tc2.cu with extra -D flags and link a side-by-side library
# usage: scripts/build_variant_bwd2.sh <tag> <nvcc -D flags...>   ->  coordx_b200/_build/libcoordx_<tag>.so   (use with COORDX_B200_LIB)
set -e
TAG=$1; shift
cd "$(dirname "$0")/.."
B=coordx_b200/_build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
  "$@" -c -o $B/trunk_bwd_tc2_$TAG.o coordx_b200/csrc/trunk_bwd_tc2.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $B/libcoordx_$TAG.so $B/coordx_abi.o $B/trunk_fwd_tc.o $B/trunk_fwd_tc_copy1.o $B/trunk_fwd_tc_copy2.o $B/trunk_fwd_tc_copy3.o $B/trunk_bwd_tc.o $B/trunk_bwd_tc2_$TAG.o
echo built $B/libcoordx_$TAG.so
