#!/bin/bash
# tuning helper: build the forward TU (all its copies) with extra -D flags and link a side-by-side library
# usage: scripts/build_variant.sh <tag> <nvcc -D flags...>   ->  coordx_b200/_build/libcoordx_<tag>.so   (use with COORDX_B200_LIB)
set -e
TAG=$1; shift
cd "$(dirname "$0")/.."
B=coordx_b200/_build
OBJS=""
for k in 0 1 2 3; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
    "$@" -DCOORDX_FWD_COPY=$k -c -o $B/trunk_fwd_tc_${TAG}_$k.o coordx_b200/csrc/trunk_fwd_tc.cu &
  OBJS="$OBJS $B/trunk_fwd_tc_${TAG}_$k.o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $B/libcoordx_$TAG.so $B/coordx_abi.o $OBJS $B/trunk_bwd_tc.o $B/trunk_bwd_tc2.o
echo built $B/libcoordx_$TAG.so
