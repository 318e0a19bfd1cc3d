#!/bin/bash
# builds coordx_b200/_build/libcoordx_trace.so: the library with the pipeline-trace stamps compiled in (-DCOORDX_TRACE=1).
# The trace scripts (scripts/fwd_trace.py, dw2_trace.py, dfeat_trace.py, bwd_trace.py) need it:  COORDX_B200_LIB=.../libcoordx_trace.so
set -e
cd "$(dirname "$0")/.."
B=coordx_b200/_build
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DCOORDX_TRACE=1"
mkdir -p $B/trace
nvcc $F -c -o $B/trace/coordx_abi.o coordx_b200/csrc/coordx_abi.cu &
nvcc $F -c -o $B/trace/trunk_bwd_tc.o coordx_b200/csrc/trunk_bwd_tc.cu &
nvcc $F -c -o $B/trace/trunk_bwd_tc2.o coordx_b200/csrc/trunk_bwd_tc2.cu &
for k in 0 1 2 3; do nvcc $F -DCOORDX_FWD_COPY=$k -c -o $B/trace/trunk_fwd_tc_$k.o coordx_b200/csrc/trunk_fwd_tc.cu & done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $B/libcoordx_trace.so $B/trace/*.o
echo built $B/libcoordx_trace.so
