#!/bin/bash
# profiles/build_variant.sh <name> [extra nvcc flags]: build csrc/ into exp/libipx_<name>.so for A/B runs (IPX_LIB=exp/libipx_<name>.so selects it)
name=$1; shift; mkdir -p /root/repo/exp
cd /root/repo/interpax_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared --threads 6 "$@" -o /root/repo/exp/libipx_$name.so ipx_eval.cu ipx_slopes.cu ipx_prep.cu ipx_sort.cu ipx_vjp.cu ipx_ppoly.cu
