#!/bin/bash
# build_variant_cl.sh NAME "-DFLAG ..." : variant of the N = 65536 cluster kernel only (jade_k_pkcl.cu), linked with the regular objects
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../jadespectrogram_b200/csrc"
out=../../tools/bin/variants; mkdir -p $out/obj_$name
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v $flags -c jade_k_pkcl.cu -o $out/obj_$name/jade_k_pkcl.o 2> $out/obj_$name/pkcl.ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/lib_$name.so jade_gpu.o jade_k_pk.o jade_k_pk2.o jade_k_pkz.o jade_k_pksmall_a.o jade_k_pksmall_b.o \
  jade_k_pkcta.o jade_k_pk3.o $out/obj_$name/jade_k_pkcl.o jade_k_warp_a.o jade_k_warp_b.o jade_k_cta.o jade_host_tables.o jade_view.o jade_axis.o
grep -A2 "pkcl3_kernelILi0" $out/obj_$name/pkcl.ptxas.log | grep -E "registers|spill" | paste - - | cut -c1-160
