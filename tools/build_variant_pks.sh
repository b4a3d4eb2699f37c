#!/bin/bash
# build_variant_pks.sh NAME "-DFLAG ..." : variant of the N <= 1024 packed kernels (jade_k_pksmall_*.cu + jade_gpu.cu)
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../jadespectrogram_b200/csrc"
out=../../tools/bin/variants; mkdir -p $out/obj_$name
NV="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v $flags"
$NV -c jade_k_pksmall_a.cu -o $out/obj_$name/jade_k_pksmall_a.o 2> $out/obj_$name/pksa.ptxas.log &
$NV -c jade_k_pksmall_b.cu -o $out/obj_$name/jade_k_pksmall_b.o 2> $out/obj_$name/pksb.ptxas.log &
$NV -c jade_gpu.cu -o $out/obj_$name/jade_gpu.o 2> $out/obj_$name/gpu.ptxas.log &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/lib_$name.so $out/obj_$name/jade_gpu.o jade_k_pk.o jade_k_pk2.o \
  $out/obj_$name/jade_k_pksmall_a.o $out/obj_$name/jade_k_pksmall_b.o jade_k_pkz.o jade_k_pkcta.o jade_k_pk3.o jade_k_pkcl.o jade_k_warp_a.o jade_k_warp_b.o jade_k_cta.o jade_host_tables.o jade_view.o jade_axis.o
grep -A2 "stft_pksmall_kernelILi16ELi0ELb0ELb0" $out/obj_$name/pksb.ptxas.log | grep -E "registers|spill" | paste - - | sed 's/ptxas info    ://g' | cut -c1-200
