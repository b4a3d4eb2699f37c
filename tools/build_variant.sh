#!/bin/bash
# build_variant.sh NAME "-DFLAG ..." : kernel-variant experiments -- compiles jade_k_pk.cu + jade_gpu.cu with extra flags and
# links them with the regular objects into tools/bin/variants/lib_NAME.so (load with JADE_GPU_LIB=...).
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../jadespectrogram_b200/csrc"
out=../../tools/bin/variants; mkdir -p $out/obj_$name
NV="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v $flags"
$NV -c jade_k_pk.cu -o $out/obj_$name/jade_k_pk.o 2> $out/obj_$name/pk.ptxas.log &
$NV -c jade_gpu.cu -o $out/obj_$name/jade_gpu.o 2> $out/obj_$name/gpu.ptxas.log &
$NV -c jade_k_pk2.cu -o $out/obj_$name/jade_k_pk2.o 2> $out/obj_$name/pk2.ptxas.log &
$NV -c jade_k_pkz.cu -o $out/obj_$name/jade_k_pkz.o 2> $out/obj_$name/pkz.ptxas.log &
if [ -n "$3" ]; then $NV -c jade_k_pkcta.cu -o $out/obj_$name/jade_k_pkcta.o 2> $out/obj_$name/pkcta.ptxas.log & fi
wait
PKCTA=jade_k_pkcta.o; [ -n "$3" ] && PKCTA=$out/obj_$name/jade_k_pkcta.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/lib_$name.so $out/obj_$name/jade_gpu.o $out/obj_$name/jade_k_pk.o $out/obj_$name/jade_k_pk2.o $out/obj_$name/jade_k_pkz.o \
  jade_k_pksmall_a.o jade_k_pksmall_b.o $PKCTA jade_k_pk3.o jade_k_pkcl.o jade_k_warp_a.o jade_k_warp_b.o jade_k_cta.o jade_host_tables.o jade_view.o jade_axis.o
grep -A2 "pkz2048_kernelILb0ELi0" $out/obj_$name/pkz.ptxas.log | grep -E "registers|spill" | paste - - | sed 's/ptxas info    ://g' | cut -c1-200
grep -A2 "pk2048_kernelILi1ELb0ELi0\|pk2048_kernelILi0ELb0ELi0\|pk2048_kernelILi1ELb0ELi1" $out/obj_$name/pk.ptxas.log | grep -E "registers|spill" | paste - - | sed 's/ptxas info    ://g' | cut -c1-200
