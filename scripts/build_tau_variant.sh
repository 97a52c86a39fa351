#!/bin/bash
# A/B build that recompiles only tau_kernel.cu with extra -D flags and links the other objects of the default build:
# scripts/build_tau_variant.sh <tag> [-D...] -> vgsim_b200/libvgsim_b200_<tag>.so  (select with VGSIM_B200_LIB=...)
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
out=/tmp/vgsim_tauvar_$tag; mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c vgsim_b200/csrc/tau_kernel.cu -o $out/tau_kernel.o 2>&1 | grep -v "offline compilation\|warning\|\^\|Remark\|const int K\|^$" || true
objs=""
for f in capi prep_kernels direct_kernel genealogy_kernel curves_kernel archive_kernel test_taps; do objs="$objs vgsim_b200/csrc/$f.o"; done
nvcc -shared -o vgsim_b200/libvgsim_b200_$tag.so $out/tau_kernel.o $objs -lcudart_static -ldl -lrt -lpthread 2>&1 | grep -v "offline compilation" || true
echo vgsim_b200/libvgsim_b200_$tag.so
