#!/bin/bash
# A/B build of the SAME library with extra -D flags: scripts/build_variant.sh <tag> [-D...]; -> vgsim_b200/libvgsim_b200_<tag>.so
# (select it with VGSIM_B200_LIB=...; the default library is built by vgsim_b200/build.py)
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
out=/tmp/vgsim_variant_$tag; mkdir -p $out
pids=()
for f in capi tau_kernel prep_kernels direct_kernel genealogy_kernel curves_kernel archive_kernel test_taps; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c vgsim_b200/csrc/$f.cu -o $out/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o vgsim_b200/libvgsim_b200_$tag.so $out/*.o -lcudart_static -ldl -lrt -lpthread
echo vgsim_b200/libvgsim_b200_$tag.so
