#!/bin/bash
# Build an experimental variant of librscape_b200.so with extra nvcc flags into r-scape_b200/build/alt/<name>.so
# (selected at run time with RSCAPE_B200_LIB=<path>).  Usage: tools/build_variant.sh e4 -DRSB_EPI_GROUPS=4
set -e
name=$1; shift
cd "$(dirname "$0")/../r-scape_b200"
mkdir -p build/alt/$name
for f in gram_tcgen05 pack stats correct_hist hits treesubs nullgen msaprep peer_reduce capi; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c csrc/$f.cu -o build/alt/$name/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/alt/$name.so build/alt/$name/*.o -lcudart -ldl
echo built build/alt/$name.so
