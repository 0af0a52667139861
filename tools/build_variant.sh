#!/bin/sh
# Experimental build of libghostcwt under another name: tools/build_variant.sh <tag> "<-D flags>"
# -> ghost_b200/variants/libghostcwt_<tag>.so, selected at run time with GCWT_LIB=<path>.
set -e
tag=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
bdir=$root/ghost_b200/variants/build_$tag
mkdir -p "$bdir"
cd "$root/ghost_b200/csrc"
for f in cabi generic_path fast_path sigtools host_path; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 $flags -c $f.cu -o "$bdir/$f.o" &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$root/ghost_b200/variants/libghostcwt_$tag.so" "$bdir"/*.o -lcudart
rm -rf "$bdir"
echo built "ghost_b200/variants/libghostcwt_$tag.so"
