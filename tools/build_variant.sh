#!/bin/bash
# Builds zdr_b200/csrc/variants/lib_<name>.so = the library with render.cu (or FILE=<other>.cu) compiled with extra
# -D flags (kernel-tuning experiments; select at run time with ZDR_B200_LIB=<path>).
# usage: [FILE=bvh_build] build_variant.sh name -DX=1 ...
set -e
cd "$(dirname "$0")/../zdr_b200/csrc"
name=$1; shift
mkdir -p variants
make -s -j4 > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off \
     --expt-relaxed-constexpr "$@" -c -o variants/${FILE:=render}_$name.o $FILE.cu
objs=""
for o in api render persistent bvh_build envmap obj_loader; do
  if [ "$o" == "$FILE" ]; then objs="$objs variants/${FILE}_$name.o"; else objs="$objs $o.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_$name.so $objs -Xcompiler -fPIC
rm -f variants/${FILE}_$name.o
echo variants/lib_$name.so
