#!/bin/bash
# Builds zdr_b200/csrc/variants/lib_<name>.so = the library with the files of FILE (default "render"; e.g. FILE="render persistent")
# recompiled with extra -D flags (kernel-tuning experiments; select at run time with ZDR_B200_LIB=<path>).  FMAD=true builds
# those files with floating-point contraction (the "fma" performance build).
# usage: [FILE="render persistent"] [FMAD=true] build_variant.sh name -DX=1 ...
set -e
cd "$(dirname "$0")/../zdr_b200/csrc"
name=$1; shift
mkdir -p variants
make -s -j4 > /dev/null
objs=""
for o in api render persistent bvh_build envmap obj_loader; do
  if [[ " ${FILE:=render} " == *" $o "* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=${FMAD:-false} -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off \
         --expt-relaxed-constexpr "$@" -c -o variants/${o}_$name.o $o.cu &
    objs="$objs variants/${o}_$name.o"
  else objs="$objs $o.o"; fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_$name.so $objs -Xcompiler -fPIC
rm -f variants/*_$name.o
echo variants/lib_$name.so
