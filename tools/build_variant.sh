#!/bin/bash
# Builds zdr_b200/csrc/variants/lib_<name>.so = the library with the files of FILE (default "render"; e.g. FILE="render persistent",
# FILE="render_fast persistent_fast") recompiled with extra -D flags (kernel-tuning experiments; select at run time with
# ZDR_B200_LIB=<path>).  FMAD=true builds the exact files with floating-point contraction; FAST_FLAGS replaces --use_fast_math for
# the *_fast files (e.g. FAST_FLAGS="-fmad=true -prec-div=false -prec-sqrt=false": no flush to zero).
# usage: [FILE="render persistent"] [FMAD=true] build_variant.sh name -DX=1 ...
set -e
cd "$(dirname "$0")/../zdr_b200/csrc"
name=$1; shift
mkdir -p variants
make -s -j4 > /dev/null
objs=""
for o in api render wavefront persistent render_fast wavefront_fast persistent_fast bvh_build envmap obj_loader; do
  if [[ " ${FILE:=render} " == *" $o "* ]]; then
    math="-fmad=${FMAD:-false}"
    [[ $o == *_fast ]] && math="${FAST_FLAGS:---use_fast_math}"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo $math -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off \
         --expt-relaxed-constexpr "$@" -c -o variants/${o}_$name.o $o.cu &
    objs="$objs variants/${o}_$name.o"
  else objs="$objs $o.o"; fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_$name.so $objs -Xcompiler -fPIC
rm -f variants/*_$name.o
echo variants/lib_$name.so
