#!/bin/bash
# A/B builds of single kernels: tools/build_variant.sh TAG file.cu -DMACRO=... [more flags]
# compiles fsgm_b200/csrc/<file.cu> with the extra flags into fsgm_b200/build_TAG/ and links fsgm_b200/libfsgm_TAG.so with the
# other objects of the default build (run fsgm_b200/build.py first).  Select at run time with FSGM_LIB=fsgm_b200/libfsgm_TAG.so.
set -e
TAG=$1; SRC=$2; shift 2
HERE=$(cd "$(dirname "$0")/.." && pwd)/fsgm_b200
mkdir -p $HERE/build_$TAG
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O2,-ffp-contract=off,-fvisibility=hidden --fmad=false \
     "$@" -c $HERE/csrc/$SRC -o $HERE/build_$TAG/$SRC.o
OBJS=$(ls $HERE/build/*.o | grep -v "/$SRC.o")
nvcc -shared -o $HERE/libfsgm_$TAG.so $OBJS $HERE/build_$TAG/$SRC.o -Xcompiler -fPIC
echo $HERE/libfsgm_$TAG.so
