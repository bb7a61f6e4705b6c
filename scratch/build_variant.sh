#!/bin/bash
# usage: build_variant.sh <out.so> <extra nvcc -D flags...>   (A/B builds of libfsgm.so for kernel experiments)
out=$1; shift
tmp=$(mktemp -d)
cd /root/repo/fsgm_b200/csrc
for f in *.cu; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O2,-ffp-contract=off,-fvisibility=hidden --fmad=false "$@" -c $f -o $tmp/$f.o > $tmp/$f.log 2>&1 &
done
wait
grep -l "error" $tmp/*.log | xargs -r cat
nvcc -shared -o $out $tmp/*.o -Xcompiler -fPIC 2> /dev/null
ls -la $out
rm -rf $tmp
