#!/bin/bash
# scratch/build_variant.sh NAME FILE.cu "FLAGS" -> mcos_b200/libmcosb200_NAME.so: the library with FILE.cu recompiled with FLAGS
# (A/B timing of compile-time variants: MCOSB_LIB_PATH=$PWD/mcos_b200/libmcosb200_NAME.so python ...)
set -e
cd "$(dirname "$0")/../mcos_b200/csrc"
name=$1; shift
mkdir -p /tmp/tb/$name
objs=""
for f in *.o; do objs="$objs $f"; done
while [ $# -ge 2 ]; do
  src=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c $src -o /tmp/tb/$name/${src%.cu}.o
  objs=$(echo $objs | sed "s# ${src%.cu}.o# /tmp/tb/$name/${src%.cu}.o#; s#^${src%.cu}.o#/tmp/tb/$name/${src%.cu}.o#")
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libmcosb200_$name.so $objs
echo built ../libmcosb200_$name.so
