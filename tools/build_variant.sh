#!/bin/bash
# Kernel tuning helper: builds ffmpeg-ht_b200/variants/lib_<name>.so with extra -D flags for one .cu file.
#   tools/build_variant.sh <name> <file.cu> "<-D flags>"      then run with B200J2K_LIB=ffmpeg-ht_b200/variants/lib_<name>.so
set -e
cd "$(dirname "$0")/../ffmpeg-ht_b200"
name=$1; src=$2; flags=$3
mkdir -p variants build
NV="nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Icsrc -I../include"
objs=""
for f in csrc/*.cu; do
  b=$(basename $f .cu)
  if [ "$b" == "$(basename $src .cu)" ]; then
    $NV $flags -c $f -o variants/${b}_$name.o; objs="$objs variants/${b}_$name.o"
  else
    objs="$objs build/$b.o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_$name.so $objs build/j2k_parse.o build/j2k_splitter.o
rm -f variants/*_$name.o
echo variants/lib_$name.so
