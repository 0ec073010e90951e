#!/bin/bash
# usage: tools/build_variant.sh NAME [nvcc -D flags...]  ->  variants/libmpvss_NAME.so  (tuning builds; git-ignored,
# selected at run time with MPVSS_B200_LIB=variants/libmpvss_NAME.so)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p variants/obj_$name
C=mpvss_rs_b200/csrc
for f in api comm modp_api modp ec_api ec hash; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-pthread "$@" -c $C/$f.cu -o variants/obj_$name/$f.o &
done
wait
nvcc -shared -o variants/libmpvss_$name.so variants/obj_$name/*.o $C/sha256_ni.o -Xcompiler -pthread -ldl
echo built variants/libmpvss_$name.so
