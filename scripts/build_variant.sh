#!/bin/bash
# Builds an A/B variant of the trace library: scripts/build_variant.sh NAME "-DB200RT_TRAV_THRESHOLD=8 ..."
# -> ipu_ray_lib_b200/variants/libb200rt_NAME.so (same nif.o as the product build); compare with scripts/gpu_ab2.sh.
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
FLAGS="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ --fmad=false -prec-div=true -prec-sqrt=true -ftz=false -diag-suppress 549"
mkdir -p ipu_ray_lib_b200/variants /tmp/b200rt_variants
/usr/local/cuda/bin/nvcc $FLAGS "$@" -c -o /tmp/b200rt_variants/b200rt_$NAME.o ipu_ray_lib_b200/csrc/b200rt.cu
[ -f ipu_ray_lib_b200/csrc/nif.o ] || make ipu_ray_lib_b200/csrc/nif.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o ipu_ray_lib_b200/variants/libb200rt_$NAME.so /tmp/b200rt_variants/b200rt_$NAME.o ipu_ray_lib_b200/csrc/nif.o -cudart static
echo built $NAME
