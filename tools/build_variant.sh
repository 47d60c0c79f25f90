#!/bin/bash
# A kernel-variant build of the library for experiments / checking builds:
#   tools/build_variant.sh NAME -DRTM_STRM_BARSYNC=1 ...   ->  rtm_gpu_b200/variants/librtm_b200_NAME.so
# (select it with RTM_LIB_PATH=rtm_gpu_b200/variants/librtm_b200_NAME.so; .so files are git-ignored but travel with gpurun)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p rtm_gpu_b200/variants /tmp/rtm_variant_$name
objs=""
for s in rtm_gpu_b200/csrc/*.cu rtm_gpu_b200/csrc/*.cpp rtm_gpu_b200/csrc/host/*.cpp; do
  case $s in */main.cpp) continue;; esac
  o=/tmp/rtm_variant_$name/$(basename $s).o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-O2,-pthread --threads 4 -Iinclude "$@" -c $s -o $o
  objs="$objs $o"
done
nvcc -shared -o rtm_gpu_b200/variants/librtm_b200_$name.so $objs -lcudart_static -ldl -lpthread -lrt
ls -la rtm_gpu_b200/variants/librtm_b200_$name.so
