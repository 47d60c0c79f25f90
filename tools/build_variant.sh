#!/bin/bash
# tools/build_variant.sh NAME "-DRTM_KNR=8 ..."  -> rtm_gpu_b200/build/variants/librtm_NAME.so (experiments only)
set -e
name=$1; defs=$2
mkdir -p rtm_gpu_b200/build/variants /tmp/rtmv_$name
cd rtm_gpu_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-O2,-pthread $defs"
nvcc $F -c rtm_engine.cu -o /tmp/rtmv_$name/e.o 2>&1 | grep -v warning || true
objs=/tmp/rtmv_$name/e.o
for f in host_abi.cpp rtm_nccl.cpp driver.cpp host/fd_operator.cpp host/model.cpp host/config.cpp host/resample.cpp host/segy_io.cpp host/poststack.cpp; do
  o=/tmp/rtmv_$name/$(basename $f).o
  [ -f ../build/$(basename $f).o ] && objs="$objs ../build/$(basename $f).o" && continue
  nvcc $F -c $f -o $o; objs="$objs $o"
done
nvcc -shared -o ../build/variants/librtm_$name.so $objs -lcudart_static -ldl -lpthread -lrt 2>&1 | grep -v warning || true
echo built $name
