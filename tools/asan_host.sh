#!/bin/bash
# tools/asan_host.sh -- the host C++ of the library under AddressSanitizer + UBSan: builds a variant
# (/tmp/rtm_asan/librtm_asan.so: instrumented host objects + the normal device object) and runs the CPU tests
# that exercise the host code through the C ABI.  Needs a normal build first (python -m rtm_gpu_b200.build).
set -e
out=/tmp/rtm_asan; mkdir -p $out
cd "$(dirname "$0")/../rtm_gpu_b200/csrc"
objs=""
for f in host_abi.cpp rtm_nccl.cpp driver.cpp host/fd_operator.cpp host/model.cpp host/config.cpp host/resample.cpp host/segy_io.cpp host/poststack.cpp; do
  o=$out/$(basename $f).o
  g++ -std=c++17 -O1 -g -fPIC -ffp-contract=off -pthread -fsanitize=address,undefined -fno-omit-frame-pointer \
      -I../../include -I. -Ihost -I/usr/local/cuda/include -c $f -o $o
  objs="$objs $o"
done
nvcc -shared -o $out/librtm_asan.so $objs ../build/rtm_engine.cu.o -lcudart_static -ldl -lpthread -lrt -Xlinker -lasan -Xlinker -lubsan 2>&1 | grep -v deprecated || true
cd ../..
RTM_LIB_PATH=$out/librtm_asan.so LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
  ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 \
  python -m pytest tests/test_host.py tests/test_driver_frontend.py -q -s 2>&1 | tee $out/log.txt | tail -3
echo "sanitizer reports: $(grep -ci 'runtime error\|AddressSanitizer' $out/log.txt)"
