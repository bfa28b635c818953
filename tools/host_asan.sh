#!/bin/bash
# The host library's CPU tests (readers, index, DEFLATE decoder, matcher) under AddressSanitizer + UBSan.
# Builds an instrumented libnimpress_host.so in place, runs tests/test_host_cpu.py, restores the normal build.
set -e
cd "$(dirname "$0")/../nimpress_b200/host"
LIB="$(pwd)/../lib/libnimpress_host.so"
cp "$LIB" /tmp/libnimpress_host.orig.so
trap 'cp /tmp/libnimpress_host.orig.so "$LIB"' EXIT
g++ -std=c++17 -O1 -g -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -shared -o ../lib/libnimpress_host.so \
    inputs.cpp variant_source.cpp region_index.cpp fast_inflate.cpp stats.cpp driver.cpp host_api.cpp -L../lib -lnimpress_cuda -lz -lpthread -Wl,-rpath,'$ORIGIN'
cd ../..
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 \
    python -m pytest tests/test_host_cpu.py -x -q -p no:cacheprovider
