#!/bin/bash
# The kernel simulation (tests/sim/jpsim.cpp: the kernels' algorithm from jpgpu_core.h + the host planner) built with
# AddressSanitizer, and the simulation parity tests run on it - corrupted, flooded and truncated scans included.  No GPU.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/asan
g++ -O1 -g -fPIC -shared -std=c++17 -fsanitize=address -fno-omit-frame-pointer -I/usr/local/cuda/include \
    -o build/asan/libjpsim.so tests/sim/jpsim.cpp jpeg_rust_b200/csrc/jpgpu_host.cpp
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 JPSIM_LIB=$PWD/build/asan/libjpsim.so \
    python -m pytest tests/test_sim_parity.py -x -q "$@"
