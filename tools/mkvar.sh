#!/bin/bash
# Builds a variant of libjpgpu.so with extra compile-time switches into build/variants/<name>.so (git-ignored, but it
# travels to the GPU box).  usage: tools/mkvar.sh name "-DJPGPU_PHASE_SYMBOLS=4 ..."
# Switches: JPGPU_PHASE_SYMBOLS, JPGPU_PF_BYTES, JPGPU_PIECE_SHIFT, JPGPU_SEQ_THREADS, JPGPU_TILES_PER_CTA, JPGPU_IDCT_MIN_CTAS.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $2 -shared -o build/variants/$1.so jpeg_rust_b200/csrc/jpgpu_kernels.cu jpeg_rust_b200/csrc/jpgpu_api.cu jpeg_rust_b200/csrc/jpgpu_host.cpp jpeg_rust_b200/csrc/jpgpu_multi.cpp -lcudart -lpthread 2>&1 | grep -E "error" || true
ls -la build/variants/$1.so | awk '{print $5, $9}'
