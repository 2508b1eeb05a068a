#!/bin/bash
# Host-only robustness run (no GPU): builds jpgpu_host.cpp + tools/asan_fuzz_host.cpp with AddressSanitizer and UBSan and
# feeds jpgpu_parse / jpgpu_parse_scans / jpgpu_geometry / jpgpu_plan_info with mutated headers of the given files under every
# extension and layout.  usage: tools/asan_fuzz_host.sh [iterations] [seed] [files...]   (default: synthetic files + fixtures)
set -e
cd "$(dirname "$0")/.."
B=build/asan; mkdir -p $B
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -Iinclude -Ijpeg_rust_b200/csrc -I/usr/local/cuda/include \
    tools/asan_fuzz_host.cpp jpeg_rust_b200/csrc/jpgpu_host.cpp -o $B/fuzz_host -lpthread
N=${1:-5000}; SEED=${2:-1}; shift 2 2>/dev/null || true
if [ $# -eq 0 ]; then
  python - <<PY
import sys; sys.path.insert(0, ".")
from jpeg_rust_b200 import synth
open("$B/a420.jpg", "wb").write(synth.synth_jpeg(1, 48, 32, "420"))
open("$B/a444dri.jpg", "wb").write(synth.synth_jpeg(2, 40, 24, "444", restart_interval=2))
open("$B/agray.jpg", "wb").write(synth.synth_jpeg(3, 33, 17, "gray", optimize=True))
open("$B/aplanar.jpg", "wb").write(synth.synth_jpeg(4, 48, 32, "420", planar_scans=True))
PY
  set -- $B/a420.jpg $B/a444dri.jpg $B/agray.jpg $B/aplanar.jpg tests/golden/fixtures/huff_simple0.jpg
fi
$B/fuzz_host $N $SEED "$@"
