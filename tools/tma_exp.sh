#!/bin/bash
# GPU box: the IDCT/colour kernel with bulk-asynchronous coefficient loads (build/variants/tma.so, -DJPGPU_IDCT_TMA=1)
# against the shipped kernel, 512 x 1080p, every sampling mode.
cd $GRAFT_REPO_ROOT
cp jpeg_rust_b200/lib/libjpgpu.so /tmp/orig.so
for lib in orig tma; do
  [ $lib = tma ] && cp build/variants/tma.so jpeg_rust_b200/lib/libjpgpu.so
  for sub in 420 444 422 gray; do
    timeout 200 python bench.py --subsampling $sub --images 512 --distinct 64 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$lib $sub', 'value', round(d['value']), 'idct ms', round(r['ms_per_launch'],3), 'frac', round(r['frac'],3))"
  done
done
cp /tmp/orig.so jpeg_rust_b200/lib/libjpgpu.so
