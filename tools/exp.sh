#!/bin/bash
# Runs on the GPU box (gpurun -- 'bash tools/exp.sh'): benches every variant library under build/variants/ and every
# environment variant listed in build/envs.txt (lines: <label> VAR=value ...; run-time switches: JPGPU_SUBSEQ_BITS,
# JPGPU_LOOKBACK_BITS, JPGPU_SEG_BITS, JPGPU_GROUPS, JPGPU_WRITE_PARTS, JPGPU_INTERVAL_MODE) and prints one line each:
# value (Mpixel/s), ms per step, per-kernel ms.  This is how the sweeps quoted in DESIGN.md were measured.
cd $GRAFT_REPO_ROOT
cp jpeg_rust_b200/lib/libjpgpu.so /tmp/orig.so
run() {  # label, extra env...
  label=$1; shift
  env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-parity $BENCH_ARGS > /tmp/b.json 2>/tmp/b.err
  python - "$label" <<PY
import json,sys
try:
    d=json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
    print(sys.argv[1].ljust(28), round(d["value"]), round(d["ms_per_step"],3), {n[:8]:round(v["ms"],3) for n,v in d["roofline"]["kernels"].items()}, [round(x,3) for x in d["roofline"]["kernels"]["prepass_count+scan+write"].get("split_ms",[])])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("/tmp/b.err").read()[-300:])
PY
}
shopt -s nullglob
for so in build/variants/*.so; do
  cp $so jpeg_rust_b200/lib/libjpgpu.so
  run $(basename $so .so) X=1
done
cp /tmp/orig.so jpeg_rust_b200/lib/libjpgpu.so
if [ -f build/envs.txt ]; then
  while read -r label envs; do [ -n "$label" ] && run $label $envs; done < build/envs.txt
fi
