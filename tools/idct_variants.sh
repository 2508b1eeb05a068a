#!/bin/bash
# Runs on the GPU box: the IDCT/colour kernel of every sampling mode on 512 x 1080p - bench line (roofline.frac against
# the measured HBM peak) and one `ncu --set full` capture of the kernel each.  usage: bash tools/idct_variants.sh <tag>
cd $GRAFT_REPO_ROOT
T=${1:-r02}
for sub in 444 422 440 gray 420; do
  timeout 300 python bench.py --subsampling $sub --images 512 --distinct 64 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-parity > gpurun_out/${T}_idct_$sub.json 2> gpurun_out/${T}_idct_$sub.err
  python - $sub gpurun_out/${T}_idct_$sub.json <<PY
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1].ljust(5), "value", round(d["value"]), "idct ms", round(r["ms_per_launch"],3), "GB/s", round(r["achieved"]), "frac", round(r["frac"],3), "of 8TB/s", round(r["frac_of_nominal_8TBps"],3))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:idct_colour_kernel -s 4 -c 1 -o gpurun_out/${T}_idct_$sub python bench.py --subsampling $sub --images 512 --distinct 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-parity > gpurun_out/${T}_idct_${sub}_ncu.log 2>&1
done
