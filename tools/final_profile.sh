#!/bin/bash
# Evidence run (1 GPU, tools/gpu.sh -- 'bash tools/final_profile.sh <tag>'): GPU tests, both bench arms, ncu launch list
# of the bench command, ncu --set full of one group's kernel chain (second decode, graphs off so that launches arrive
# in API order), the IDCT/colour kernel of every sampling mode, compute-sanitizer.  Outputs land in gpurun_out/<tag>_*;
# profiles/summarize_ncu.py turns the .ncu-rep files into the summaries committed under profiles/.
cd $GRAFT_REPO_ROOT
T=${1:-r02z}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_1024img.json 2> gpurun_out/${T}_bench.err; tail -c 400 gpurun_out/${T}_bench_1024img.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench_reference.json
JPGPU_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_1024img.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-parity > gpurun_out/${T}_ncu_bench.log 2>&1
JPGPU_GRAPH=0 timeout 500 ncu --set full --clock-control none --import-source on -s 26 -c 8 -o gpurun_out/${T}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-parity > gpurun_out/${T}_ncu_full.log 2>&1; tail -3 gpurun_out/${T}_ncu_full.log | cut -c1-200
timeout 200 compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py 2>&1 | tail -3 > gpurun_out/${T}_sanitizer_memcheck.log; cat gpurun_out/${T}_sanitizer_memcheck.log
for tool in initcheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool python tests/sanitizer_smoke.py 2>&1 | tail -3 > gpurun_out/${T}_sanitizer_$tool.log; cat gpurun_out/${T}_sanitizer_$tool.log
done
bash tools/idct_variants.sh $T
