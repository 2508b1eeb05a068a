#!/bin/bash
# Evidence run (1 GPU, gpurun -- 'bash tools/final_profile.sh'): tests, both bench arms, ncu launch list of the bench
# command, ncu --set full of the seven full-batch kernels (launches 42..48 of this command line), config-4 line,
# compute-sanitizer.  Outputs land in gpurun_out/<T>_*; profiles/summarize_ncu.py turns the .ncu-rep into the summary.
cd $GRAFT_REPO_ROOT
T=r01za
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_1024img.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench_1024img.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_launches_1024img.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -s 42 -c 7 -o gpurun_out/${T}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_full.log 2>&1; tail -3 gpurun_out/${T}_ncu_full.log | cut -c1-200
timeout 300 python bench.py --images 64 --distinct 16 --width 3840 --height 2160 --subsampling 444 --restart-interval 16 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${T}_bench_4k444_dri16.json 2>> gpurun_out/${T}_bench.err; tail -c 400 gpurun_out/${T}_bench_4k444_dri16.json; echo
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --waves 16 > gpurun_out/${T}_bench_waves16.json 2>> gpurun_out/${T}_bench.err; tail -c 400 gpurun_out/${T}_bench_waves16.json; echo
timeout 200 compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py 2>&1 | tail -3 > gpurun_out/${T}_sanitizer_memcheck.log; cat gpurun_out/${T}_sanitizer_memcheck.log
