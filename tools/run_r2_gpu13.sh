#!/bin/bash
# round-2 GPU session 13: final validation: GPU tests, smoke, full bench, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest13.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest13.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest13.log | tail -8 | cut -c1-300
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-120
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; cut -c1-200 gpurun_out/r2_bench_final.json; tail -2 gpurun_out/r2_bench_final.err
timeout 300 python tools/time_layers.py 592 296 > gpurun_out/r2_layers13.log 2>&1; cat gpurun_out/r2_layers13.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --global-batch 592 > gpurun_out/r2_bench_small.json 2>/dev/null &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --global-batch 592 > gpurun_out/r2_ncu_launches.log 2>&1
wc -l gpurun_out/r2_ncu_launches.csv
