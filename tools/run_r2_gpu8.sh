#!/bin/bash
# round-2 GPU session 8: full GPU test suite, sparsity-path timings, bench (with CPU baseline), ncu launch list of a short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest8.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest8.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest8.log | tail -12 | cut -c1-300
timeout 300 python tools/time_sparse_paths.py 296 > gpurun_out/r2_sparse_paths.jsonl 2> gpurun_out/r2_sparse_paths.err; cut -c1-400 gpurun_out/r2_sparse_paths.jsonl
timeout 300 python tools/time_layers.py 592 296 > gpurun_out/r2_layers8.log 2>&1; cat gpurun_out/r2_layers8.log
timeout 900 python bench.py > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; cut -c1-1500 gpurun_out/r2_bench8.json; tail -3 gpurun_out/r2_bench8.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --global-batch 592 > gpurun_out/r2_bench_small.json 2>/dev/null &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --global-batch 592 > gpurun_out/r2_ncu_launches.log 2>&1
tail -2 gpurun_out/r2_ncu_launches.log | cut -c1-300; wc -l gpurun_out/r2_ncu_launches.csv
