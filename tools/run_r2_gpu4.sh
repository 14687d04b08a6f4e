#!/bin/bash
# round-2 GPU session 4: full GPU test suite again, sparsity-path timings, bench with CPU baseline, ncu launch list + full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest4.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest4.log
timeout 600 python tools/time_sparse_paths.py 296 > gpurun_out/r2_sparse_paths.jsonl 2> gpurun_out/r2_sparse_paths.err
timeout 900 python bench.py > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest4.log | tail -15; cat gpurun_out/r2_sparse_paths.jsonl; tail -3 gpurun_out/r2_sparse_paths.err; cat gpurun_out/r2_bench4.json | cut -c1-6000; tail -3 gpurun_out/r2_bench4.err
