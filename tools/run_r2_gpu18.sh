#!/bin/bash
# round-2 GPU session 18: final validation + final bench of the committed tree
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest18.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest18.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest18.log | tail -6 | cut -c1-300
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-100
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; cut -c1-200 gpurun_out/r2_bench_final.json; tail -2 gpurun_out/r2_bench_final.err
timeout 300 python tools/time_layers.py 592 296 > gpurun_out/r2_layers18.log 2>&1; cat gpurun_out/r2_layers18.log
timeout 300 python tools/time_sparse_paths.py 296 > gpurun_out/r2_sparse_paths.jsonl 2> gpurun_out/r2_sparse_paths.err; cut -c1-330 gpurun_out/r2_sparse_paths.jsonl
