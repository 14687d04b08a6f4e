#!/bin/bash
# round-2 GPU session 12: final validation of the tree as committed: GPU tests, smoke, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest12.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest12.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest12.log | tail -8 | cut -c1-300
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; cut -c1-200 gpurun_out/r2_bench12.json; tail -2 gpurun_out/r2_bench12.err
