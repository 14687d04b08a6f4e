#!/bin/bash
# round-2 GPU session 28: LIF_TENSOR (split-weight form) as the engine default: full GPU suite, smoke, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest28.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest28.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest28.log | tail -10 | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2_bench_tclif.json 2> gpurun_out/r2_bench_tclif.err; cut -c1-250 gpurun_out/r2_bench_tclif.json; tail -3 gpurun_out/r2_bench_tclif.err
