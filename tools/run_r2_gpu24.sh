#!/bin/bash
# round-2 GPU session 24: LIF_TENSOR as the engine default: new parity tests, full GPU suite, smoke, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "lif_tensor or full_size or production_shape" > gpurun_out/r2_gputest24a.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest24a.log
grep -E "passed|failed|FAILED|^E  |Error" gpurun_out/r2_gputest24a.log | tail -12 | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest24.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest24.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest24.log | tail -8 | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | cut -c1-200
timeout 900 python bench.py > gpurun_out/r2_bench_tclif.json 2> gpurun_out/r2_bench_tclif.err; cut -c1-250 gpurun_out/r2_bench_tclif.json; tail -3 gpurun_out/r2_bench_tclif.err
