#!/bin/bash
# round-2 GPU session 14: shared-space expander stores: conv tests, conv2 timing, full tests, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "spiking_conv_binary or block_sparse or tile_skip" 2>&1 | tail -2
timeout 200 python tools/time_conv2.py 296 10
timeout 900 python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest14.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest14.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest14.log | tail -6 | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err; cut -c1-200 gpurun_out/r2_bench14.json; tail -2 gpurun_out/r2_bench14.err
