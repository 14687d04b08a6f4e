#!/bin/bash
# round-2 GPU session 11: validate the deeper operand ring + conv1 select tree: tests, layer times, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest11.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest11.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest11.log | tail -8 | cut -c1-300
timeout 200 python tools/time_conv2.py 296 10
timeout 300 python tools/time_conv1.py 296 10 | grep "lif_mode=103\|lif_mode=0"
timeout 300 python tools/time_layers.py 592 296 > gpurun_out/r2_layers11.log 2>&1; cat gpurun_out/r2_layers11.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err; cut -c1-200 gpurun_out/r2_bench11.json; tail -2 gpurun_out/r2_bench11.err
python __graft_entry__.py smoke 2>&1 | tail -2
