#!/bin/bash
# round-2 GPU session 3: full GPU test suite, layer times, quick bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest3.log
timeout 300 python tools/time_layers.py 592 296 > gpurun_out/r2_layers3.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_gputest3.log | tail -15; cat gpurun_out/r2_layers3.log; cat gpurun_out/r2_bench3.json; tail -5 gpurun_out/r2_bench3.err
