#!/bin/bash
# round-2 GPU session 2: bit-packed spikes + conv1 LIF variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest2.log
timeout 300 python tools/time_conv1.py 296 10 > gpurun_out/r2_conv1_variants.log 2>&1
timeout 300 python tools/time_layers.py 592 296 > gpurun_out/r2_layers2.log 2>&1
tail -15 gpurun_out/r2_gputest2.log; cat gpurun_out/r2_conv1_variants.log gpurun_out/r2_layers2.log
