#!/bin/bash
# round-2 GPU session 19: first run of the tensor-core-leak conv1 (SNNQP_LIF_TENSOR): flips vs exact + timing
mkdir -p gpurun_out
SNNQP_C1_MODES=0,103,2,203 timeout 240 python tools/time_conv1.py 296 10 > gpurun_out/r2_conv1_tclif.txt 2>&1; echo "exit $?" >> gpurun_out/r2_conv1_tclif.txt
cat gpurun_out/r2_conv1_tclif.txt | cut -c1-250
