#!/bin/bash
# round-2 GPU session 21: tensor-core-leak conv1 after the epilogue / producer-wait changes
mkdir -p gpurun_out
SNNQP_C1_MODES=0,2,203 timeout 240 python tools/time_conv1.py 296 10 > gpurun_out/r2_conv1_tclif.txt 2>&1; echo "exit $?" >> gpurun_out/r2_conv1_tclif.txt
cat gpurun_out/r2_conv1_tclif.txt | cut -c1-250
