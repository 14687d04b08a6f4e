#!/bin/bash
# round-2 GPU session 29: ncu --set full of the shipped tensor-core-leak conv1 (bit-packed output, 148 samples per launch)
mkdir -p gpurun_out
SNNQP_C1_MODES=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv1_tclif -s 5 -c 1 -o gpurun_out/r2_tclif -f python tools/time_conv1.py 148 2 > gpurun_out/r2_tclif_ncu.log 2>&1; tail -2 gpurun_out/r2_tclif_ncu.log
