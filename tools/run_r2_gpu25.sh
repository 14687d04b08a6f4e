#!/bin/bash
# round-2 GPU session 25: conv1 flip rates and times of the LIF modes on both synthetic weight sets
# (SNNQP_C1_STABLE=1: the StableRNG weights / frames of the production-shape test)
mkdir -p gpurun_out
SNNQP_C1_MODES=0,103,2 timeout 240 python tools/time_conv1.py 296 10 2>&1 | grep "bits=1" | tee gpurun_out/r2_conv1_tclif.txt
SNNQP_C1_STABLE=1 SNNQP_C1_MODES=0,103,2 timeout 240 python tools/time_conv1.py 296 10 2>&1 | grep "bits=1" | sed 's/^/stable: /' | tee -a gpurun_out/r2_conv1_tclif.txt
for a in "20 0" "20 1"; do timeout 120 python tools/probe_tclif_error.py $a 2>&1 | tail -3; done | tee gpurun_out/r2_tclif_error.txt
