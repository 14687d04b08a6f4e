#!/bin/bash
# round-2 GPU session 26: membrane error of the tensor-core-leak conv1 vs the reference op order (accumulation precision of tcgen05 kind::f16)
mkdir -p gpurun_out
for a in "1 0" "2 0" "20 0" "20 1"; do timeout 120 python tools/probe_tclif_error.py $a 2>&1 | tail -3; done | tee gpurun_out/r2_tclif_error.txt
