#!/bin/bash
# round-2 GPU session 27: split-weight LIF_TENSOR (shipped form): flips / times on both weight sets, network-level effect
mkdir -p gpurun_out
SNNQP_C1_MODES=0,103,2 timeout 240 python tools/time_conv1.py 296 10 2>&1 | grep "bits=1" | tee gpurun_out/r2_conv1_tclif.txt
SNNQP_C1_STABLE=1 SNNQP_C1_MODES=0,103,2 timeout 240 python tools/time_conv1.py 296 10 2>&1 | grep "bits=1" | sed 's/^/stable: /' | tee -a gpurun_out/r2_conv1_tclif.txt
for a in "20 0" "20 1"; do timeout 120 python tools/probe_tclif_error.py $a 2>&1 | tail -3; done | tee gpurun_out/r2_tclif_error.txt
timeout 300 python tools/probe_lif_modes_logits.py 592 2>&1 | tail -4 | tee gpurun_out/r2_lif_modes_logits.txt
