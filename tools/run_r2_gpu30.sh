#!/bin/bash
# round-2 GPU session 30: LIF_TENSOR conv1 with slot-pair hand-offs and uniform TMEM addresses: time, flips, parity tests
mkdir -p gpurun_out
SNNQP_C1_MODES=0,2 timeout 240 python tools/time_conv1.py 296 10 2>&1 | grep "lif_mode=2" | tee gpurun_out/r2_conv1_tclif_pairs.txt
SNNQP_C1_STABLE=1 SNNQP_C1_MODES=0,2 timeout 240 python tools/time_conv1.py 296 10 2>&1 | grep "lif_mode=2" | sed 's/^/stable: /' | tee -a gpurun_out/r2_conv1_tclif_pairs.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "lif_tensor or production_shape or densities" 2>&1 | tail -4 | cut -c1-250
