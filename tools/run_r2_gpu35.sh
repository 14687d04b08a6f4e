#!/bin/bash
# round-2 GPU session 35: 3x3 tile kernel with one MMA-issuing warp (libsnnqp_m1.so) vs two (default build)
mkdir -p gpurun_out
for v in m1 ""; do
  lib=$PWD/snnquantprune_b200/libsnnqp${v:+_$v}.so
  SNNQP_LIB=$lib timeout 200 python tools/time_conv2.py 296 10 | sed "s/^/${v:-two issuers}: /"
  SNNQP_LIB=$lib timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "binary_bit_exact or sparse or skip or packed" 2>&1 | tail -1
done | tee gpurun_out/r2_conv2_mma_issuers.txt
timeout 200 python tools/time_layers.py 592 296 2>&1 | head -4
