#!/bin/bash
# round-2 GPU session 34: 3x3 tile kernel with 7 (shipped) vs 8 weight taps resident in TMEM (libsnnqp_t8.so)
mkdir -p gpurun_out
for v in "" t8; do
  lib=$PWD/snnquantprune_b200/libsnnqp${v:+_$v}.so
  SNNQP_LIB=$lib timeout 200 python tools/time_conv2.py 296 10 | sed "s/^/${v:-shipped}: /"
  SNNQP_LIB=$lib timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "binary_bit_exact" 2>&1 | tail -1
done | tee gpurun_out/r2_conv2_tmem_taps.txt
