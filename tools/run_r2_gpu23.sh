#!/bin/bash
# round-2 GPU session 23 (historical: the SNNQP_TC_SAT / SNNQP_TC_POOLF build macros it compared were removed after the measurement): LIF_TENSOR epilogue variants: compares 1 FSET + 3 FFMA.SAT (shipped) vs 4 FFMA.SAT; pool by LOP3 OR (shipped) vs FADD2 + FADD.SAT
mkdir -p gpurun_out; : > gpurun_out/r2_conv1_tclif_variants.txt
for v in "" s4p s3p1 s4p1; do
  lib=$PWD/snnquantprune_b200/libsnnqp${v:+_$v}.so
  SNNQP_LIB=$lib SNNQP_C1_MODES=2,203 timeout 200 python tools/time_conv1.py 296 10 | grep "bits=1" | sed "s/^/${v:-shipped}: /" | tee -a gpurun_out/r2_conv1_tclif_variants.txt
done
