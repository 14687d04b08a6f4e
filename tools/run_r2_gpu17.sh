#!/bin/bash
# round-2 GPU session 17: number of expander warps of the bit-packed tile kernel (A/B builds via SNNQP_LIB)
mkdir -p gpurun_out
for n in 4 3 2 6; do
  if [ "$n" = "4" ]; then lib=$PWD/snnquantprune_b200/libsnnqp.so; else lib=$PWD/snnquantprune_b200/libsnnqp_exp$n.so; fi
  SNNQP_LIB=$lib timeout 200 python tools/time_conv2.py 296 10 | sed "s/^/expander warps $n: /"
done | tee gpurun_out/r2_conv2_expander_warps.txt
