#!/bin/bash
# round-2 GPU session 1: tests, baseline layer times, conv2 TMEM/smem split experiment
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest1.log
python tools/time_layers.py 592 296 > gpurun_out/r2_layers1.log 2>&1
for s in 0 72 52 42 33 32 2; do
  if [ "$s" = "0" ]; then python tools/time_conv2.py 296 10; else SNNQP_C2_SPLIT=$s python tools/time_conv2.py 296 10; fi
done > gpurun_out/r2_conv2_split.log 2>&1
tail -5 gpurun_out/r2_gputest1.log; cat gpurun_out/r2_layers1.log gpurun_out/r2_conv2_split.log
