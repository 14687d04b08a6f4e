#!/bin/bash
# round-2 GPU session 7: ncu --set full captures (with source) of the stand-alone conv1 / conv2 kernels and the fused head
mkdir -p gpurun_out
timeout 120 python tools/prof_kernels.py 148 > gpurun_out/r2_prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_prof_plain.log; exit 0; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_conv1_umma|k_conv3x3_umma|k_head_fused" -s 3 -c 3 \
    -o gpurun_out/r2_prof_head -f python tools/prof_kernels.py 148 > gpurun_out/r2_prof_ncu.log 2>&1
tail -5 gpurun_out/r2_prof_ncu.log; ls -la gpurun_out/r2_prof_head.ncu-rep
