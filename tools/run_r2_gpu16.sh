#!/bin/bash
# round-2 GPU session 16: conv1 producer waits: nanosleep poll (shipped) vs hardware-suspended try_wait (libsnnqp_suspend.so)
mkdir -p gpurun_out
timeout 200 python tools/time_conv1.py 296 10 | grep "lif_mode=103" | tee gpurun_out/r2_conv1_wait_modes.txt
SNNQP_LIB=$PWD/snnquantprune_b200/libsnnqp_suspend.so timeout 200 python tools/time_conv1.py 296 10 | grep "lif_mode=103" | sed 's/^/suspend: /' | tee -a gpurun_out/r2_conv1_wait_modes.txt
