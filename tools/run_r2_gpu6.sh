#!/bin/bash
# round-2 GPU session 6: fused head kernel bring-up, short timeouts (a deadlocked kernel must not eat the budget)
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_head" > gpurun_out/r2_gputest6.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/r2_gputest6.log
tail -25 gpurun_out/r2_gputest6.log | cut -c1-300
if [ "$rc" != "0" ]; then echo "fused head failed: skipping the timing runs"; exit 0; fi
timeout 200 python tools/time_layers.py 592 296 > gpurun_out/r2_layers6.log 2>&1; cat gpurun_out/r2_layers6.log
timeout 200 python tools/time_layers.py 1184 148 > gpurun_out/r2_layers6b.log 2>&1; tail -6 gpurun_out/r2_layers6b.log
timeout 300 python tools/time_sparse_paths.py 296 > gpurun_out/r2_sparse_paths.jsonl 2> gpurun_out/r2_sparse_paths.err; cat gpurun_out/r2_sparse_paths.jsonl
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "block_sparse or tile_skip" 2>&1 | tail -3
