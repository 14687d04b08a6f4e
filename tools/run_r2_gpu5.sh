#!/bin/bash
# round-2 GPU session 5: fused head kernel bring-up
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_head or block_sparse or tile_skip" > gpurun_out/r2_gputest5.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest5.log
tail -25 gpurun_out/r2_gputest5.log
timeout 300 python tools/time_layers.py 592 296 > gpurun_out/r2_layers5.log 2>&1; cat gpurun_out/r2_layers5.log
timeout 300 python tools/time_layers.py 1184 148 > gpurun_out/r2_layers5b.log 2>&1; tail -3 gpurun_out/r2_layers5b.log
timeout 300 python tools/time_sparse_paths.py 296 > gpurun_out/r2_sparse_paths.jsonl 2> gpurun_out/r2_sparse_paths.err; cat gpurun_out/r2_sparse_paths.jsonl
