#!/bin/bash
# round-2 GPU session 9: tile kernel with 16 epilogue warps
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "spiking_conv_binary or block_sparse or tile_skip" > gpurun_out/r2_gputest9a.log 2>&1; rc=$?
tail -4 gpurun_out/r2_gputest9a.log | cut -c1-300
if [ "$rc" != "0" ]; then grep -E "^E  " gpurun_out/r2_gputest9a.log | head -8 | cut -c1-300; exit 0; fi
timeout 200 python tools/time_conv2.py 296 10 | tee gpurun_out/r2_conv2_tile16.log
timeout 300 python tools/time_sparse_paths.py 296 > gpurun_out/r2_sparse_paths.jsonl 2> gpurun_out/r2_sparse_paths.err; cut -c1-420 gpurun_out/r2_sparse_paths.jsonl
timeout 300 python tools/time_layers.py 592 296 > gpurun_out/r2_layers9.log 2>&1; cat gpurun_out/r2_layers9.log
timeout 900 python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest9.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest9.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest9.log | tail -12 | cut -c1-300
