#!/bin/bash
# round-2 GPU session 8: pad-free tile kernel bring-up, full GPU test suite, sparsity-path timings, bench, ncu launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "spiking_conv_binary or block_sparse or tile_skip" > gpurun_out/r2_gputest8a.log 2>&1; rc=$?
tail -4 gpurun_out/r2_gputest8a.log | cut -c1-300
if [ "$rc" != "0" ]; then echo "TILE KERNEL FAILED (rc $rc): falling back to the strip kernel for the rest of this session"; grep -E "^E  " gpurun_out/r2_gputest8a.log | head -8 | cut -c1-300; export SNNQP_CONV_STRIPS=1; fi
timeout 200 python tools/time_conv2.py 296 10 | tee gpurun_out/r2_conv2_tile.log
SNNQP_CONV_STRIPS=1 timeout 200 python tools/time_conv2.py 296 10 | tee -a gpurun_out/r2_conv2_tile.log
timeout 900 python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest8.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest8.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest8.log | tail -12 | cut -c1-300
timeout 300 python tools/time_sparse_paths.py 296 > gpurun_out/r2_sparse_paths.jsonl 2> gpurun_out/r2_sparse_paths.err; cut -c1-420 gpurun_out/r2_sparse_paths.jsonl
timeout 300 python tools/time_layers.py 592 296 > gpurun_out/r2_layers8.log 2>&1; cat gpurun_out/r2_layers8.log
timeout 900 python bench.py > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; cut -c1-1200 gpurun_out/r2_bench8.json; tail -3 gpurun_out/r2_bench8.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --global-batch 592 > gpurun_out/r2_bench_small.json 2>/dev/null &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --global-batch 592 > gpurun_out/r2_ncu_launches.log 2>&1
tail -2 gpurun_out/r2_ncu_launches.log | cut -c1-300; wc -l gpurun_out/r2_ncu_launches.csv
