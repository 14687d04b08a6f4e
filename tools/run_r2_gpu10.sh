#!/bin/bash
# round-2 GPU session 10: final N=1 numbers -- full bench with CPU baseline, reference arm, ncu launch list, ncu --set full of the head kernels
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "block_sparse" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; cut -c1-900 gpurun_out/r2_bench_final.json; tail -2 gpurun_out/r2_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-600 gpurun_out/r2_bench_reference_arm.json
for cfg in "--bits 4 --prune 0.8 --batch 256" "--bits 2 --prune 0.9 --batch 512" "--T 10"; do
  timeout 300 python bench.py $cfg --steps 10 --no-cpu-baseline 2>/dev/null | cut -c1-260
done | tee gpurun_out/r2_bench_other_configs.jsonl
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --global-batch 592 > gpurun_out/r2_bench_small.json 2>/dev/null &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --global-batch 592 > gpurun_out/r2_ncu_launches.log 2>&1
wc -l gpurun_out/r2_ncu_launches.csv
timeout 120 python tools/prof_kernels.py 148 > gpurun_out/r2_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_conv1_umma|k_conv3x3_tile" -s 5 -c 5 \
    -o gpurun_out/r2_prof_final -f python tools/prof_kernels.py 148 > gpurun_out/r2_prof_ncu.log 2>&1
tail -2 gpurun_out/r2_prof_ncu.log; ls -la gpurun_out/r2_prof_final.ncu-rep
