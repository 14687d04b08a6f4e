#!/bin/bash
# round-2 GPU session 32: bench lines of the final build: N=1 default, the other BASELINE.json configs, the reference arm
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; cut -c1-200 gpurun_out/r2_bench_final.json; tail -2 gpurun_out/r2_bench_final.err
for cfg in "--bits 4 --prune 0.8 --batch 256" "--bits 2 --prune 0.9 --batch 512" "--T 10" "--lif exact" "--lif fast"; do
  timeout 300 python bench.py $cfg --steps 10 --no-cpu-baseline 2>/dev/null | cut -c1-330
done | tee gpurun_out/r2_bench_other_configs.jsonl
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
