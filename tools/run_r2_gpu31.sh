#!/bin/bash
# round-2 GPU session 31: final validation of the LIF_TENSOR default + the profile artefacts of the same tree
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest31.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest31.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest31.log | tail -6 | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-120
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; cut -c1-200 gpurun_out/r2_bench_final.json; tail -2 gpurun_out/r2_bench_final.err
timeout 300 python tools/time_layers.py 592 296 > gpurun_out/r2_layers31.log 2>&1; cat gpurun_out/r2_layers31.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --global-batch 592 > gpurun_out/r2_bench_small.json 2>/dev/null &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --global-batch 592 > gpurun_out/r2_ncu_launches.log 2>&1
wc -l gpurun_out/r2_ncu_launches.csv
timeout 120 python tools/prof_kernels.py 148 > gpurun_out/r2_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_conv1_tclif|k_conv1_umma|k_conv3x3_tile" -s 5 -c 5 \
    -o gpurun_out/r2_prof_final -f python tools/prof_kernels.py 148 > gpurun_out/r2_prof_ncu.log 2>&1
tail -2 gpurun_out/r2_prof_ncu.log; ls -la gpurun_out/r2_prof_final.ncu-rep
