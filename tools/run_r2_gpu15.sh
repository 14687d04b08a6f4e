#!/bin/bash
# round-2 GPU session 15: validation of the committed tree
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail 6 > gpurun_out/r2_gputest15.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputest15.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_gputest15.log | tail -8 | cut -c1-300
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-100
timeout 200 python tools/time_conv1.py 296 10 | grep "bits=1 lif_mode=103"
